// ref_evaluator_shim.cpp -- builds the REFERENCE's own C++ evaluator into oracle/_ref/.
// TEST INFRASTRUCTURE ONLY.  No reference source is copied: the two headers are #included
// from where they lie under /root/reference (-I given by oracle/Makefile), exactly as the
// reference's Cython glue does (macr_lightgcn/evaluator/cpp/apt_evaluate_foldout.pyx:11-20).
// The shim only gives the two functions C linkage so ctypes can call them.
#include "tools.h"             // c_top_k_array_index   (tools.h:24-33)
#include "evaluate_foldout.h"  // evaluate_foldout      (evaluate_foldout.h:115-195)

extern "C" {
void ref_c_top_k_array_index(float *scores_pt, int columns_num, int rows_num, int top_k,
                             int thread_num, int *rankings_pt) {
  c_top_k_array_index(scores_pt, columns_num, rows_num, top_k, thread_num, rankings_pt);
}
void ref_evaluate_foldout(int users_num, int *rankings, int rank_len, int **ground_truths,
                          int *ground_truths_num, int thread_num, float *results) {
  evaluate_foldout(users_num, rankings, rank_len, ground_truths, ground_truths_num, thread_num,
                   results);
}
}
