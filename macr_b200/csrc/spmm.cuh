// spmm.cuh -- internal declarations of the LightGCN propagation kernels (spmm.cu)
#pragma once
#include "train_kernels.cuh"

namespace macr {

// a dense [rows][64] operand that may live in two allocations (E0 = concat(user table, item
// table) is never materialised): rows < split come from a, the rest from b
struct RowSrc {
  const float *a, *b;
  long long split;
};

// Y = A X (+ add); optionally acc_out = (acc_in + Y) (/ acc_div when > 0)
int launch_spmm(const int32_t *rowptr, const int32_t *col, const float *val, int64_t n_rows,
                RowSrc X, const float *add, float *Y, RowSrc acc_in, float *acc_out,
                float acc_div, cudaStream_t s);
int launch_lgcn_propagate(const int32_t *rowptr, const int32_t *col, const float *val,
                          const float *U, int64_t n_users, const float *I, int64_t n_items,
                          int n_layers, float *Emean, float *tmp, cudaStream_t s);
int launch_scatter_rows(PlanBufs planU, const float *gU, PlanBufs planI, const float *gI, int B,
                        int64_t n_users, float div, float *out, cudaStream_t s);
int launch_l2_rows(PlanBufs planU, PlanBufs planI, int B, const float *U, const float *I,
                   int64_t n_users, float lam, float *grad, cudaStream_t s);

}  // namespace macr
