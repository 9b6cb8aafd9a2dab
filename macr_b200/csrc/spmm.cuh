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

// Static work decomposition of one adjacency (built once, spmm.cu:build_spmm_plan): every row is
// cut into segments of at most kSegNnz nonzeros, one half-warp per segment, so popular items
// (27 504 nonzeros in ml_10m) no longer leave a few SMs working long after the grid has drained.
// Rows of one segment are finished in place; the others go through `partial` and a fixed-order
// combine pass (deterministic).
struct SpmmPlan {
  int n_seg = 0, n_multi = 0, n_wide = 0;
  int32_t *seg_row = nullptr, *seg_start = nullptr, *seg_end = nullptr, *seg_slot = nullptr;
  int32_t *multi_row = nullptr, *multi_slot0 = nullptr, *multi_nseg = nullptr;
  int32_t *wide_idx = nullptr;  // multi rows of more than kCombineWide segments: one CTA each
  float *partial = nullptr;  // [sum of segments of multi-segment rows][64]
};
// ranges (nullable): {a0, a1, b0, b1} -- only rows in [a0,a1) or [b0,b1) get segments (the rows a
// rank owns when the adjacency is row-partitioned); the segmentation of a row never depends on it
int build_spmm_plan(const int32_t *d_rowptr, int64_t n_rows, SpmmPlan *out,
                    const int64_t *ranges = nullptr);
void free_spmm_plan(SpmmPlan *p);

// Y = A X (+ add); optionally acc_out = (acc_in + Y) (/ acc_div when > 0).
// plan == nullptr: stateless kernel (one CTA per 16 rows).  x_nonzero (nullable): bitmap over
// the rows of X; nonzeros whose X row is flagged all-zero are skipped (row-sparse gradients).
// row_needed (nullable, planned path only): bitmap over the output rows; rows whose bit is clear are
// not computed at all (their outputs keep whatever they held) -- a training step reads the last
// forward layer at the batch's <= 3B rows only; the rows that are computed are bit-identical.
int launch_spmm_planned(const SpmmPlan *plan, const int32_t *rowptr, const int32_t *col,
                        const float *val, int64_t n_rows, RowSrc X, const float *add, float *Y,
                        RowSrc acc_in, float *acc_out, float acc_div, const uint32_t *x_nonzero,
                        cudaStream_t s, const uint32_t *row_needed = nullptr);
int launch_mark_batch_rows(const StepState *st, int B, int64_t n_users, uint32_t *bitmap, cudaStream_t s);
int launch_spmm(const int32_t *rowptr, const int32_t *col, const float *val, int64_t n_rows,
                RowSrc X, const float *add, float *Y, RowSrc acc_in, float *acc_out,
                float acc_div, cudaStream_t s);
int launch_lgcn_propagate(const int32_t *rowptr, const int32_t *col, const float *val,
                          const float *U, int64_t n_users, const float *I, int64_t n_items,
                          int n_layers, float *Emean, float *tmp, cudaStream_t s,
                          const SpmmPlan *plan = nullptr);
int launch_scatter_rows(PlanBufs planU, const float *gU, PlanBufs planI, const float *gI, int B,
                        int64_t n_users, float div, float *out, uint32_t *nz_bitmap, cudaStream_t s);
int launch_l2_rows(PlanBufs planU, PlanBufs planI, int B, const float *U, const float *I,
                   int64_t n_users, float lam, float *grad, cudaStream_t s);

}  // namespace macr
