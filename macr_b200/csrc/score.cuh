// score.cuh -- pieces of the scoring path shared by score.cu (exact fp32) and score_tc.cu (tcgen05)
#pragma once
#include "common.cuh"

namespace macr {

// the order rule of every top-K list: score descending, ties -> lower item id first
__device__ __forceinline__ bool score_better(float sa, int ia, float sb, int ib) {
  return sa > sb || (sa == sb && ia < ib);
}

// sorted-list insert, list held one rank per lane; (cs,cid) warp-uniform
__device__ __forceinline__ void score_list_insert(float &ls, int &li, float cs, int cid, int lane,
                                                  unsigned kmask, int K) {
  const unsigned bal = __ballot_sync(0xffffffffu, score_better(ls, li, cs, cid)) & kmask;
  const int pos = __popc(bal);
  const float us = __shfl_up_sync(0xffffffffu, ls, 1);
  const int ui = __shfl_up_sync(0xffffffffu, li, 1);
  if (pos < K) {
    if (lane == pos) {
      ls = cs;
      li = cid;
    } else if (lane > pos) {
      ls = us;
      li = ui;
    }
  }
}

// exact fp32 kernel restricted to the rows queued in row_map[0 .. *n_rows_dev) (device-side
// count: the launch is a no-op when the queue is empty); results are scattered to
// out_ids/out_scores[row_map[slot]].  Used by the tcgen05 path for overflowed rows.
size_t score_exact_workspace_bytes(int T, long long n_items, int K);
int score_exact_rows(const float *Uq, int T, const float *It, long long n_items,
                     const float *sig_i, const float *sig_u, float c, const int32_t *mask_rowptr,
                     const int32_t *mask_col, int K, int id_off, const int32_t *row_map,
                     const int *n_rows_dev, int32_t *out_ids, float *out_scores, void *ws,
                     size_t ws_bytes, cudaStream_t s);

}  // namespace macr
