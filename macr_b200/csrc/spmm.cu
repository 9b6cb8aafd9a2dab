// spmm.cu -- K6: CSR SpMM for LightGCN's propagation  E_{k+1} = A_hat E_k
// (macr_lightgcn/LightGCN.py:288-309; the reference runs it as 100 row folds of
// tf.sparse_tensor_dense_matmul per layer, :257-269,:297-305).
//
// Layout: A_hat as CSR (int32 rowptr / col, fp32 val), dense operand row-major [N][64] fp32.
// One half-warp per output row: the 16 lanes first load 16 (col,val) pairs with one coalesced
// access each, then broadcast them lane by lane so 16 independent 256-byte row gathers are in
// flight per half-warp; every lane owns one float4 (4 of the 64 columns) of the output row.
// HBM/L2-bound: algorithmic bytes = 8*nnz + 4*(N+1) + 8*N*d.
#include "spmm.cuh"

#include <algorithm>
#include <new>
#include <vector>

namespace macr {

__device__ __forceinline__ const float4 *row_ptr2(const RowSrc &s, long long r, int hl) {
  const float *base = (r < s.split) ? s.a + r * kD : s.b + (r - s.split) * kD;
  return reinterpret_cast<const float4 *>(base) + hl;
}

constexpr int kLongRow = 256;  // rows with more nonzeros are summed by the whole CTA

// Y[r] (+ epilogue) from the accumulated row `acc` (lane hl holds columns 4*hl .. 4*hl+3)
__device__ __forceinline__ void spmm_epilogue(float4 acc, long long r, int hl,
                                              const float *__restrict__ add, float *Y,
                                              const RowSrc &acc_in, float *acc_out, float acc_div) {
  if (add) {
    const float4 t = reinterpret_cast<const float4 *>(add + r * kD)[hl];
    acc.x += t.x;
    acc.y += t.y;
    acc.z += t.z;
    acc.w += t.w;
  }
  if (Y) reinterpret_cast<float4 *>(Y + r * kD)[hl] = acc;
  if (acc_out) {
    float4 z = *row_ptr2(acc_in, r, hl);
    z.x += acc.x;
    z.y += acc.y;
    z.z += acc.z;
    z.w += acc.w;
    if (acc_div > 0.f) {  // tf.reduce_mean over the stacked layers: sum / (L+1)
      z.x = __fdiv_rn(z.x, acc_div);
      z.y = __fdiv_rn(z.y, acc_div);
      z.z = __fdiv_rn(z.z, acc_div);
      z.w = __fdiv_rn(z.w, acc_div);
    }
    reinterpret_cast<float4 *>(acc_out + r * kD)[hl] = z;
  }
}

// nonzeros [start, end) taken in groups of 16 with stride `stride` (in nonzeros)
__device__ __forceinline__ float4 spmm_accumulate(const int32_t *__restrict__ col,
                                                  const float *__restrict__ val, int start, int end,
                                                  int stride, const RowSrc &X, int hl,
                                                  unsigned hmask) {
  float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
  for (int base = start; base < end; base += stride) {
    const int e = base + hl;
    const int c = e < end ? col[e] : 0;
    const float a = e < end ? val[e] : 0.f;
    const int cnt = min(16, end - base);
#pragma unroll 4
    for (int k = 0; k < cnt; ++k) {
      const int cc = __shfl_sync(hmask, c, k, 16);
      const float aa = __shfl_sync(hmask, a, k, 16);
      const float4 x = *row_ptr2(X, cc, hl);
      acc.x = fmaf(aa, x.x, acc.x);
      acc.y = fmaf(aa, x.y, acc.y);
      acc.z = fmaf(aa, x.z, acc.z);
      acc.w = fmaf(aa, x.w, acc.w);
    }
  }
  return acc;
}

// One CTA owns 16 consecutive rows.  Rows of at most kLongRow nonzeros: one half-warp per row,
// nonzeros in order (the oracle's chain).  Longer rows (popular items: 27 504 nonzeros in
// ml_10m) would leave one half-warp working long after the rest of the grid has drained, so
// the 16 half-warps of the CTA take their 16-nonzero groups round-robin and the 16 partial rows
// are summed through shared memory in a fixed order (deterministic; differs from the sequential
// chain only in rounding).
__global__ void __launch_bounds__(256)
spmm_csr_kernel(const int32_t *__restrict__ rowptr, const int32_t *__restrict__ col,
                const float *__restrict__ val, long long n_rows, RowSrc X,
                const float *__restrict__ add, float *Y, RowSrc acc_in, float *acc_out,
                float acc_div) {
  __shared__ int s_start[16], s_end[16];
  __shared__ float4 s_part[16][16];
  const int lane = threadIdx.x & 31, hl = lane & 15, hw = threadIdx.x >> 4;
  const unsigned hmask = (lane < 16) ? 0x0000ffffu : 0xffff0000u;
  const long long r0 = (long long)blockIdx.x * 16;
  const long long r = r0 + hw;
  int start = 0, end = 0;
  if (r < n_rows) {
    start = rowptr[r];
    end = rowptr[r + 1];
  }
  if (hl == 0) {
    s_start[hw] = start;
    s_end[hw] = end;
  }
  if (r < n_rows && end - start <= kLongRow) {
    const float4 acc = spmm_accumulate(col, val, start, end, 16, X, hl, hmask);
    spmm_epilogue(acc, r, hl, add, Y, acc_in, acc_out, acc_div);
  }
  __syncthreads();
#pragma unroll 1
  for (int j = 0; j < 16; ++j) {
    const int st = s_start[j], en = s_end[j];
    if (en - st <= kLongRow) continue;  // uniform across the CTA
    s_part[hw][hl] = spmm_accumulate(col, val, st + hw * 16, en, 256, X, hl, hmask);
    __syncthreads();
    if (hw == 0) {
      float4 acc = s_part[0][hl];
#pragma unroll
      for (int k = 1; k < 16; ++k) {
        const float4 p = s_part[k][hl];
        acc.x += p.x;
        acc.y += p.y;
        acc.z += p.z;
        acc.w += p.w;
      }
      spmm_epilogue(acc, r0 + j, hl, add, Y, acc_in, acc_out, acc_div);
    }
    __syncthreads();
  }
}

// ---- planned (segmented) variant ------------------------------------------------------------
constexpr int kSegNnz = 64;

// one half-warp per segment (<= 64 nonzeros: 4 groups of 16 gathers, all 16 of a group in flight)
__global__ void __launch_bounds__(256)
spmm_seg_kernel(const int32_t *__restrict__ seg_row, const int32_t *__restrict__ seg_start,
                const int32_t *__restrict__ seg_end, const int32_t *__restrict__ seg_slot, int n_seg,
                const int32_t *__restrict__ col, const float *__restrict__ val, RowSrc X,
                const uint32_t *__restrict__ x_nonzero, const uint32_t *__restrict__ row_needed,
                const float *__restrict__ add, float *Y,
                RowSrc acc_in, float *acc_out, float acc_div, float *__restrict__ partial) {
  const int lane = threadIdx.x & 31, hl = lane & 15;
  const unsigned hmask = (lane < 16) ? 0x0000ffffu : 0xffff0000u;
  const int sg = (int)(((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 4);
  if (sg >= n_seg) return;
  if (row_needed) {  // only the rows somebody reads (the batch's rows in the last forward layer)
    const int r = seg_row[sg];
    if (!((row_needed[r >> 5] >> (r & 31)) & 1u)) return;  // uniform across the half-warp
  }
  const int start = seg_start[sg], end = seg_end[sg];
  float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
  for (int base = start; base < end; base += 16) {
    const int e = base + hl;
    int c = e < end ? col[e] : -1;
    const float a = e < end ? val[e] : 0.f;
    if (x_nonzero && c >= 0 && !((x_nonzero[c >> 5] >> (c & 31)) & 1u)) c = -1;  // X row is zero
#pragma unroll
    for (int k = 0; k < 16; ++k) {
      const int cc = __shfl_sync(hmask, c, k, 16);
      const float aa = __shfl_sync(hmask, a, k, 16);
      if (cc >= 0) {  // uniform across the half-warp
        const float4 x = *row_ptr2(X, cc, hl);
        acc.x = fmaf(aa, x.x, acc.x);
        acc.y = fmaf(aa, x.y, acc.y);
        acc.z = fmaf(aa, x.z, acc.z);
        acc.w = fmaf(aa, x.w, acc.w);
      }
    }
  }
  const int slot = seg_slot[sg];
  if (slot < 0) spmm_epilogue(acc, seg_row[sg], hl, add, Y, acc_in, acc_out, acc_div);
  else reinterpret_cast<float4 *>(partial + (long long)slot * kD)[hl] = acc;
}

// rows cut into several segments: partial rows summed in a fixed order, then the epilogue.
// A row of up to kCombineWide segments is summed by one half-warp in segment order (4 partial
// rows in flight).  Longer rows -- a popular item of ml_10m has 430 segments, the head of a Zipf
// catalogue thousands -- would be one long latency chain (~0.5 us per segment, longer than the
// whole rest of the SpMM), so a whole CTA takes such a row: half-warp h sums segments h, h+16, ...
// and the 16 sums are added in half-warp order through shared memory.  Both orders depend on the
// row's segment count alone (deterministic; the same on every rank of a row-partitioned run).
constexpr int kCombineWide = 32;

__device__ __forceinline__ void add4(float4 &acc, const float4 &t) {
  acc.x += t.x;
  acc.y += t.y;
  acc.z += t.z;
  acc.w += t.w;
}

// acc += p[first], p[first + stride], ... (< n), in that order, four loads in flight
__device__ __forceinline__ void sum_partials(float4 &acc, const float4 *__restrict__ p, int first,
                                             int stride, int n) {
  int k = first;
  for (; k + 3 * stride < n; k += 4 * stride) {
    const float4 t0 = p[(long long)k * (kD / 4)];
    const float4 t1 = p[(long long)(k + stride) * (kD / 4)];
    const float4 t2 = p[(long long)(k + 2 * stride) * (kD / 4)];
    const float4 t3 = p[(long long)(k + 3 * stride) * (kD / 4)];
    add4(acc, t0);
    add4(acc, t1);
    add4(acc, t2);
    add4(acc, t3);
  }
  for (; k < n; k += stride) add4(acc, p[(long long)k * (kD / 4)]);
}

__global__ void __launch_bounds__(256)
spmm_combine_kernel(const int32_t *__restrict__ multi_row, const int32_t *__restrict__ multi_slot0,
                    const int32_t *__restrict__ multi_nseg, int n_multi,
                    const uint32_t *__restrict__ row_needed,
                    const float *__restrict__ partial, const float *__restrict__ add, float *Y,
                    RowSrc acc_in, float *acc_out, float acc_div) {
  const int hl = threadIdx.x & 15;
  const int m = (int)(((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 4);
  if (m >= n_multi) return;
  const int n = multi_nseg[m];
  if (n > kCombineWide) return;  // spmm_combine_wide_kernel's row
  if (row_needed) {
    const int r = multi_row[m];
    if (!((row_needed[r >> 5] >> (r & 31)) & 1u)) return;
  }
  const float4 *p = reinterpret_cast<const float4 *>(partial + (long long)multi_slot0[m] * kD) + hl;
  float4 acc = p[0];
  sum_partials(acc, p, 1, 1, n);
  spmm_epilogue(acc, multi_row[m], hl, add, Y, acc_in, acc_out, acc_div);
}

// one CTA per row of more than kCombineWide segments (wide_idx: indices into the multi_* arrays)
__global__ void __launch_bounds__(256)
spmm_combine_wide_kernel(const int32_t *__restrict__ wide_idx, const int32_t *__restrict__ multi_row,
                         const int32_t *__restrict__ multi_slot0, const int32_t *__restrict__ multi_nseg,
                         const uint32_t *__restrict__ row_needed,
                         const float *__restrict__ partial, const float *__restrict__ add, float *Y,
                         RowSrc acc_in, float *acc_out, float acc_div) {
  __shared__ float4 s_part[16][16];
  const int hl = threadIdx.x & 15, hw = threadIdx.x >> 4;
  const int m = wide_idx[blockIdx.x];
  if (row_needed) {
    const int r = multi_row[m];
    if (!((row_needed[r >> 5] >> (r & 31)) & 1u)) return;  // uniform across the CTA
  }
  const int n = multi_nseg[m];
  const float4 *p = reinterpret_cast<const float4 *>(partial + (long long)multi_slot0[m] * kD) + hl;
  float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
  sum_partials(acc, p, hw, 16, n);
  s_part[hw][hl] = acc;
  __syncthreads();
  if (hw == 0) {
    acc = s_part[0][hl];
#pragma unroll
    for (int k = 1; k < 16; ++k) add4(acc, s_part[k][hl]);
    spmm_epilogue(acc, multi_row[m], hl, add, Y, acc_in, acc_out, acc_div);
  }
}

// one-off, host side: the adjacency of a run never changes (LightGCN.py:257-269 builds it once)
int build_spmm_plan(const int32_t *d_rowptr, int64_t n_rows, SpmmPlan *out, const int64_t *ranges) {
  std::vector<int32_t> rp((size_t)n_rows + 1);
  MACR_CUDA(cudaMemcpy(rp.data(), d_rowptr, sizeof(int32_t) * rp.size(), cudaMemcpyDeviceToHost));
  std::vector<int32_t> srow, sstart, send, sslot, mrow, mslot0, mnseg, wide;
  int32_t n_slots = 0;
  for (int64_t r = 0; r < n_rows; ++r) {
    if (ranges && !((r >= ranges[0] && r < ranges[1]) || (r >= ranges[2] && r < ranges[3]))) continue;
    const int32_t st = rp[r], en = rp[r + 1];
    const int32_t nseg = en - st <= kSegNnz ? 1 : (en - st + kSegNnz - 1) / kSegNnz;
    if (nseg > 1) {
      if (nseg > kCombineWide) wide.push_back((int32_t)mrow.size());
      mrow.push_back((int32_t)r);
      mslot0.push_back(n_slots);
      mnseg.push_back(nseg);
    }
    for (int32_t k = 0; k < nseg; ++k) {
      srow.push_back((int32_t)r);
      sstart.push_back(st + k * kSegNnz);
      send.push_back(std::min(en, st + (k + 1) * kSegNnz));
      sslot.push_back(nseg > 1 ? n_slots++ : -1);
    }
  }
  SpmmPlan p;
  p.n_seg = (int)srow.size();
  p.n_multi = (int)mrow.size();
  p.n_wide = (int)wide.size();
  auto up = [&](const std::vector<int32_t> &v, int32_t **dst) -> int {
    MACR_CUDA(cudaMalloc(dst, sizeof(int32_t) * (v.size() + 1)));
    MACR_CUDA(cudaMemcpy(*dst, v.data(), sizeof(int32_t) * v.size(), cudaMemcpyHostToDevice));
    return MACR_OK;
  };
  int rc;
  if ((rc = up(srow, &p.seg_row)) || (rc = up(sstart, &p.seg_start)) || (rc = up(send, &p.seg_end)) ||
      (rc = up(sslot, &p.seg_slot)) || (rc = up(mrow, &p.multi_row)) ||
      (rc = up(mslot0, &p.multi_slot0)) || (rc = up(mnseg, &p.multi_nseg)) ||
      (rc = up(wide, &p.wide_idx)))
    return rc;
  MACR_CUDA(cudaMalloc(&p.partial, sizeof(float) * kD * ((size_t)n_slots + 1)));
  *out = p;
  return MACR_OK;
}

void free_spmm_plan(SpmmPlan *p) {
  cudaFree(p->seg_row), cudaFree(p->seg_start), cudaFree(p->seg_end), cudaFree(p->seg_slot);
  cudaFree(p->multi_row), cudaFree(p->multi_slot0), cudaFree(p->multi_nseg), cudaFree(p->partial);
  cudaFree(p->wide_idx);
  *p = SpmmPlan();
}

int launch_spmm_planned(const SpmmPlan *plan, const int32_t *rowptr, const int32_t *col,
                        const float *val, int64_t n_rows, RowSrc X, const float *add, float *Y,
                        RowSrc acc_in, float *acc_out, float acc_div, const uint32_t *x_nonzero,
                        cudaStream_t s, const uint32_t *row_needed) {
  if (!plan) return launch_spmm(rowptr, col, val, n_rows, X, add, Y, acc_in, acc_out, acc_div, s);
  if (plan->n_seg == 0) return MACR_OK;
  spmm_seg_kernel<<<(unsigned)(((long long)plan->n_seg * 16 + 255) / 256), 256, 0, s>>>(
      plan->seg_row, plan->seg_start, plan->seg_end, plan->seg_slot, plan->n_seg, col, val, X,
      x_nonzero, row_needed, add, Y, acc_in, acc_out, acc_div, plan->partial);
  MACR_LAUNCH_CHECK();
  if (plan->n_multi) {
    spmm_combine_kernel<<<(unsigned)(((long long)plan->n_multi * 16 + 255) / 256), 256, 0, s>>>(
        plan->multi_row, plan->multi_slot0, plan->multi_nseg, plan->n_multi, row_needed, plan->partial,
        add, Y, acc_in, acc_out, acc_div);
    MACR_LAUNCH_CHECK();
  }
  if (plan->n_wide) {
    spmm_combine_wide_kernel<<<plan->n_wide, 256, 0, s>>>(plan->wide_idx, plan->multi_row,
                                                          plan->multi_slot0, plan->multi_nseg,
                                                          row_needed, plan->partial, add, Y, acc_in,
                                                          acc_out, acc_div);
    MACR_LAUNCH_CHECK();
  }
  return MACR_OK;
}

int launch_spmm(const int32_t *rowptr, const int32_t *col, const float *val, int64_t n_rows,
                RowSrc X, const float *add, float *Y, RowSrc acc_in, float *acc_out,
                float acc_div, cudaStream_t s) {
  if (n_rows == 0) return MACR_OK;
  const long long threads = n_rows * 16;
  spmm_csr_kernel<<<(unsigned)((threads + 255) / 256), 256, 0, s>>>(rowptr, col, val, n_rows, X,
                                                                    add, Y, acc_in, acc_out,
                                                                    acc_div);
  MACR_LAUNCH_CHECK();
  return MACR_OK;
}

int launch_lgcn_propagate(const int32_t *rowptr, const int32_t *col, const float *val,
                          const float *U, int64_t n_users, const float *I, int64_t n_items,
                          int n_layers, float *Emean, float *tmp, cudaStream_t s,
                          const SpmmPlan *plan) {
  const int64_t N = n_users + n_items;
  RowSrc e0{U, I, n_users};
  if (n_layers == 0) {
    MACR_CUDA(cudaMemcpyAsync(Emean, U, sizeof(float) * n_users * kD, cudaMemcpyDeviceToDevice, s));
    MACR_CUDA(cudaMemcpyAsync(Emean + n_users * kD, I, sizeof(float) * n_items * kD,
                              cudaMemcpyDeviceToDevice, s));
    return MACR_OK;
  }
  float *buf[2] = {tmp, tmp + N * kD};
  RowSrc x = e0;
  for (int k = 0; k < n_layers; ++k) {
    const bool last = k == n_layers - 1;
    float *y = last ? nullptr : buf[k & 1];
    RowSrc accin = (k == 0) ? e0 : RowSrc{Emean, Emean, N};
    int rc = launch_spmm_planned(plan, rowptr, col, val, N, x, nullptr, y, accin, Emean,
                                 last ? (float)(n_layers + 1) : 0.f, nullptr, s);
    if (rc) return rc;
    if (!last) x = RowSrc{y, y, N};
  }
  return MACR_OK;
}

// bitmap over the N = U + I node rows of the rows the batch reads: users, pos and neg items
__global__ void __launch_bounds__(256)
mark_batch_rows_kernel(const StepState *st, int B, long long n_users, uint32_t *__restrict__ bitmap) {
  const int e = blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= 3 * B) return;
  const int32_t *ids = st->ids_base + st->step_idx * 3LL * B;
  const long long r = ids[e] + (e < B ? 0 : n_users);
  atomicOr(&bitmap[r >> 5], 1u << (r & 31));
}

int launch_mark_batch_rows(const StepState *st, int B, int64_t n_users, uint32_t *bitmap, cudaStream_t s) {
  mark_batch_rows_kernel<<<(3 * B + 255) / 256, 256, 0, s>>>(st, B, n_users, bitmap);
  MACR_LAUNCH_CHECK();
  return MACR_OK;
}

// d(Emean)/(L+1) for the unique touched rows into a zeroed dense [N][64] buffer
__global__ void __launch_bounds__(256)
scatter_rows_kernel(PlanBufs planU, const float *__restrict__ gU, PlanBufs planI,
                    const float *__restrict__ gI, int maxU, long long n_users, float div,
                    float *__restrict__ out, uint32_t *__restrict__ nz_bitmap) {
  const int lane = threadIdx.x & 31;
  const int wid = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  const bool item = wid >= maxU;
  const int slot = item ? wid - maxU : wid;
  const PlanBufs &pl = item ? planI : planU;
  if (slot >= *pl.n_uniq) return;
  const long long r = pl.uniq_rows[slot] + (item ? n_users : 0);
  float2 g = reinterpret_cast<const float2 *>((item ? gI : gU) + (long long)slot * kD)[lane];
  g.x = __fdiv_rn(g.x, div);
  g.y = __fdiv_rn(g.y, div);
  reinterpret_cast<float2 *>(out + r * kD)[lane] = g;
  if (nz_bitmap && lane == 0) atomicOr(nz_bitmap + (r >> 5), 1u << (r & 31));
}

int launch_scatter_rows(PlanBufs planU, const float *gU, PlanBufs planI, const float *gI, int B,
                        int64_t n_users, float div, float *out, uint32_t *nz_bitmap,
                        cudaStream_t s) {
  const int warps = 3 * B;
  scatter_rows_kernel<<<(warps + 7) / 8, 256, 0, s>>>(planU, gU, planI, gI, B, n_users, div, out,
                                                      nz_bitmap);
  MACR_LAUNCH_CHECK();
  return MACR_OK;
}

// L2 slices of the raw-row lookups (LightGCN.py:148-150,525-528): grad[row] += lam*raw[row]
// once per occurrence in the batch
__global__ void __launch_bounds__(256)
l2_rows_kernel(PlanBufs planU, PlanBufs planI, int maxU, const float *__restrict__ U,
               const float *__restrict__ I, long long n_users, float lam, float *__restrict__ grad) {
  const int lane = threadIdx.x & 31;
  const int wid = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  const bool item = wid >= maxU;
  const int slot = item ? wid - maxU : wid;
  const PlanBufs &pl = item ? planI : planU;
  if (slot >= *pl.n_uniq) return;
  const long long r = pl.uniq_rows[slot];
  const int cnt = pl.seg_off[slot + 1] - pl.seg_off[slot];
  const float2 raw = reinterpret_cast<const float2 *>((item ? I : U) + r * kD)[lane];
  float2 *gp = reinterpret_cast<float2 *>(grad + (r + (item ? n_users : 0)) * kD) + lane;
  float2 g = *gp;
  const float lx = __fmul_rn(lam, raw.x), ly = __fmul_rn(lam, raw.y);
  for (int k = 0; k < cnt; ++k) {
    g.x = __fadd_rn(g.x, lx);
    g.y = __fadd_rn(g.y, ly);
  }
  *gp = g;
}

int launch_l2_rows(PlanBufs planU, PlanBufs planI, int B, const float *U, const float *I,
                   int64_t n_users, float lam, float *grad, cudaStream_t s) {
  const int warps = 3 * B;
  l2_rows_kernel<<<(warps + 7) / 8, 256, 0, s>>>(planU, planI, B, U, I, n_users, lam, grad);
  MACR_LAUNCH_CHECK();
  return MACR_OK;
}

}  // namespace macr

using namespace macr;

extern "C" int macr_spmm_csr(const int32_t *rowptr, const int32_t *col, const float *val,
                             int64_t n_rows, const float *X, int d, float *Y,
                             macr_stream_t stream) {
  MACR_CHECK_ARG(d == kD, "macr_spmm_csr: d must be %d (got %d)", kD, d);
  MACR_CHECK_ARG(rowptr && col && val && X && Y, "macr_spmm_csr: null pointer");
  RowSrc x{X, X, n_rows};
  return launch_spmm(rowptr, col, val, n_rows, x, nullptr, Y, x, nullptr, 0.f, as_stream(stream));
}

struct macr_spmm_plan {
  macr::SpmmPlan plan;
  int64_t n_rows;
};

extern "C" int macr_spmm_plan_create(const int32_t *rowptr, int64_t n_rows, macr_spmm_plan **out) {
  MACR_CHECK_ARG(rowptr && out && n_rows >= 0, "macr_spmm_plan_create: bad argument");
  macr_spmm_plan *h = new (std::nothrow) macr_spmm_plan();
  MACR_CHECK_ARG(h, "macr_spmm_plan_create: out of host memory");
  h->n_rows = n_rows;
  int rc = build_spmm_plan(rowptr, n_rows, &h->plan);
  if (rc) {
    delete h;
    return rc;
  }
  *out = h;
  return MACR_OK;
}

extern "C" int macr_spmm_plan_destroy(macr_spmm_plan *h) {
  if (!h) return MACR_OK;
  free_spmm_plan(&h->plan);
  delete h;
  return MACR_OK;
}

extern "C" int macr_spmm_csr_planned(const macr_spmm_plan *plan, const int32_t *rowptr,
                                     const int32_t *col, const float *val, int64_t n_rows,
                                     const float *X, int d, float *Y, macr_stream_t stream) {
  MACR_CHECK_ARG(d == kD, "macr_spmm_csr_planned: d must be %d (got %d)", kD, d);
  MACR_CHECK_ARG(plan && rowptr && col && val && X && Y, "macr_spmm_csr_planned: null pointer");
  MACR_CHECK_ARG(plan->n_rows == n_rows, "macr_spmm_csr_planned: plan was built for %lld rows",
                 (long long)plan->n_rows);
  RowSrc x{X, X, n_rows};
  return launch_spmm_planned(&plan->plan, rowptr, col, val, n_rows, x, nullptr, Y, x, nullptr, 0.f,
                             nullptr, as_stream(stream));
}

extern "C" int macr_lgcn_propagate_planned(const macr_spmm_plan *plan, const int32_t *rowptr,
                                           const int32_t *col, const float *val, const float *U,
                                           int64_t n_users, const float *I, int64_t n_items, int d,
                                           int n_layers, float *Emean, float *tmp,
                                           macr_stream_t stream) {
  MACR_CHECK_ARG(d == kD, "macr_lgcn_propagate_planned: d must be %d (got %d)", kD, d);
  MACR_CHECK_ARG(plan && rowptr && col && val && U && I && Emean && tmp,
                 "macr_lgcn_propagate_planned: null pointer");
  MACR_CHECK_ARG(plan->n_rows == n_users + n_items,
                 "macr_lgcn_propagate_planned: plan was built for %lld rows", (long long)plan->n_rows);
  MACR_CHECK_ARG(n_layers >= 0 && n_layers <= 16, "macr_lgcn_propagate_planned: bad n_layers %d",
                 n_layers);
  return launch_lgcn_propagate(rowptr, col, val, U, n_users, I, n_items, n_layers, Emean, tmp,
                               as_stream(stream), &plan->plan);
}

extern "C" int macr_lgcn_propagate(const int32_t *rowptr, const int32_t *col, const float *val,
                                   const float *U, int64_t n_users, const float *I,
                                   int64_t n_items, int d, int n_layers, float *Emean, float *tmp,
                                   macr_stream_t stream) {
  MACR_CHECK_ARG(d == kD, "macr_lgcn_propagate: d must be %d (got %d)", kD, d);
  MACR_CHECK_ARG(rowptr && col && val && U && I && Emean && tmp, "macr_lgcn_propagate: null pointer");
  MACR_CHECK_ARG(n_layers >= 0 && n_layers <= 16, "macr_lgcn_propagate: bad n_layers %d", n_layers);
  return launch_lgcn_propagate(rowptr, col, val, U, n_users, I, n_items, n_layers, Emean, tmp,
                               as_stream(stream));
}
