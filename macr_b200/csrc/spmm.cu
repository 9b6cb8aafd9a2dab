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

namespace macr {

__device__ __forceinline__ const float4 *row_ptr2(const RowSrc &s, long long r, int hl) {
  const float *base = (r < s.split) ? s.a + r * kD : s.b + (r - s.split) * kD;
  return reinterpret_cast<const float4 *>(base) + hl;
}

__global__ void __launch_bounds__(256)
spmm_csr_kernel(const int32_t *__restrict__ rowptr, const int32_t *__restrict__ col,
                const float *__restrict__ val, long long n_rows, RowSrc X,
                const float *__restrict__ add, float *Y, RowSrc acc_in, float *acc_out,
                float acc_div) {
  const int lane = threadIdx.x & 31, hl = lane & 15;
  const unsigned hmask = (lane < 16) ? 0x0000ffffu : 0xffff0000u;
  const long long r = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 4;
  if (r >= n_rows) return;
  const int start = rowptr[r], end = rowptr[r + 1];
  float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
  for (int base = start; base < end; base += 16) {
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
                          int n_layers, float *Emean, float *tmp, cudaStream_t s) {
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
    int rc = launch_spmm(rowptr, col, val, N, x, nullptr, y, accin, Emean,
                         last ? (float)(n_layers + 1) : 0.f, s);
    if (rc) return rc;
    if (!last) x = RowSrc{y, y, N};
  }
  return MACR_OK;
}

// d(Emean)/(L+1) for the unique touched rows into a zeroed dense [N][64] buffer
__global__ void __launch_bounds__(256)
scatter_rows_kernel(PlanBufs planU, const float *__restrict__ gU, PlanBufs planI,
                    const float *__restrict__ gI, int maxU, long long n_users, float div,
                    float *__restrict__ out) {
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
}

int launch_scatter_rows(PlanBufs planU, const float *gU, PlanBufs planI, const float *gI, int B,
                        int64_t n_users, float div, float *out, cudaStream_t s) {
  const int warps = 3 * B;
  scatter_rows_kernel<<<(warps + 7) / 8, 256, 0, s>>>(planU, gU, planI, gI, B, n_users, div, out);
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
