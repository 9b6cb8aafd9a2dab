// train_kernels.cu -- sm_100a kernels of the MACR training step
//   K1+K2 gather + five dots        (macr_mf/model.py:35-37,186-187,194-196,219)
//   K3    B x B gated BCE grid      (model.py:204-217 + its autodiff)
//   K5a   batch plan (dedup)        (TF-1.14 optimizer.py _deduplicate_indexed_slices)
//   K5b   Adam, TF dense semantics  (TF-1.14 adam.py _apply_sparse_shared / ApplyAdam)
// Design notes and rooflines: DESIGN.md sections 3-4.
#include "train_kernels.cuh"

namespace macr {

// ---------------------------------------------------------------------------------------------
// small device helpers
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ float ex2_approx(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
__device__ __forceinline__ float lg2_approx(float x) {
  float y;
  asm("lg2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
__device__ __forceinline__ float rcp_approx(float x) {
  float y;
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
// gates are only O(B) per step: full-precision expf and IEEE division
__device__ __forceinline__ float sigmoid_precise(float x) {
  return __fdiv_rn(1.0f, __fadd_rn(1.0f, expf(-x)));
}
__device__ __forceinline__ float step_lr_t(const StepState *st, float lr_or_lrt) {
  if (st == nullptr) return lr_or_lrt;
  // adam.py: lr * sqrt(1 - beta2_power) / (1 - beta1_power), fp32, left to right
  return __fdiv_rn(__fmul_rn(lr_or_lrt, __fsqrt_rn(__fsub_rn(1.0f, st->b2p))),
                   __fsub_rn(1.0f, st->b1p));
}
__device__ __forceinline__ const int32_t *step_ids(const StepState *st, const int32_t *direct,
                                                   int B, int which) {
  if (st == nullptr) return direct;
  return st->ids_base + st->step_idx * 3LL * B + (long long)which * B;
}

// ---------------------------------------------------------------------------------------------
// K1+K2: one warp per triple; lane l owns elements [2l, 2l+1] of every 64-wide row
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
gather_dots_kernel(const float *__restrict__ Ue, const float *__restrict__ Ie,
                   const float *__restrict__ Ur, const float *__restrict__ Ir,
                   const float *__restrict__ w, const float *__restrict__ wu,
                   const int32_t *u_, const int32_t *p_, const int32_t *n_, const StepState *st,
                   int B, float *__restrict__ yp, float *__restrict__ yn, float *__restrict__ sp,
                   float *__restrict__ sn, float *__restrict__ su, float *__restrict__ regsq) {
  const int lane = threadIdx.x & 31;
  const int b = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (b >= B) return;
  const int32_t *u = step_ids(st, u_, B, 0), *p = step_ids(st, p_, B, 1),
                *n = step_ids(st, n_, B, 2);
  const long long ur = u[b], pr = p[b], nr = n[b];
  const float2 ue = reinterpret_cast<const float2 *>(Ue + ur * kD)[lane];
  const float2 pe = reinterpret_cast<const float2 *>(Ie + pr * kD)[lane];
  const float2 ne = reinterpret_cast<const float2 *>(Ie + nr * kD)[lane];
  const float2 wv = reinterpret_cast<const float2 *>(w)[lane];
  const float2 wuv = reinterpret_cast<const float2 *>(wu)[lane];
  float a0 = ue.x * pe.x + ue.y * pe.y;
  float a1 = ue.x * ne.x + ue.y * ne.y;
  float a2 = pe.x * wv.x + pe.y * wv.y;
  float a3 = ne.x * wv.x + ne.y * wv.y;
  float a4 = ue.x * wuv.x + ue.y * wuv.y;
  float a5;
  if (Ur == Ue && Ir == Ie) {
    a5 = ue.x * ue.x + ue.y * ue.y + pe.x * pe.x + pe.y * pe.y + ne.x * ne.x + ne.y * ne.y;
  } else {  // LightGCN: L2 term on the raw rows (LightGCN.py:148-150,525-526)
    const float2 u0 = reinterpret_cast<const float2 *>(Ur + ur * kD)[lane];
    const float2 p0 = reinterpret_cast<const float2 *>(Ir + pr * kD)[lane];
    const float2 n0 = reinterpret_cast<const float2 *>(Ir + nr * kD)[lane];
    a5 = u0.x * u0.x + u0.y * u0.y + p0.x * p0.x + p0.y * p0.y + n0.x * n0.x + n0.y * n0.y;
  }
  a0 = warp_sum(a0);
  a1 = warp_sum(a1);
  a2 = warp_sum(a2);
  a3 = warp_sum(a3);
  a4 = warp_sum(a4);
  a5 = warp_sum(a5);
  if (lane == 0) {
    yp[b] = a0;
    yn[b] = a1;
    sp[b] = a2;
    sn[b] = a3;
    su[b] = a4;
    regsq[b] = a5;
  }
}

int launch_gather_dots(const float *Ue, const float *Ie, const float *Ur, const float *Ir,
                       const float *w, const float *wu, const int32_t *u, const int32_t *p,
                       const int32_t *n, const StepState *st, int B, float *yp, float *yn,
                       float *sp, float *sn, float *su, float *regsq, cudaStream_t s) {
  const int wpb = 8;
  gather_dots_kernel<<<(B + wpb - 1) / wpb, wpb * 32, 0, s>>>(Ue, Ie, Ur, Ir, w, wu, u, p, n, st,
                                                              B, yp, yn, sp, sn, su, regsq);
  MACR_LAUNCH_CHECK();
  return MACR_OK;
}

// ---------------------------------------------------------------------------------------------
// K3: the B x B grid.  CTA tile = (16*RI) rows x (16*RJ) columns, 256 threads as 16(tx) x 16(ty),
// each thread owns an RI x RJ register micro-tile (rows ty+16r, columns tx+16c): the whole inner
// loop runs on registers; row / column partial sums meet in shared memory once per tile.
// The [B,B] matrices are never written anywhere.
//
// Per (i,j) pair the reference evaluates (model.py:204-211)
//   P=(yp_j*a_i)*g_i  s=sig(P)  lossP=-log(s+1e-10)      dlossP/dP = -s(1-s)/(s+1e-10)
//   N=(yn_j*an_i)*g_i t=sig(N)  lossN=-log((1-t)+1e-10)  dlossN/dN =  t(1-t)/((1-t)+1e-10)
// Fast path (taken when s >= 2^-9 and 1-t >= 2^-9, where fp32 "+1e-10" is a no-op exactly as in
// the reference's own fp32 arithmetic): one rcp serves both sigmoids, one lg2 serves both logs,
// and the gradients collapse to s-1 and t  ->  4 MUFU ops per pair.  Otherwise the literal
// formulas are evaluated (8 MUFU ops).
// ---------------------------------------------------------------------------------------------
template <int RI, int RJ, bool kMasked, bool kGrad>
__global__ void __launch_bounds__(256, (RI * RJ >= 64) ? 2 : 3)
grid_bce_kernel(const float *__restrict__ yp, const float *__restrict__ yn,
                const float *__restrict__ sp, const float *__restrict__ sn,
                const float *__restrict__ su, int B, int Bpad, float *__restrict__ rowP_part,
                float *__restrict__ rowN_part, float *__restrict__ colP_part,
                float *__restrict__ colN_part, float *__restrict__ losspart) {
  constexpr int TI = 16 * RI, TJ = 16 * RJ;
  constexpr int TMAX = TI > TJ ? TI : TJ;
  constexpr float kLog2e = 1.4426950408889634f;
  __shared__ float sA[TI], sAN[TI], sG[TI];
  __shared__ float sYp[TJ], sYn[TJ];
  __shared__ float sRed[2][TMAX][17];
  __shared__ float sLoss[8];

  const int tid = threadIdx.x, tx = tid & 15, ty = tid >> 4;
  const int i0 = blockIdx.y * TI, j0 = blockIdx.x * TJ;

  for (int t = tid; t < TI; t += 256) {
    const int i = i0 + t;
    const bool ok = i < B;
    sA[t] = ok ? sigmoid_precise(sp[i]) : 0.f;
    sAN[t] = ok ? sigmoid_precise(sn[i]) : 0.f;
    sG[t] = ok ? sigmoid_precise(su[i]) : 0.f;
  }
  for (int t = tid; t < TJ; t += 256) {
    const int j = j0 + t;
    const bool ok = j < B;
    sYp[t] = ok ? yp[j] : 0.f;
    sYn[t] = ok ? yn[j] : 0.f;
  }
  __syncthreads();

  float a[RI], an[RI], gl[RI], ag[RI], ang[RI], wr[RI];
  float ypj[RJ], ynj[RJ], wc[RJ];
#pragma unroll
  for (int r = 0; r < RI; ++r) {
    const int t = ty + 16 * r;
    a[r] = sA[t];
    an[r] = sAN[t];
    const float g = sG[t];
    gl[r] = -g * kLog2e;  // exp(-P) = ex2((yp*a)*(-g*log2e))
    ag[r] = a[r] * g;
    ang[r] = an[r] * g;
    wr[r] = (i0 + t < B) ? 1.f : 0.f;
  }
#pragma unroll
  for (int c = 0; c < RJ; ++c) {
    const int t = tx + 16 * c;
    ypj[c] = sYp[t];
    ynj[c] = sYn[t];
    wc[c] = (j0 + t < B) ? 1.f : 0.f;
  }

  float colP[RJ], colN[RJ], rowP[RI], rowN[RI];
#pragma unroll
  for (int c = 0; c < RJ; ++c) colP[c] = colN[c] = 0.f;
#pragma unroll
  for (int r = 0; r < RI; ++r) rowP[r] = rowN[r] = 0.f;
  float lgacc = 0.f;  // sum of log2(.) terms; loss = -ln2 * sum

#pragma unroll
  for (int r = 0; r < RI; ++r) {
#pragma unroll
    for (int c = 0; c < RJ; ++c) {
      const float eP = ex2_approx((ypj[c] * a[r]) * gl[r]);   // exp(-P)
      const float eN = ex2_approx((ynj[c] * an[r]) * gl[r]);  // exp(-N)
      const float DP = 1.0f + eP, DN = 1.0f + eN;
      float lg, dP, dN;
      if (eP <= 500.0f && eN >= 0.00390625f && eN <= 1.0e18f) {
        const float rr = rcp_approx(DP * DN);
        const float s = rr * DN, t = rr * DP;
        const float q = 1.0f - t;
        lg = lg2_approx(s * q);
        dP = s - 1.0f;
        dN = t;
      } else {
        const float s = rcp_approx(DP), t = rcp_approx(DN);
        const float se = s + kBceEps, q = (1.0f - t) + kBceEps;
        lg = lg2_approx(se) + lg2_approx(q);
        dP = -(s * (1.0f - s)) * rcp_approx(se);
        dN = (t * (1.0f - t)) * rcp_approx(q);
      }
      if (kMasked) {
        const float mk = wr[r] * wc[c];
        lg *= mk;
        dP *= mk;
        dN *= mk;
      }
      lgacc += lg;
      if (kGrad) {
        colP[c] = fmaf(dP, ag[r], colP[c]);
        colN[c] = fmaf(dN, ang[r], colN[c]);
        rowP[r] = fmaf(dP, ypj[c], rowP[r]);
        rowN[r] = fmaf(dN, ynj[c], rowN[r]);
      }
    }
  }

  // ---- per-tile reductions -------------------------------------------------------------
  if (kGrad) {
#pragma unroll
    for (int r = 0; r < RI; ++r) {
      sRed[0][ty + 16 * r][tx] = rowP[r];
      sRed[1][ty + 16 * r][tx] = rowN[r];
    }
    __syncthreads();
    for (int t = tid; t < 2 * TI; t += 256) {
      const int which = t / TI, row = t - which * TI;
      float acc = 0.f;
#pragma unroll
      for (int k = 0; k < 16; ++k) acc += sRed[which][row][k];
      float *dst = which ? rowN_part : rowP_part;
      dst[(size_t)blockIdx.x * Bpad + i0 + row] = acc;
    }
    __syncthreads();
#pragma unroll
    for (int c = 0; c < RJ; ++c) {
      sRed[0][tx + 16 * c][ty] = colP[c];
      sRed[1][tx + 16 * c][ty] = colN[c];
    }
    __syncthreads();
    for (int t = tid; t < 2 * TJ; t += 256) {
      const int which = t / TJ, colm = t - which * TJ;
      float acc = 0.f;
#pragma unroll
      for (int k = 0; k < 16; ++k) acc += sRed[which][colm][k];
      float *dst = which ? colN_part : colP_part;
      dst[(size_t)blockIdx.y * Bpad + j0 + colm] = acc;
    }
  }
  lgacc = warp_sum(lgacc);
  if ((tid & 31) == 0) sLoss[tid >> 5] = lgacc;
  __syncthreads();
  if (tid == 0) {
    float acc = 0.f;
#pragma unroll
    for (int k = 0; k < 8; ++k) acc += sLoss[k];
    losspart[blockIdx.y * gridDim.x + blockIdx.x] = acc;
  }
}

// per-b epilogue of the grid: fold the tile partials, finish the chain rule through the three
// sigmoids, add the alpha / beta branch gradients (model.py:213-217).
// 256 threads = 64 batch positions x 4 partial arrays, so the nblk-long folds run 4-wide and
// coalesced; thread (b, 0) then finishes the scalar math.
__global__ void __launch_bounds__(256)
grid_finalize_kernel(const float *__restrict__ sp, const float *__restrict__ sn,
                     const float *__restrict__ su, int B, int Bpad, int nblk, float alpha,
                     float beta, const float *__restrict__ rowP_part,
                     const float *__restrict__ rowN_part, const float *__restrict__ colP_part,
                     const float *__restrict__ colN_part, float *__restrict__ d_yp,
                     float *__restrict__ d_yn, float *__restrict__ d_sp, float *__restrict__ d_sn,
                     float *__restrict__ d_su, float *__restrict__ litem,
                     float *__restrict__ luser, int want_grad) {
  __shared__ float sh[4][64];
  const int bl = threadIdx.x & 63, which = threadIdx.x >> 6;
  const int b = blockIdx.x * 64 + bl;
  if (want_grad) {
    const float *src = which == 0 ? rowP_part : which == 1 ? rowN_part : which == 2 ? colP_part
                                                                                     : colN_part;
    float acc = 0.f;
    if (b < B) {
      int k = 0;
      for (; k + 8 <= nblk; k += 8) {
        float v[8];
#pragma unroll
        for (int q = 0; q < 8; ++q) v[q] = src[(size_t)(k + q) * Bpad + b];
#pragma unroll
        for (int q = 0; q < 8; ++q) acc += v[q];
      }
      for (; k < nblk; ++k) acc += src[(size_t)k * Bpad + b];
    }
    sh[which][bl] = acc;
    __syncthreads();
  }
  if (which != 0 || b >= B) return;
  const float a = sigmoid_precise(sp[b]), an = sigmoid_precise(sn[b]),
              g = sigmoid_precise(su[b]);
  const float ea = a + kBceEps, ean = (1.0f - an) + kBceEps;
  const float eg = g + kBceEps, eg1 = (1.0f - g) + kBceEps;
  litem[b] = -logf(ea) - logf(ean);
  luser[b] = -logf(eg) - logf(eg1);
  if (!want_grad) return;
  const float rp = sh[0][bl], rn = sh[1][bl], cp = sh[2][bl], cn = sh[3][bl];
  const float invB = 1.0f / (float)B;
  const float invBB = invB * invB;
  d_yp[b] = cp * invBB;
  d_yn[b] = cn * invBB;
  const float da = rp * invBB * g - alpha * invB / ea;
  const float dan = rn * invBB * g + alpha * invB / ean;
  const float dg = (rp * a + rn * an) * invBB + beta * invB * (1.0f / eg1 - 1.0f / eg);
  d_sp[b] = da * (a * (1.0f - a));
  d_sn[b] = dan * (an * (1.0f - an));
  d_su[b] = dg * (g * (1.0f - g));
}

GridWs grid_ws_layout(int B, void *base) {
  GridWs w;
  w.tile = (B >= 2048) ? 128 : 64;
  w.nblk = (B + w.tile - 1) / w.tile;
  w.Bpad = w.nblk * w.tile;
  float *p = reinterpret_cast<float *>(base);
  const size_t band = (size_t)w.nblk * w.Bpad;
  w.rowP = p;
  w.rowN = p + band;
  w.colP = p + 2 * band;
  w.colN = p + 3 * band;
  w.losspart = p + 4 * band;
  const size_t lp = ((size_t)w.nblk * w.nblk + 3) & ~(size_t)3;
  w.litem = w.losspart + lp;
  w.luser = w.litem + w.Bpad;
  w.bytes = (4 * band + lp + 2 * (size_t)w.Bpad) * sizeof(float);
  return w;
}

template <int R, bool kGrad>
static void launch_grid_t(const float *yp, const float *yn, const float *sp, const float *sn,
                          const float *su, int B, const GridWs &ws, cudaStream_t s) {
  dim3 grid(ws.nblk, ws.nblk);
  if (B % ws.tile == 0)
    grid_bce_kernel<R, R, false, kGrad><<<grid, 256, 0, s>>>(
        yp, yn, sp, sn, su, B, ws.Bpad, ws.rowP, ws.rowN, ws.colP, ws.colN, ws.losspart);
  else
    grid_bce_kernel<R, R, true, kGrad><<<grid, 256, 0, s>>>(
        yp, yn, sp, sn, su, B, ws.Bpad, ws.rowP, ws.rowN, ws.colP, ws.colN, ws.losspart);
}

int launch_grid_bce(const float *yp, const float *yn, const float *sp, const float *sn,
                    const float *su, int B, float alpha, float beta, const GridWs &ws,
                    float *d_yp, float *d_yn, float *d_sp, float *d_sn, float *d_su,
                    int want_grad, cudaStream_t s) {
  if (ws.tile == 128) {
    if (want_grad) launch_grid_t<8, true>(yp, yn, sp, sn, su, B, ws, s);
    else launch_grid_t<8, false>(yp, yn, sp, sn, su, B, ws, s);
  } else {
    if (want_grad) launch_grid_t<4, true>(yp, yn, sp, sn, su, B, ws, s);
    else launch_grid_t<4, false>(yp, yn, sp, sn, su, B, ws, s);
  }
  MACR_LAUNCH_CHECK();
  grid_finalize_kernel<<<(B + 63) / 64, 256, 0, s>>>(sp, sn, su, B, ws.Bpad, ws.nblk, alpha,
                                                       beta, ws.rowP, ws.rowN, ws.colP, ws.colN,
                                                       d_yp, d_yn, d_sp, d_sn, d_su, ws.litem,
                                                       ws.luser, want_grad);
  MACR_LAUNCH_CHECK();
  return MACR_OK;
}

// deterministic block sum in double (fixed tree); result valid in thread 0
__device__ double block_sum_1024(double v, double *sh) {
  const int tid = threadIdx.x;
  sh[tid] = v;
  __syncthreads();
  for (int o = 512; o > 0; o >>= 1) {
    if (tid < o) sh[tid] += sh[tid + o];
    __syncthreads();
  }
  const double r = sh[0];
  __syncthreads();
  return r;
}

__global__ void __launch_bounds__(1024)
reduce_losses_kernel(const float *__restrict__ losspart, int nparts, const float *__restrict__ litem,
                     const float *__restrict__ luser, const float *__restrict__ regsq, int B,
                     float alpha, float beta, float decay, int batch_size_flag,
                     float *__restrict__ losses3, const StepState *st) {
  __shared__ double sh[1024];
  const int tid = threadIdx.x;
  double a0 = 0, a1 = 0, a2 = 0, a3 = 0;
  for (int k = tid; k < nparts; k += 1024) a0 += losspart[k];
  for (int k = tid; k < B; k += 1024) {
    a1 += litem[k];
    a2 += luser[k];
    if (regsq) a3 += regsq[k];
  }
  a0 = block_sum_1024(a0, sh);
  a1 = block_sum_1024(a1, sh);
  a2 = block_sum_1024(a2, sh);
  a3 = block_sum_1024(a3, sh);
  if (tid == 0) {
    const double invB = 1.0 / (double)B;
    const float l_ori = (float)(-0.6931471805599453 * a0 * invB * invB);
    const float l_item = (float)(a1 * invB), l_user = (float)(a2 * invB);
    if (st == nullptr) {
      losses3[0] = l_ori;
      losses3[1] = l_item;
      losses3[2] = l_user;
    } else {
      const float reg = decay * ((float)(a3 * 0.5) / (float)batch_size_flag);
      const float mf = l_ori + alpha * l_item + beta * l_user;
      float *out = st->loss_base + st->step_idx * 4;
      out[0] = mf + reg;
      out[1] = mf;
      out[2] = reg;
      out[3] = l_ori;
    }
  }
}

int launch_reduce_losses(const GridWs &ws, const float *regsq, int B, float alpha, float beta,
                         float decay, int batch_size_flag, float *losses3, const StepState *st,
                         cudaStream_t s) {
  reduce_losses_kernel<<<1, 1024, 0, s>>>(ws.losspart, ws.nblk * ws.nblk, ws.litem, ws.luser,
                                          regsq, B, alpha, beta, decay, batch_size_flag, losses3,
                                          st);
  MACR_LAUNCH_CHECK();
  return MACR_OK;
}

// ---------------------------------------------------------------------------------------------
// K5a: batch plan -- group the batch positions that hit the same table row, deterministically.
// All-pairs counting instead of a sort (n <= 16384 ids, so n^2 equality tests spread over the
// whole GPU cost a few microseconds and need no table-sized scratch):
//   phase 1 (every CTA): 32 positions x 8 sub-lanes; the full id array is staged in shared
//     memory and each position learns  rank  = #earlier positions with the same row,
//     total = #positions with the same row,  lead = first position with the same row.
//   phase 2 (last CTA to finish, per table): positions with rank 0 are segment leaders; one block
//     scan over the positions gives their slot (first-occurrence order, the order
//     array_ops.unique produces) and the segment offsets; every position then drops itself at
//     seg_off[slot] + rank, i.e. ascending position inside a segment -- the order TF's
//     unsorted_segment_sum adds in.
// ---------------------------------------------------------------------------------------------
struct PlanTable {
  const int32_t *ids;
  int ids_off, n_ids, blocks;
  PlanBufs out;
  uint32_t *bitmap;
  int32_t *total, *lead;
  unsigned *counter;
};

constexpr int kPlanPos = 128;  // positions per CTA in phase 1 (x 8 sub-lanes = 1024 threads)

__global__ void __launch_bounds__(1024)
batch_plan_kernel(PlanTable t0, PlanTable t1, const StepState *st, int B) {
  extern __shared__ __align__(16) int32_t sids[];
  const bool second = (int)blockIdx.x >= t0.blocks;
  const PlanTable &t = second ? t1 : t0;
  const int blk = second ? blockIdx.x - t0.blocks : blockIdx.x;
  const int tid = threadIdx.x, n = t.n_ids;
  const int32_t *ids = (st ? st->ids_base + st->step_idx * 3LL * B : t.ids) + t.ids_off;
  const int n4 = (n + 3) >> 2;
  if ((n & 3) == 0 && (reinterpret_cast<uintptr_t>(ids) & 15) == 0) {
    const int4 *g4 = reinterpret_cast<const int4 *>(ids);
    int4 *d4 = reinterpret_cast<int4 *>(sids);
    for (int e = tid; e < n4; e += 1024) d4[e] = g4[e];
  } else {
    for (int e = tid; e < n4 * 4; e += 1024) sids[e] = e < n ? ids[e] : -1;
  }
  __syncthreads();

  const int q = blk * kPlanPos + (tid >> 3), sub = tid & 7;
  const int my = q < n ? sids[q] : -2;
  int rank = 0, total = 0, lead = 0x7fffffff;
  const int4 *s4 = reinterpret_cast<const int4 *>(sids);
#pragma unroll 4
  for (int i = sub; i < n4; i += 8) {
    const int4 x = s4[i];
    if (x.x == my || x.y == my || x.z == my || x.w == my) {  // rare: matches are sparse
      const int j = 4 * i;
      const int e0 = x.x == my, e1 = x.y == my, e2 = x.z == my, e3 = x.w == my;
      total += e0 + e1 + e2 + e3;
      rank += (e0 & (j < q)) + (e1 & (j + 1 < q)) + (e2 & (j + 2 < q)) + (e3 & (j + 3 < q));
      const int f = e0 ? j : e1 ? j + 1 : e2 ? j + 2 : j + 3;
      lead = min(lead, f);
    }
  }
#pragma unroll
  for (int o = 1; o < 8; o <<= 1) {
    rank += __shfl_xor_sync(0xffffffffu, rank, o);
    total += __shfl_xor_sync(0xffffffffu, total, o);
    lead = min(lead, __shfl_xor_sync(0xffffffffu, lead, o));
  }
  if (sub == 0 && q < n) {
    t.out.rank[q] = rank;
    t.total[q] = total;
    t.lead[q] = lead;
    if (t.bitmap) atomicOr(&t.bitmap[(uint32_t)my >> 5], 1u << (my & 31));
  }
}

// phase 2: one CTA per table.  rank / total / slot / seg_off are staged in shared memory (16-bit:
// n <= 16384) so the sequential passes every thread makes over its consecutive positions never
// wait on global memory; everything that touches global memory is coalesced and independent.
__global__ void __launch_bounds__(1024)
batch_plan_tail_kernel(PlanTable t0, PlanTable t1, const StepState *st, int B) {
  extern __shared__ __align__(16) unsigned short sh16[];
  __shared__ int s_scan[2][32];
  const PlanTable &t = blockIdx.x == 0 ? t0 : t1;
  const int tid = threadIdx.x, n = t.n_ids;
  if (n == 0) return;
  const int32_t *ids = (st ? st->ids_base + st->step_idx * 3LL * B : t.ids) + t.ids_off;
  const int npad = ((n + 1023) / 1024) * 1024;
  unsigned short *rankS = sh16, *totalS = sh16 + npad, *slotS = sh16 + 2 * npad,
                 *offS = sh16 + 3 * npad;
  constexpr int kMaxPer = 16;
  int leadv[kMaxPer], idv[kMaxPer];
#pragma unroll
  for (int k = 0; k < kMaxPer; ++k) {
    const int x = tid + k * 1024;
    if (x < npad) {
      rankS[x] = x < n ? (unsigned short)t.out.rank[x] : (unsigned short)1;
      totalS[x] = x < n ? (unsigned short)t.total[x] : (unsigned short)0;
      leadv[k] = x < n ? t.lead[x] : 0;
      idv[k] = x < n ? ids[x] : 0;
    }
  }
  __syncthreads();
  const int per = npad >> 10;
  const int q0 = tid * per, q1 = q0 + per;
  int cntL = 0, sumT = 0;
  for (int x = q0; x < q1; ++x)
    if (rankS[x] == 0) {
      cntL += 1;
      sumT += totalS[x];
    }
  int iL = cntL, iT = sumT;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    const int a = __shfl_up_sync(0xffffffffu, iL, o), b = __shfl_up_sync(0xffffffffu, iT, o);
    if ((tid & 31) >= o) {
      iL += a;
      iT += b;
    }
  }
  if ((tid & 31) == 31) {
    s_scan[0][tid >> 5] = iL;
    s_scan[1][tid >> 5] = iT;
  }
  __syncthreads();
  int baseL = 0, baseT = 0, allL = 0;
#pragma unroll
  for (int w = 0; w < 32; ++w) {
    const int a = s_scan[0][w], b = s_scan[1][w];
    if (w < (tid >> 5)) {
      baseL += a;
      baseT += b;
    }
    allL += a;
  }
  int slot = baseL + iL - cntL, off = baseT + iT - sumT;
  for (int x = q0; x < q1; ++x)
    if (rankS[x] == 0) {
      slotS[x] = (unsigned short)slot;
      offS[slot] = (unsigned short)off;
      off += totalS[x];
      ++slot;
    }
  __syncthreads();
#pragma unroll
  for (int k = 0; k < kMaxPer; ++k) {
    const int x = tid + k * 1024;
    if (x < n) {
      const int sl = slotS[leadv[k]];
      const int rk = rankS[x];
      t.out.pslot[x] = sl;
      t.out.seg_pos[(int)offS[sl] + rk] = x;
      if (rk == 0) t.out.uniq_rows[sl] = idv[k];
    }
    if (x < allL) {
      t.out.seg_off[x] = offS[x];
      t.out.done[x] = 0;
    }
  }
  if (tid == 0) {
    t.out.seg_off[allL] = n;
    *t.out.n_uniq = allL;
  }
}

// touched-row bitmaps of both tables straight from the ids (lets the dense sweep start without
// waiting for the plan)
__global__ void __launch_bounds__(256)
mark_touched_kernel(const StepState *st, const int32_t *ids_direct, int B, uint32_t *bmU,
                    uint32_t *bmI) {
  const int e = blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= 3 * B) return;
  const int32_t *ids = st ? st->ids_base + st->step_idx * 3LL * B : ids_direct;
  const uint32_t r = (uint32_t)ids[e];
  atomicOr(&(e < B ? bmU : bmI)[r >> 5], 1u << (r & 31));
}

int launch_mark_touched(const StepState *st, const int32_t *ids, int B, uint32_t *bmU,
                        uint32_t *bmI, cudaStream_t s) {
  mark_touched_kernel<<<(3 * B + 255) / 256, 256, 0, s>>>(st, ids, B, bmU, bmI);
  MACR_LAUNCH_CHECK();
  return MACR_OK;
}

// scratch layout behind PlanBufs (rank, pslot, done) and PlanTable (total, lead, counter)
size_t plan_ws_bytes(int n_ids) { return sizeof(int32_t) * (5 * (size_t)n_ids + 16); }

int plan_init() {  // opt in to 64 KB dynamic shared memory once (outside any stream capture)
  static bool attr_set = false;
  if (!attr_set) {
    MACR_CUDA(cudaFuncSetAttribute(batch_plan_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                   16384 * 4));
    MACR_CUDA(cudaFuncSetAttribute(batch_plan_tail_kernel,
                                   cudaFuncAttributeMaxDynamicSharedMemorySize, 16384 * 8));
    attr_set = true;
  }
  return MACR_OK;
}

PlanBufs plan_carve(int32_t *uniq_rows, int32_t *seg_off, int32_t *seg_pos, int32_t *n_uniq,
                    void *ws, int n_ids) {
  int32_t *p = reinterpret_cast<int32_t *>(ws);
  PlanBufs b;
  b.uniq_rows = uniq_rows;
  b.seg_off = seg_off;
  b.seg_pos = seg_pos;
  b.n_uniq = n_uniq;
  b.rank = p;
  b.pslot = p + n_ids;
  b.done = p + 2 * (size_t)n_ids;
  b.total = p + 3 * (size_t)n_ids;
  b.lead = p + 4 * (size_t)n_ids;
  b.counter = reinterpret_cast<unsigned *>(p + 5 * (size_t)n_ids);
  return b;
}

int launch_batch_plan2(const int32_t *ids0, const StepState *st, int ids0_off, int n_ids0,
                       PlanBufs out0, uint32_t *bitmap0, const int32_t *ids1, int ids1_off,
                       int n_ids1, PlanBufs out1, uint32_t *bitmap1, cudaStream_t s) {
  const int nmax = n_ids0 > n_ids1 ? n_ids0 : n_ids1;
  MACR_CHECK_ARG(nmax <= 16384, "batch plan supports at most 16384 ids per table (batch <= 8192)");
  int rci = plan_init();
  if (rci) return rci;
  PlanTable t0{ids0, ids0_off, n_ids0, (n_ids0 + kPlanPos - 1) / kPlanPos, out0, bitmap0, out0.total, out0.lead,
               out0.counter};
  PlanTable t1{ids1, ids1_off, n_ids1, (n_ids1 + kPlanPos - 1) / kPlanPos, out1, bitmap1, out1.total, out1.lead,
               out1.counter};
  const size_t smem = (size_t)((nmax + 3) / 4) * 16;
  const int B = st ? n_ids0 : 0;  // trainer convention: table 0 = users (B ids)
  batch_plan_kernel<<<t0.blocks + t1.blocks, 1024, smem, s>>>(t0, t1, st, B);
  MACR_LAUNCH_CHECK();
  const size_t smem2 = (size_t)(((nmax + 1023) / 1024) * 1024) * 8;
  batch_plan_tail_kernel<<<n_ids1 > 0 ? 2 : 1, 1024, smem2, s>>>(t0, t1, st, B);
  MACR_LAUNCH_CHECK();
  return MACR_OK;
}

// ---------------------------------------------------------------------------------------------
// K5b: Adam.  The sweep is the HBM-bound part of the step: 24 B per table element
// (read + write of var, m, v), 128-bit streaming accesses, 4 independent float4 triples in
// flight per thread.  Rows touched by the batch are skipped here (bitmap) and handled by
// adam_rows_kernel once their gradient is known, so the sweep can overlap the B x B grid.
// ---------------------------------------------------------------------------------------------
struct SweepTable {
  float4 *var, *m, *v;
  long long n4;  // rows * 16
  const uint32_t *bitmap;
};

template <int UNROLL>
__global__ void __launch_bounds__(256)
adam_sweep_kernel(SweepTable t0, SweepTable t1, float lr_or_lrt, const StepState *st, float b1,
                  float b2, float eps) {
  const float lr_t = step_lr_t(st, lr_or_lrt);
  const long long total = t0.n4 + t1.n4;
  const long long stride = (long long)gridDim.x * blockDim.x * UNROLL;
  for (long long base = (long long)blockIdx.x * blockDim.x * UNROLL + threadIdx.x; base < total;
       base += stride) {
    float4 x[UNROLL], mm[UNROLL], vv[UNROLL];
    float4 *px[UNROLL], *pm[UNROLL], *pv[UNROLL];
    bool live[UNROLL];
#pragma unroll
    for (int k = 0; k < UNROLL; ++k) {
      long long e = base + (long long)k * blockDim.x;
      live[k] = e < total;
      if (live[k]) {
        const bool second = e >= t0.n4;
        const SweepTable &t = second ? t1 : t0;
        if (second) e -= t0.n4;
        const long long row = e >> 4;
        if (t.bitmap && ((t.bitmap[row >> 5] >> (row & 31)) & 1u)) live[k] = false;
        px[k] = t.var + e;
        pm[k] = t.m + e;
        pv[k] = t.v + e;
      }
    }
#pragma unroll
    for (int k = 0; k < UNROLL; ++k)
      if (live[k]) {
        x[k] = ld_stream(px[k]);
        mm[k] = ld_stream(pm[k]);
        vv[k] = ld_stream(pv[k]);
      }
#pragma unroll
    for (int k = 0; k < UNROLL; ++k)
      if (live[k]) {
        // rows never touched so far have m = v = 0: the update is the identity, bit for bit
        // (m*b1 = 0, v*b2 = 0, var - (lr_t*0)/(0+eps) = var) -> nothing to compute or write
        const bool zero = mm[k].x == 0.f && mm[k].y == 0.f && mm[k].z == 0.f && mm[k].w == 0.f &&
                          vv[k].x == 0.f && vv[k].y == 0.f && vv[k].z == 0.f && vv[k].w == 0.f;
        if (zero) continue;
        adam_decay_only(x[k].x, mm[k].x, vv[k].x, lr_t, b1, b2, eps);
        adam_decay_only(x[k].y, mm[k].y, vv[k].y, lr_t, b1, b2, eps);
        adam_decay_only(x[k].z, mm[k].z, vv[k].z, lr_t, b1, b2, eps);
        adam_decay_only(x[k].w, mm[k].w, vv[k].w, lr_t, b1, b2, eps);
        st_stream(px[k], x[k]);
        st_stream(pm[k], mm[k]);
        st_stream(pv[k], vv[k]);
      }
  }
}

int launch_adam_sweep2(float *var0, float *m0, float *v0, int64_t rows0, const uint32_t *bm0,
                       float *var1, float *m1, float *v1, int64_t rows1, const uint32_t *bm1,
                       float lr_t, const StepState *st, float b1, float b2, float eps,
                       cudaStream_t s) {
  SweepTable t0{(float4 *)var0, (float4 *)m0, (float4 *)v0, rows0 * (kD / 4), bm0};
  SweepTable t1{(float4 *)var1, (float4 *)m1, (float4 *)v1, rows1 * (kD / 4), bm1};
  const long long total = t0.n4 + t1.n4;
  if (total == 0) return MACR_OK;
  constexpr int UNROLL = 4;
  long long blocks = (total + 256LL * UNROLL - 1) / (256LL * UNROLL);
  const long long cap = (long long)sm_count() * 8;
  if (blocks > cap) blocks = cap;
  adam_sweep_kernel<UNROLL><<<(int)blocks, 256, 0, s>>>(t0, t1, lr_t, st, b1, b2, eps);
  MACR_LAUNCH_CHECK();
  return MACR_OK;
}

// ---------------------------------------------------------------------------------------------
// row gradients of the unique touched rows, summed per table row (the IndexedSlices dedup).
//   user row r :  sum_b  dyp_b*Ie[p_b] + dyn_b*Ie[n_b] + dsu_b*w_user (+ lam*Ur[r])
//   item row r :  sum_q  dy_q*Ue[u_b] + ds_q*w (+ lam*Ir[r]),  q<B: pos role, q>=B: neg role
// Popular items own segments hundreds of positions long, so the unit of work is <= 32
// consecutive entries of one segment: the warp launched for position q works only if
// rank[q] % 32 == 0.  Its lanes fetch the 32 entries' metadata with one coalesced access each,
// then the row gathers (lane l owns dims 2l, 2l+1) are issued 4 entries at a time.  Segments
// longer than one unit leave per-unit partials in `unit_part`; the last unit to arrive
// (per-segment ticket) adds them in unit order, so the sum is order-deterministic.
// Extra CTAs at the end of the grid reduce grad(w) = sum_b dsp_b*pe_b + dsn_b*ne_b and
// grad(w_user) = sum_b dsu_b*ue_b into per-CTA partials (fixed composition, fixed order).
// ---------------------------------------------------------------------------------------------
constexpr int kRowWarps = 8;
constexpr int kUnit = 32;
constexpr int kWgradPerCta = 64;  // batch positions per w-gradient CTA

__global__ void __launch_bounds__(kRowWarps * 32)
row_grads_kernel(const float *__restrict__ Ue, const float *__restrict__ Ie,
                 const float *__restrict__ Ur, const float *__restrict__ Ir,
                 const float *__restrict__ w, const float *__restrict__ wu, const StepState *st,
                 const int32_t *u_, const int32_t *p_, const int32_t *n_, int B,
                 const float *__restrict__ d_yp, const float *__restrict__ d_yn,
                 const float *__restrict__ d_sp, const float *__restrict__ d_sn,
                 const float *__restrict__ d_su, float lam, PlanBufs planU, PlanBufs planI,
                 float *__restrict__ gU, float *__restrict__ gI, float *unit_part, int pos_ctas,
                 float *__restrict__ gw_part, float *__restrict__ gwu_part) {
  __shared__ float2 sW[kRowWarps][32], sWU[kRowWarps][32];
  const int lane = threadIdx.x & 31, wl = threadIdx.x >> 5;
  const int32_t *u = step_ids(st, u_, B, 0), *p = step_ids(st, p_, B, 1),
                *n = step_ids(st, n_, B, 2);

  if ((int)blockIdx.x >= pos_ctas) {  // ---- grad(w), grad(w_user) partials ----
    const int cta = blockIdx.x - pos_ctas;
    float2 aw = make_float2(0.f, 0.f), awu = make_float2(0.f, 0.f);
    const int b0 = cta * kWgradPerCta + wl * (kWgradPerCta / kRowWarps);
#pragma unroll 4
    for (int k = 0; k < kWgradPerCta / kRowWarps; ++k) {
      const int b = b0 + k;
      if (b < B) {
        const float dsp = d_sp[b], dsn = d_sn[b], dsu = d_su[b];
        const float2 pe = reinterpret_cast<const float2 *>(Ie + (long long)p[b] * kD)[lane];
        const float2 ne = reinterpret_cast<const float2 *>(Ie + (long long)n[b] * kD)[lane];
        const float2 ue = reinterpret_cast<const float2 *>(Ue + (long long)u[b] * kD)[lane];
        aw.x += dsp * pe.x + dsn * ne.x;
        aw.y += dsp * pe.y + dsn * ne.y;
        awu.x += dsu * ue.x;
        awu.y += dsu * ue.y;
      }
    }
    sW[wl][lane] = aw;
    sWU[wl][lane] = awu;
    __syncthreads();
    if (wl == 0) {
      float2 a = make_float2(0.f, 0.f), c = make_float2(0.f, 0.f);
#pragma unroll
      for (int k = 0; k < kRowWarps; ++k) {
        a.x += sW[k][lane].x;
        a.y += sW[k][lane].y;
        c.x += sWU[k][lane].x;
        c.y += sWU[k][lane].y;
      }
      reinterpret_cast<float2 *>(gw_part + (long long)cta * kD)[lane] = a;
      reinterpret_cast<float2 *>(gwu_part + (long long)cta * kD)[lane] = c;
    }
    return;
  }

  const int wid = blockIdx.x * kRowWarps + wl;  // one warp per batch position: users, then items
  if (wid >= 3 * B) return;
  const bool item = wid >= B;
  const int q = item ? wid - B : wid;
  const PlanBufs &pl = item ? planI : planU;
  const int rk = pl.rank[q];
  if (rk % kUnit != 0) return;
  const int slot = pl.pslot[q];
  const int s0 = pl.seg_off[slot], s1 = pl.seg_off[slot + 1];
  const int total = s1 - s0;
  const int base = s0 + rk;
  const int cnt = min(kUnit, s1 - base);
  const long long r = pl.uniq_rows[slot];
  const float2 bias = reinterpret_cast<const float2 *>(item ? w : wu)[lane];
  const float2 raw = reinterpret_cast<const float2 *>((item ? Ir : Ur) + r * kD)[lane];
  const float lx = lam * raw.x, ly = lam * raw.y;

  // metadata of this unit's entries, one entry per lane
  int ra = 0, rb = 0;
  float ca = 0.f, cb = 0.f, cs = 0.f;
  if (lane < cnt) {
    const int qq = pl.seg_pos[base + lane];
    if (!item) {
      ra = p[qq];
      rb = n[qq];
      ca = d_yp[qq];
      cb = d_yn[qq];
      cs = d_su[qq];
    } else {
      const bool is_pos = qq < B;
      const int b = is_pos ? qq : qq - B;
      ra = u[b];
      ca = is_pos ? d_yp[b] : d_yn[b];
      cs = is_pos ? d_sp[b] : d_sn[b];
    }
  }
  const float *tabA = item ? Ue : Ie;
  float2 g = make_float2(0.f, 0.f);
  for (int e0 = 0; e0 < cnt; e0 += 4) {
    float2 va[4], vb[4];
    float fa[4], fb[4], fs[4];
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const int e = min(e0 + k, cnt - 1);
      const int ia = __shfl_sync(0xffffffffu, ra, e);
      const int ib = __shfl_sync(0xffffffffu, rb, e);
      fa[k] = __shfl_sync(0xffffffffu, ca, e);
      fb[k] = __shfl_sync(0xffffffffu, cb, e);
      fs[k] = __shfl_sync(0xffffffffu, cs, e);
      va[k] = reinterpret_cast<const float2 *>(tabA + (long long)ia * kD)[lane];
      vb[k] = item ? make_float2(0.f, 0.f)
                   : reinterpret_cast<const float2 *>(Ie + (long long)ib * kD)[lane];
    }
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      if (e0 + k < cnt) {
        float x = fa[k] * va[k].x + fb[k] * vb[k].x + fs[k] * bias.x;
        float y = fa[k] * va[k].y + fb[k] * vb[k].y + fs[k] * bias.y;
        if (lam != 0.f) {
          x += lx;
          y += ly;
        }
        g.x += x;
        g.y += y;
      }
    }
  }
  float *gout = (item ? gI : gU) + (long long)slot * kD;
  if (total <= kUnit) {
    reinterpret_cast<float2 *>(gout)[lane] = g;
    return;
  }
  // multi-unit segment: publish the partial, the last unit to arrive folds them in unit order
  float *mypart = unit_part + ((long long)(item ? B : 0) + q) * kD;
  reinterpret_cast<float2 *>(mypart)[lane] = g;
  __threadfence();
  int ticket = 0;
  if (lane == 0) ticket = atomicAdd(&pl.done[slot], 1);
  ticket = __shfl_sync(0xffffffffu, ticket, 0);
  const int n_units = (total + kUnit - 1) / kUnit;
  if (ticket != n_units - 1) return;
  __threadfence();
  float2 acc = make_float2(0.f, 0.f);
  for (int k = 0; k < n_units; ++k) {
    const int qk = pl.seg_pos[s0 + k * kUnit];
    const float2 v = __ldcg(reinterpret_cast<const float2 *>(
                                unit_part + ((long long)(item ? B : 0) + qk) * kD) + lane);
    acc.x += v.x;
    acc.y += v.y;
  }
  reinterpret_cast<float2 *>(gout)[lane] = acc;
}

int row_grads_max_parts(int B) { return (B + kWgradPerCta - 1) / kWgradPerCta; }

int launch_row_grads(const float *Ue, const float *Ie, const float *Ur, const float *Ir,
                     const float *w, const float *wu, const StepState *st, const int32_t *u,
                     const int32_t *p, const int32_t *n, int B, const float *d_yp,
                     const float *d_yn, const float *d_sp, const float *d_sn, const float *d_su,
                     float lam, PlanBufs planU, PlanBufs planI, float *gU, float *gI,
                     float *unit_part, float *gw_part, float *gwu_part, int *n_part,
                     cudaStream_t s) {
  const int pos_ctas = (3 * B + kRowWarps - 1) / kRowWarps;
  const int w_ctas = row_grads_max_parts(B);
  row_grads_kernel<<<pos_ctas + w_ctas, kRowWarps * 32, 0, s>>>(
      Ue, Ie, Ur, Ir, w, wu, st, u, p, n, B, d_yp, d_yn, d_sp, d_sn, d_su, lam, planU, planI, gU,
      gI, unit_part, pos_ctas, gw_part, gwu_part);
  MACR_LAUNCH_CHECK();
  if (n_part) *n_part = w_ctas;
  return MACR_OK;
}

// Adam on the touched rows (gradient known); clears the rows' bitmap bits for the next step
__global__ void __launch_bounds__(256)
adam_rows_kernel(float *U, float *mU, float *vU, PlanBufs planU, const float *__restrict__ gU,
                 uint32_t *bmU, float *I, float *mI, float *vI, PlanBufs planI,
                 const float *__restrict__ gI, uint32_t *bmI, int maxU, float lr_or_lrt,
                 const StepState *st, float b1, float b2, float eps) {
  const int lane = threadIdx.x & 31;
  const int wid = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  const bool item = wid >= maxU;
  const int slot = item ? wid - maxU : wid;
  const PlanBufs &pl = item ? planI : planU;
  if (slot >= *pl.n_uniq) return;
  const float lr_t = step_lr_t(st, lr_or_lrt);
  const float omb1 = __fsub_rn(1.0f, b1), omb2 = __fsub_rn(1.0f, b2);
  const long long r = pl.uniq_rows[slot];
  float2 *pv = reinterpret_cast<float2 *>((item ? I : U) + r * kD) + lane;
  float2 *pm = reinterpret_cast<float2 *>((item ? mI : mU) + r * kD) + lane;
  float2 *pvv = reinterpret_cast<float2 *>((item ? vI : vU) + r * kD) + lane;
  const float2 g = reinterpret_cast<const float2 *>((item ? gI : gU) + (long long)slot * kD)[lane];
  float2 x = *pv, m = *pm, v = *pvv;
  adam_with_grad(x.x, m.x, v.x, g.x, lr_t, b1, b2, omb1, omb2, eps);
  adam_with_grad(x.y, m.y, v.y, g.y, lr_t, b1, b2, omb1, omb2, eps);
  *pv = x;
  *pm = m;
  *pvv = v;
  uint32_t *bm = item ? bmI : bmU;
  if (bm && lane == 0) atomicAnd(&bm[r >> 5], ~(1u << (r & 31)));
}

int launch_adam_rows2(float *U, float *mU, float *vU, PlanBufs planU, const float *gU,
                      uint32_t *bmU, float *I, float *mI, float *vI, PlanBufs planI,
                      const float *gI, uint32_t *bmI, int max_rows, float lr_t,
                      const StepState *st, float b1, float b2, float eps, cudaStream_t s) {
  // max_rows = B: up to B unique user rows then up to 2B unique item rows
  const int warps = 3 * max_rows;
  adam_rows_kernel<<<(warps + 7) / 8, 256, 0, s>>>(U, mU, vU, planU, gU, bmU, I, mI, vI, planI, gI,
                                                   bmI, max_rows, lr_t, st, b1, b2, eps);
  MACR_LAUNCH_CHECK();
  return MACR_OK;
}

// ApplyAdam on w and w_user: fold the per-CTA partials (fixed order), then
// m += (g-m)(1-b1); v += (g*g-v)(1-b2); var -= (m*lr_t)/(sqrt(v)+eps)   (training_ops.cc)
__global__ void __launch_bounds__(1024)
adam_vec2_kernel(float *w, float *mw, float *vw, float *wu, float *mwu, float *vwu,
                 const float *__restrict__ gw_part, const float *__restrict__ gwu_part, int n_part,
                 float lr_or_lrt, const StepState *st, float b1, float b2, float eps) {
  __shared__ float sh[2][16][kD];
  const int k = threadIdx.x & 63, grp = threadIdx.x >> 6;  // 16 groups x 64 dims
  const int per = (n_part + 15) / 16;
  const int lo = grp * per, hi = min(n_part, lo + per);
  float a = 0.f, c = 0.f;
  for (int q = lo; q < hi; ++q) {
    a += gw_part[(long long)q * kD + k];
    c += gwu_part[(long long)q * kD + k];
  }
  sh[0][grp][k] = a;
  sh[1][grp][k] = c;
  __syncthreads();
  if (threadIdx.x < 2 * kD) {
    const int which = threadIdx.x >> 6;
    float g = 0.f;
#pragma unroll
    for (int q = 0; q < 16; ++q) g += sh[which][q][k];
    const float lr_t = step_lr_t(st, lr_or_lrt);
    const float omb1 = __fsub_rn(1.0f, b1), omb2 = __fsub_rn(1.0f, b2);
    float *var = which ? wu : w, *m = which ? mwu : mw, *v = which ? vwu : vw;
    const float mn = __fadd_rn(m[k], __fmul_rn(__fsub_rn(g, m[k]), omb1));
    const float vn = __fadd_rn(v[k], __fmul_rn(__fsub_rn(__fmul_rn(g, g), v[k]), omb2));
    m[k] = mn;
    v[k] = vn;
    var[k] = __fsub_rn(var[k], __fdiv_rn(__fmul_rn(mn, lr_t), __fadd_rn(__fsqrt_rn(vn), eps)));
  }
}

int launch_adam_vec2(float *w, float *mw, float *vw, float *wu, float *mwu, float *vwu,
                     const float *gw_part, const float *gwu_part, int n_part, float lr_t,
                     const StepState *st, float b1, float b2, float eps, cudaStream_t s) {
  adam_vec2_kernel<<<1, 1024, 0, s>>>(w, mw, vw, wu, mwu, vwu, gw_part, gwu_part, n_part, lr_t, st,
                                      b1, b2, eps);
  MACR_LAUNCH_CHECK();
  return MACR_OK;
}

// every element has a gradient (LightGCN tables): sparse formula over all rows
__global__ void __launch_bounds__(256)
adam_dense_kernel(float4 *var, float4 *m, float4 *v, const float4 *__restrict__ grad,
                  long long n4, float lr_or_lrt, const StepState *st, float b1, float b2,
                  float eps) {
  const float lr_t = step_lr_t(st, lr_or_lrt);
  const float omb1 = __fsub_rn(1.0f, b1), omb2 = __fsub_rn(1.0f, b2);
  const long long stride = (long long)gridDim.x * blockDim.x;
  for (long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x; e < n4; e += stride) {
    float4 x = ld_stream(var + e), mm = ld_stream(m + e), vv = ld_stream(v + e);
    const float4 g = ld_stream(grad + e);
    adam_with_grad(x.x, mm.x, vv.x, g.x, lr_t, b1, b2, omb1, omb2, eps);
    adam_with_grad(x.y, mm.y, vv.y, g.y, lr_t, b1, b2, omb1, omb2, eps);
    adam_with_grad(x.z, mm.z, vv.z, g.z, lr_t, b1, b2, omb1, omb2, eps);
    adam_with_grad(x.w, mm.w, vv.w, g.w, lr_t, b1, b2, omb1, omb2, eps);
    st_stream(var + e, x);
    st_stream(m + e, mm);
    st_stream(v + e, vv);
  }
}

int launch_adam_dense(float *var, float *m, float *v, const float *grad, int64_t n_elems,
                      float lr_t, const StepState *st, float b1, float b2, float eps,
                      cudaStream_t s) {
  const long long n4 = n_elems / 4;
  if (n4 == 0) return MACR_OK;
  long long blocks = (n4 + 255) / 256;
  const long long cap = (long long)sm_count() * 16;
  if (blocks > cap) blocks = cap;
  adam_dense_kernel<<<(int)blocks, 256, 0, s>>>((float4 *)var, (float4 *)m, (float4 *)v,
                                                (const float4 *)grad, n4, lr_t, st, b1, b2, eps);
  MACR_LAUNCH_CHECK();
  return MACR_OK;
}

// end-of-step state advance (adam.py _finish: beta powers *= beta) -- one thread
__global__ void step_advance_kernel(StepState *st, float b1, float b2, int train) {
  if (threadIdx.x == 0 && blockIdx.x == 0) {
    if (train) {
      st->b1p = __fmul_rn(st->b1p, b1);
      st->b2p = __fmul_rn(st->b2p, b2);
      st->t += 1;
    }
    st->step_idx += 1;
  }
}

int launch_step_state(StepState *st, int, float, float b1, float b2, int train, cudaStream_t s) {
  step_advance_kernel<<<1, 32, 0, s>>>(st, b1, b2, train);
  MACR_LAUNCH_CHECK();
  return MACR_OK;
}

// ---------------------------------------------------------------------------------------------
// last kernel of a step (one CTA): ApplyAdam on w / w_user from the per-CTA gradient partials,
// the loss reduction, and the step-state advance (adam.py _finish: beta powers *= beta).
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(1024)
step_tail_kernel(float *w, float *mw, float *vw, float *wu, float *mwu, float *vwu,
                 const float *__restrict__ gw_part, const float *__restrict__ gwu_part, int n_part,
                 const float *__restrict__ losspart, int nparts, const float *__restrict__ litem,
                 const float *__restrict__ luser, const float *__restrict__ regsq, int B,
                 macr_hparams hp, StepState *st, int train) {
  __shared__ double sh[1024];
  __shared__ float shg[2][8][kD];
  const int tid = threadIdx.x;
  const float lr_t = step_lr_t(st, hp.lr);
  if (train) {
    const int k = tid & 63, which = (tid >> 6) & 1, grp = tid >> 7;  // 8 groups x 2 vectors x 64
    const float *src = which ? gwu_part : gw_part;
    float a = 0.f;
    for (int q = grp; q < n_part; q += 8) a += src[(long long)q * kD + k];
    shg[which][grp][k] = a;
  }
  double a0 = 0, a1 = 0, a2 = 0, a3 = 0;
  for (int k = tid; k < nparts; k += 1024) a0 += losspart[k];
  for (int k = tid; k < B; k += 1024) {
    a1 += litem[k];
    a2 += luser[k];
    a3 += regsq[k];
  }
  // one fixed-shape reduction tree for the four sums: warp shuffles, then the 32 warp leaders
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    a0 += __shfl_xor_sync(0xffffffffu, a0, o);
    a1 += __shfl_xor_sync(0xffffffffu, a1, o);
    a2 += __shfl_xor_sync(0xffffffffu, a2, o);
    a3 += __shfl_xor_sync(0xffffffffu, a3, o);
  }
  if ((tid & 31) == 0) {
    sh[(tid >> 5) * 4 + 0] = a0;
    sh[(tid >> 5) * 4 + 1] = a1;
    sh[(tid >> 5) * 4 + 2] = a2;
    sh[(tid >> 5) * 4 + 3] = a3;
  }
  __syncthreads();  // also publishes shg
  if (tid == 0) {
    a0 = a1 = a2 = a3 = 0;
    for (int k = 0; k < 32; ++k) {
      a0 += sh[k * 4 + 0];
      a1 += sh[k * 4 + 1];
      a2 += sh[k * 4 + 2];
      a3 += sh[k * 4 + 3];
    }
  }
  if (train && tid < 2 * kD) {
    const int k = tid & 63, which = tid >> 6;
    float g = 0.f;
#pragma unroll
    for (int q = 0; q < 8; ++q) g += shg[which][q][k];
    const float omb1 = __fsub_rn(1.0f, hp.beta1), omb2 = __fsub_rn(1.0f, hp.beta2);
    float *var = which ? wu : w, *m = which ? mwu : mw, *v = which ? vwu : vw;
    const float mn = __fadd_rn(m[k], __fmul_rn(__fsub_rn(g, m[k]), omb1));
    const float vn = __fadd_rn(v[k], __fmul_rn(__fsub_rn(__fmul_rn(g, g), v[k]), omb2));
    m[k] = mn;
    v[k] = vn;
    var[k] = __fsub_rn(var[k], __fdiv_rn(__fmul_rn(mn, lr_t), __fadd_rn(__fsqrt_rn(vn), hp.eps)));
  }
  __syncthreads();  // every lr_t read of this step is done
  if (tid == 0) {
    const double invB = 1.0 / (double)B;
    const float l_ori = (float)(-0.6931471805599453 * a0 * invB * invB);
    const float l_item = (float)(a1 * invB), l_user = (float)(a2 * invB);
    const float reg = hp.decay * ((float)(a3 * 0.5) / (float)hp.batch_size_flag);
    const float mf = l_ori + hp.alpha * l_item + hp.beta * l_user;
    float *out = st->loss_base + st->step_idx * 4;
    out[0] = mf + reg;
    out[1] = mf;
    out[2] = reg;
    out[3] = l_ori;
    if (train) {
      st->b1p = __fmul_rn(st->b1p, hp.beta1);
      st->b2p = __fmul_rn(st->b2p, hp.beta2);
      st->t += 1;
    }
    st->step_idx += 1;
  }
}

int launch_step_tail(float *w, float *mw, float *vw, float *wu, float *mwu, float *vwu,
                     const float *gw_part, const float *gwu_part, int n_part, const GridWs &ws,
                     const float *regsq, int B, const macr_hparams &hp, StepState *st, int train,
                     cudaStream_t s) {
  step_tail_kernel<<<1, 1024, 0, s>>>(w, mw, vw, wu, mwu, vwu, gw_part, gwu_part, n_part,
                                      ws.losspart, ws.nblk * ws.nblk, ws.litem, ws.luser, regsq, B,
                                      hp, st, train);
  MACR_LAUNCH_CHECK();
  return MACR_OK;
}

}  // namespace macr
