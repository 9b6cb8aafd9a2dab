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
// sigmoids, add the alpha / beta branch gradients (model.py:213-217)
__global__ void __launch_bounds__(256)
grid_finalize_kernel(const float *__restrict__ sp, const float *__restrict__ sn,
                     const float *__restrict__ su, int B, int Bpad, int nblk, float alpha,
                     float beta, const float *__restrict__ rowP_part,
                     const float *__restrict__ rowN_part, const float *__restrict__ colP_part,
                     const float *__restrict__ colN_part, float *__restrict__ d_yp,
                     float *__restrict__ d_yn, float *__restrict__ d_sp, float *__restrict__ d_sn,
                     float *__restrict__ d_su, float *__restrict__ litem,
                     float *__restrict__ luser, int want_grad) {
  const int b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= B) return;
  const float a = sigmoid_precise(sp[b]), an = sigmoid_precise(sn[b]),
              g = sigmoid_precise(su[b]);
  const float ea = a + kBceEps, ean = (1.0f - an) + kBceEps;
  const float eg = g + kBceEps, eg1 = (1.0f - g) + kBceEps;
  litem[b] = -logf(ea) - logf(ean);
  luser[b] = -logf(eg) - logf(eg1);
  if (!want_grad) return;
  float rp = 0.f, rn = 0.f, cp = 0.f, cn = 0.f;
  for (int k = 0; k < nblk; ++k) {
    rp += rowP_part[(size_t)k * Bpad + b];
    rn += rowN_part[(size_t)k * Bpad + b];
    cp += colP_part[(size_t)k * Bpad + b];
    cn += colN_part[(size_t)k * Bpad + b];
  }
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
  grid_finalize_kernel<<<(B + 255) / 256, 256, 0, s>>>(sp, sn, su, B, ws.Bpad, ws.nblk, alpha,
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
// K5a: batch plan.  One CTA per table sorts (row<<32 | position) keys in shared memory
// (bitonic network), then marks segment heads and compacts them with a block scan.  The order
// inside a segment is ascending position, i.e. the order TF's unsorted_segment_sum adds in.
// ---------------------------------------------------------------------------------------------
struct PlanTable {
  const int32_t *ids;
  int ids_off, n_ids;
  long long rows;
  PlanBufs out;
  uint32_t *bitmap;
};

static int next_pow2(int n) {
  int p = 1024;
  while (p < n) p <<= 1;
  return p;
}
size_t plan_ws_bytes(int) { return 16; }

__global__ void __launch_bounds__(1024)
batch_plan_kernel(PlanTable t0, PlanTable t1, const StepState *st, int B, int npow2) {
  extern __shared__ __align__(16) unsigned long long keys[];
  __shared__ int warp_tot[32];
  __shared__ int s_total;
  const PlanTable t = blockIdx.x == 0 ? t0 : t1;
  if (t.n_ids == 0) return;
  const int tid = threadIdx.x;
  const int32_t *ids = (st ? st->ids_base + st->step_idx * 3LL * B : t.ids) + t.ids_off;
  for (int e = tid; e < npow2; e += 1024)
    keys[e] = e < t.n_ids
                  ? (((unsigned long long)(uint32_t)ids[e]) << 32) | (unsigned long long)(uint32_t)e
                  : ~0ull;
  __syncthreads();
  for (int k = 2; k <= npow2; k <<= 1) {
    for (int j = k >> 1; j > 0; j >>= 1) {
      for (int x = tid; x < (npow2 >> 1); x += 1024) {
        const int i = ((x & ~(j - 1)) << 1) | (x & (j - 1));
        const int l = i | j;
        const unsigned long long ka = keys[i], kb = keys[l];
        const bool up = (i & k) == 0;
        if ((ka > kb) == up) {
          keys[i] = kb;
          keys[l] = ka;
        }
      }
      __syncthreads();
    }
  }
  // segment heads + compaction
  const int per = npow2 >> 10;
  const int e0 = tid * per;
  int cnt = 0;
  for (int e = e0; e < e0 + per; ++e) {
    if (e < t.n_ids) {
      const uint32_t r = (uint32_t)(keys[e] >> 32);
      cnt += (e == 0 || r != (uint32_t)(keys[e - 1] >> 32)) ? 1 : 0;
    }
  }
  int incl = cnt;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    const int v = __shfl_up_sync(0xffffffffu, incl, o);
    if ((tid & 31) >= o) incl += v;
  }
  if ((tid & 31) == 31) warp_tot[tid >> 5] = incl;
  __syncthreads();
  if (tid < 32) {
    int v = warp_tot[tid];
    int inc2 = v;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const int y = __shfl_up_sync(0xffffffffu, inc2, o);
      if (tid >= o) inc2 += y;
    }
    warp_tot[tid] = inc2 - v;  // exclusive
    if (tid == 31) s_total = inc2;
  }
  __syncthreads();
  int slot = warp_tot[tid >> 5] + incl - cnt;
  for (int e = e0; e < e0 + per; ++e) {
    if (e < t.n_ids) {
      const unsigned long long k = keys[e];
      const uint32_t r = (uint32_t)(k >> 32);
      t.out.seg_pos[e] = (int32_t)(uint32_t)k;
      if (e == 0 || r != (uint32_t)(keys[e - 1] >> 32)) {
        t.out.uniq_rows[slot] = (int32_t)r;
        t.out.seg_off[slot] = e;
        if (t.bitmap) atomicOr(&t.bitmap[r >> 5], 1u << (r & 31));
        ++slot;
      }
    }
  }
  if (tid == 0) {
    t.out.seg_off[s_total] = t.n_ids;
    *t.out.n_uniq = s_total;
  }
}

int plan_init() {  // opt in to 128 KB dynamic shared memory once (outside any stream capture)
  static bool attr_set = false;
  if (!attr_set) {
    MACR_CUDA(cudaFuncSetAttribute(batch_plan_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                   16384 * 8));
    attr_set = true;
  }
  return MACR_OK;
}

int launch_batch_plan2(const int32_t *ids0, const StepState *st, int ids0_off, int n_ids0,
                       int64_t rows0, PlanBufs out0, uint32_t *bitmap0, const int32_t *ids1,
                       int ids1_off, int n_ids1, int64_t rows1, PlanBufs out1, uint32_t *bitmap1,
                       void *, cudaStream_t s) {
  const int nmax = n_ids0 > n_ids1 ? n_ids0 : n_ids1;
  MACR_CHECK_ARG(nmax <= 16384, "batch plan supports at most 16384 ids per table (batch <= 8192)");
  const int npow2 = next_pow2(nmax);
  const size_t smem = (size_t)npow2 * sizeof(unsigned long long);
  int rci = plan_init();
  if (rci) return rci;
  PlanTable t0{ids0, ids0_off, n_ids0, rows0, out0, bitmap0};
  PlanTable t1{ids1, ids1_off, n_ids1, rows1, out1, bitmap1};
  const int B = st ? n_ids0 : 0;  // trainer convention: table 0 = users (B ids)
  batch_plan_kernel<<<n_ids1 > 0 ? 2 : 1, 1024, smem, s>>>(t0, t1, st, B, npow2);
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
// row gradients of the unique touched rows.  One warp per unique row walks its segment in
// ascending batch position and sums  d(loss)/d(row)  in registers (lane l owns dims 2l, 2l+1):
//   user row r :  sum_b  dyp_b*Ie[p_b] + dyn_b*Ie[n_b] + dsu_b*w_user (+ lam*Ur[r])
//   item row r :  sum_q  dy_q*Ue[u_b] + ds_q*w (+ lam*Ir[r]),  q<B: pos role, q>=B: neg role
// and the warp's share of grad(w_user) = (sum dsu)*Ue[r] / grad(w) = (sum ds)*Ie[r].
// ---------------------------------------------------------------------------------------------
constexpr int kRowWarps = 8;

__global__ void __launch_bounds__(kRowWarps * 32)
row_grads_kernel(const float *__restrict__ Ue, const float *__restrict__ Ie,
                 const float *__restrict__ Ur, const float *__restrict__ Ir,
                 const float *__restrict__ w, const float *__restrict__ wu, const StepState *st,
                 const int32_t *u_, const int32_t *p_, const int32_t *n_, int B,
                 const float *__restrict__ d_yp, const float *__restrict__ d_yn,
                 const float *__restrict__ d_sp, const float *__restrict__ d_sn,
                 const float *__restrict__ d_su, float lam, PlanBufs planU, PlanBufs planI,
                 float *__restrict__ gU, float *__restrict__ gI, float *__restrict__ gw_part,
                 float *__restrict__ gwu_part) {
  __shared__ float2 sW[kRowWarps][32], sWU[kRowWarps][32];
  const int lane = threadIdx.x & 31, wl = threadIdx.x >> 5;
  const int wid = blockIdx.x * kRowWarps + wl;
  const int32_t *u = step_ids(st, u_, B, 0), *p = step_ids(st, p_, B, 1),
                *n = step_ids(st, n_, B, 2);
  float2 cw = make_float2(0.f, 0.f), cwu = make_float2(0.f, 0.f);
  if (wid < B) {
    if (wid < *planU.n_uniq) {
      const long long r = planU.uniq_rows[wid];
      const int s0 = planU.seg_off[wid], s1 = planU.seg_off[wid + 1];
      const float2 wuv = reinterpret_cast<const float2 *>(wu)[lane];
      const float2 raw = reinterpret_cast<const float2 *>(Ur + r * kD)[lane];
      float2 g = make_float2(0.f, 0.f);
      float csum = 0.f;
      for (int e = s0; e < s1; ++e) {
        const int b = planU.seg_pos[e];
        const float dyp = d_yp[b], dyn = d_yn[b], dsu = d_su[b];
        const float2 pe = reinterpret_cast<const float2 *>(Ie + (long long)p[b] * kD)[lane];
        const float2 ne = reinterpret_cast<const float2 *>(Ie + (long long)n[b] * kD)[lane];
        float x = dyp * pe.x + dyn * ne.x + dsu * wuv.x;
        float y = dyp * pe.y + dyn * ne.y + dsu * wuv.y;
        if (lam != 0.f) {
          x += lam * raw.x;
          y += lam * raw.y;
        }
        g.x += x;
        g.y += y;
        csum += dsu;
      }
      reinterpret_cast<float2 *>(gU + (long long)wid * kD)[lane] = g;
      const float2 ue = reinterpret_cast<const float2 *>(Ue + r * kD)[lane];
      cwu = make_float2(csum * ue.x, csum * ue.y);
    }
  } else {
    const int wi = wid - B;
    if (wi < 2 * B && wi < *planI.n_uniq) {
      const long long r = planI.uniq_rows[wi];
      const int s0 = planI.seg_off[wi], s1 = planI.seg_off[wi + 1];
      const float2 wv = reinterpret_cast<const float2 *>(w)[lane];
      const float2 raw = reinterpret_cast<const float2 *>(Ir + r * kD)[lane];
      float2 g = make_float2(0.f, 0.f);
      float csum = 0.f;
      for (int e = s0; e < s1; ++e) {
        const int q = planI.seg_pos[e];
        const bool is_pos = q < B;
        const int b = is_pos ? q : q - B;
        const float dy = is_pos ? d_yp[b] : d_yn[b];
        const float ds = is_pos ? d_sp[b] : d_sn[b];
        const float2 ue = reinterpret_cast<const float2 *>(Ue + (long long)u[b] * kD)[lane];
        float x = dy * ue.x + ds * wv.x;
        float y = dy * ue.y + ds * wv.y;
        if (lam != 0.f) {
          x += lam * raw.x;
          y += lam * raw.y;
        }
        g.x += x;
        g.y += y;
        csum += ds;
      }
      reinterpret_cast<float2 *>(gI + (long long)wi * kD)[lane] = g;
      const float2 ie = reinterpret_cast<const float2 *>(Ie + r * kD)[lane];
      cw = make_float2(csum * ie.x, csum * ie.y);
    }
  }
  sW[wl][lane] = cw;
  sWU[wl][lane] = cwu;
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
    reinterpret_cast<float2 *>(gw_part + (long long)blockIdx.x * kD)[lane] = a;
    reinterpret_cast<float2 *>(gwu_part + (long long)blockIdx.x * kD)[lane] = c;
  }
}

int row_grads_max_parts(int B) { return (3 * B + kRowWarps - 1) / kRowWarps; }

int launch_row_grads(const float *Ue, const float *Ie, const float *Ur, const float *Ir,
                     const float *w, const float *wu, const StepState *st, const int32_t *u,
                     const int32_t *p, const int32_t *n, int B, const float *d_yp,
                     const float *d_yn, const float *d_sp, const float *d_sn, const float *d_su,
                     float lam, PlanBufs planU, PlanBufs planI, float *gU, float *gI,
                     float *gw_part, float *gwu_part, int *n_part, cudaStream_t s) {
  const int blocks = row_grads_max_parts(B);
  row_grads_kernel<<<blocks, kRowWarps * 32, 0, s>>>(Ue, Ie, Ur, Ir, w, wu, st, u, p, n, B, d_yp,
                                                     d_yn, d_sp, d_sn, d_su, lam, planU, planI, gU,
                                                     gI, gw_part, gwu_part);
  MACR_LAUNCH_CHECK();
  if (n_part) *n_part = blocks;
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

}  // namespace macr
