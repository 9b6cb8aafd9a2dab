// train_kernels.cu -- sm_100a kernels of the MACR training step
//   K1+K2 gather + five dots        (macr_mf/model.py:35-37,186-187,194-196,219)
//   K3    B x B gated BCE grid      (model.py:204-217 + its autodiff)
//   K5a   batch plan (dedup)        (TF-1.14 optimizer.py _deduplicate_indexed_slices)
//         row gradients + fused Adam on the touched rows + step tail
//   K5b   Adam, TF dense semantics  (TF-1.14 adam.py _apply_sparse_shared / ApplyAdam)
// Design notes and rooflines: DESIGN.md sections 3-4.
#include <stdlib.h>

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

struct GateOut {
  float *gA, *gAN, *gG, *litem, *luser;
  int item_only;
};
__device__ __forceinline__ void gate_values(float sp, float sn, float su, float &a, float &an,
                                            float &g, float &litem, float &luser);

// ---------------------------------------------------------------------------------------------
// K1+K2: one warp per triple; lane l owns elements [2l, 2l+1] of every 64-wide row
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
gather_dots_kernel(const float *__restrict__ Ue, const float *__restrict__ Ie,
                   const float *__restrict__ Ur, const float *__restrict__ Ir,
                   const float *__restrict__ w, const float *__restrict__ wu,
                   const int32_t *u_, const int32_t *p_, const int32_t *n_, const StepState *st,
                   int B, float *__restrict__ yp, float *__restrict__ yn, float *__restrict__ sp,
                   float *__restrict__ sn, float *__restrict__ su, float *__restrict__ regsq,
                   float *__restrict__ snap, GateOut gates) {
  pdl_trigger();  // the grid kernel's CTAs may take their slots (they wait for this grid's results)
  const int lane = threadIdx.x & 31;
  const int b = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (b >= B) return;
  const int32_t *u = step_ids(st, u_, B, 0), *p = step_ids(st, p_, B, 1),
                *n = step_ids(st, n_, B, 2);
  const long long ur = u[b], pr = p[b], nr = n[b];
  const float2 ue = reinterpret_cast<const float2 *>(Ue + ur * kD)[lane];
  const float2 pe = reinterpret_cast<const float2 *>(Ie + pr * kD)[lane];
  const float2 ne = reinterpret_cast<const float2 *>(Ie + nr * kD)[lane];
  if (snap) {  // row snapshot [3][B][64] (users | pos | neg): row_grads reads these, so the Adam
               // update of a row can run while other rows' gradients are still being formed
    reinterpret_cast<float2 *>(snap + (long long)b * kD)[lane] = ue;
    reinterpret_cast<float2 *>(snap + ((long long)B + b) * kD)[lane] = pe;
    reinterpret_cast<float2 *>(snap + (2LL * B + b) * kD)[lane] = ne;
  }
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
    if (gates.gA) {  // gates + branch losses for the B x B grid (computed once per position)
      float a, an, g, li, lu;
      gate_values(a2, a3, a4, a, an, g, li, lu);
      if (gates.item_only) g = 1.0f, lu = 0.0f;  // no user branch in this graph
      gates.gA[b] = a;
      gates.gAN[b] = an;
      gates.gG[b] = g;
      gates.litem[b] = li;
      gates.luser[b] = lu;
    }
  }
}

int launch_gather_dots(const float *Ue, const float *Ie, const float *Ur, const float *Ir,
                       const float *w, const float *wu, const int32_t *u, const int32_t *p,
                       const int32_t *n, const StepState *st, int B, float *yp, float *yn,
                       float *sp, float *sn, float *su, float *regsq, float *snap,
                       const GridWs *gates, cudaStream_t s) {
  const int wpb = 8;
  GateOut go{};
  if (gates)
    go = GateOut{gates->gA, gates->gAN, gates->gG, gates->litem, gates->luser,
                 gates->item_gate_only};
  gather_dots_kernel<<<(B + wpb - 1) / wpb, wpb * 32, 0, s>>>(
      Ue, Ie, Ur, Ir, w, wu, u, p, n, st, B, yp, yn, sp, sn, su, regsq, snap, go);
  MACR_LAUNCH_CHECK();
  return MACR_OK;
}

// ---------------------------------------------------------------------------------------------
// K3: the B x B grid.  CTA tile = (16*RI) rows x (16*RJ) columns, 256 threads as 16(tx) x 16(ty),
// each thread owns an RI x RJ register micro-tile (rows ty+16r, columns tx+16c): the whole inner
// loop runs on registers; row / column partial sums meet in shared memory once per tile.
// The [B,B] matrices are never written anywhere.
//
// Per (i,j) pair the reference evaluates (model.py:204-211), with c_i = sig(sp_i)*sig(su_i) etc.
//   P = yp_j*c_i   s = sig(P)  lossP = -log(s+1e-10)      dlossP/dP = -s(1-s)/(s+1e-10)
//   N = yn_j*cn_i  t = sig(N)  lossN = -log((1-t)+1e-10)  dlossN/dN =  t(1-t)/((1-t)+1e-10)
// Fast path (s >= 2^-9 and 1-t >= 2^-9, where the fp32 "+1e-10" is a no-op exactly as in the
// reference's own fp32 arithmetic):  with eP = exp(-P), eN = exp(-N), D = (1+eP)(1+eN)
//   lossP + lossN = ln D + N           dlossP/dP = (1+eN)/D - 1        dlossN/dN = (1+eP)/D
// i.e. 2 ex2 + 1 rcp + 1 lg2 = 4 MUFU and 13 FP32 instructions per pair; the "+N" term is a
// rank-1 sum added once per tile.  A tile whose column scores are bounded so that every pair is
// in the fast range (|yp_j|*max_i c_i <= 6.2, |yn_j|*max_i cn_i <= 5.5) runs without any
// per-pair range test; other tiles test every pair and fall back to the literal formulas
// (8 MUFU) where needed.
//
// Extra CTAs at the end of the grid fold each band's partial sums in fixed order and finish the
// chain rule through the three sigmoids (see grid_fold_band) -- no separate finalize launch.
// ---------------------------------------------------------------------------------------------
struct GridOut {
  float *d_yp, *d_yn, *d_sp, *d_sn, *d_su;
  // loss outputs (written by the loss folder CTA): st != nullptr -> {loss, mf, reg, L_ori} at
  // st->loss_base + 4*st->step_idx, else {L_ori, L_item, L_user} at losses3
  const float *regsq;
  macr_hparams hp;
  const StepState *st;
  float *losses3;
};

template <bool kChecked, bool kMasked, bool kGrad>
__device__ __forceinline__ void grid_pair(float ypj, float ynj, float agl, float angl, float mk,
                                          float &lgacc, float &colP, float &colN, float &rowP,
                                          float &rowN) {
  const float xP = ypj * agl, xN = ynj * angl;  // -P*log2(e), -N*log2(e)
  const float eP = ex2_approx(xP), eN = ex2_approx(xN);
  const float DP = 1.0f + eP, DN = 1.0f + eN;
  float lg, dP, dN;
  if (!kChecked || (eP <= 500.0f && eN >= 0.00390625f && eN <= 1.0e18f)) {
    const float D = DP * DN;
    const float rr = rcp_approx(D);
    lg = -lg2_approx(D);  // + xN, added per tile as a rank-1 sum
    dN = DP * rr;
    dP = DN * rr - 1.0f;
  } else {
    const float s = rcp_approx(DP), t = rcp_approx(DN);
    const float se = s + kBceEps, q = (1.0f - t) + kBceEps;
    lg = (lg2_approx(se) + lg2_approx(q)) - xN;
    dP = -(s * (1.0f - s)) * rcp_approx(se);
    dN = (t * (1.0f - t)) * rcp_approx(q);
  }
  if (kMasked) lg *= mk;
  lgacc += lg;
  if (kGrad) {  // column sums carry the -log2(e) factor of agl / angl; it is divided out once
    colP = fmaf(dP, agl, colP);
    colN = fmaf(dN, angl, colN);
    rowP = fmaf(dP, ypj, rowP);
    rowN = fmaf(dN, ynj, rowN);
  }
}

// Four pairs of one row at once (unchecked, unmasked tiles): the four D = (1+eP)(1+eN) share ONE
// reciprocal and ONE logarithm -- lg2(D0 D1 D2 D3) is the sum of the four logarithms and
// 1/D_k = (product of the others) / (D0 D1 D2 D3) -- so a pair costs 2.5 MUFU operations instead
// of 4 (2 ex2 + 1/4 rcp + 1/4 lg2) for nine extra multiplies per quad on the FMA pipe, which has
// the slack (the kernel is bound by the 16-lane MUFU pipe).  Range: unchecked tiles have
// eP <= e^6.2, eN <= e^5.5, so D <= 1.3e5 and the product of four <= 2.4e20: far inside fp32, and
// its reciprocal is a normal number.  Each 1/D_k carries two more roundings (relative 1.2e-7).
template <bool kGrad>
__device__ __forceinline__ void grid_quad(const float *ypj, const float *ynj, float agl, float angl,
                                          float &lgacc, float *colP, float *colN, float &rowP,
                                          float &rowN) {
  float DP[4], DN[4], D[4];
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    DP[k] = 1.0f + ex2_approx(ypj[k] * agl);
    DN[k] = 1.0f + ex2_approx(ynj[k] * angl);
    D[k] = DP[k] * DN[k];
  }
  const float p01 = D[0] * D[1], p23 = D[2] * D[3], p = p01 * p23;
  lgacc -= lg2_approx(p);  // + xN, added per tile as a rank-1 sum
  if (kGrad) {
    const float inv = rcp_approx(p);
    const float i01 = p23 * inv, i23 = p01 * inv;  // 1/(D0 D1), 1/(D2 D3)
    const float rr[4] = {D[1] * i01, D[0] * i01, D[3] * i23, D[2] * i23};
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const float dN = DP[k] * rr[k];
      const float dP = fmaf(DN[k], rr[k], -1.0f);
      colP[k] = fmaf(dP, agl, colP[k]);
      colN[k] = fmaf(dN, angl, colN[k]);
      rowP = fmaf(dP, ypj[k], rowP);
      rowN = fmaf(dN, ynj[k], rowN);
    }
  }
}

// Band folds without any synchronisation cost in the tile CTAs: every partial-sum slot holds a
// sentinel bit pattern (kPartEmpty, a NaN no arithmetic produces) until its tile CTA stores the
// value -- a single 4-byte store, so a reader sees either the sentinel or the value.  One extra
// CTA per row band and per column band, placed at the END of the grid, polls its band's slots
// in fixed order (deterministic sum), re-arms them, and finishes the chain rule through the
// three sigmoids with the alpha / beta branch gradients (model.py:213-217).  The folders are
// dispatched with the last wave of tiles and never hold more than nblk_i + nblk_j CTA slots.
constexpr unsigned kPartEmpty = 0xffffffffu;

__device__ __forceinline__ float not_sentinel(float v) {
  return __float_as_uint(v) == kPartEmpty ? __uint_as_float(0x7fc00000u) : v;
}

// sum of n slots `stride` floats apart, in index order; waits for slots that are still empty
// and re-arms every slot it consumed.  Loads are issued kFoldBatch at a time (one L2 round trip
// per batch).
constexpr int kFoldBatch = 16;

__device__ __forceinline__ float take_partials(float *slot, int n, size_t stride) {
  float acc = 0.f;
  for (int k0 = 0; k0 < n; k0 += kFoldBatch) {
    unsigned bits[kFoldBatch];
    for (;;) {
      bool ready = true;
#pragma unroll
      for (int q = 0; q < kFoldBatch; ++q) {
        bits[q] = 0;
        if (k0 + q < n)
          asm volatile("ld.volatile.global.u32 %0, [%1];"
                       : "=r"(bits[q])
                       : "l"(slot + (size_t)(k0 + q) * stride));
      }
#pragma unroll
      for (int q = 0; q < kFoldBatch; ++q) ready = ready && bits[q] != kPartEmpty;
      if (ready) break;
      __nanosleep(100);
    }
#pragma unroll
    for (int q = 0; q < kFoldBatch; ++q)
      if (k0 + q < n) {
        acc += __uint_as_float(bits[q]);
        asm volatile("st.volatile.global.u32 [%0], %1;" ::"l"(slot + (size_t)(k0 + q) * stride),
                     "r"(kPartEmpty)
                     : "memory");
      }
  }
  return acc;
}

// the loss folder: sum of the tile loss partials (sentinel-armed slots like the band partials)
// and of the per-position branch losses / L2 squares -> the step's loss scalars (model.py:217-221)
__device__ void grid_fold_losses(int B, const GridWs &ws, const GridOut &out, double *sh) {
  const int tid = threadIdx.x;
  const int nparts = ws.nblk_i * ws.ngrp_j;
  // everything that does not depend on the tiles first: destination, per-position sums
  float *dst = out.losses3;
  if (out.st != nullptr) dst = out.st->loss_base + out.st->step_idx * 4;
  double a1 = 0, a2 = 0, a3 = 0;
#pragma unroll 4
  for (int k = tid; k < B; k += 256) {
    a1 += ws.litem[k];
    a2 += ws.luser[k];
    if (out.regsq) a3 += out.regsq[k];
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    a1 += __shfl_xor_sync(0xffffffffu, a1, o);
    a2 += __shfl_xor_sync(0xffffffffu, a2, o);
    a3 += __shfl_xor_sync(0xffffffffu, a3, o);
  }
  if ((tid & 31) == 0) {
    sh[(tid >> 5) * 4 + 1] = a1;
    sh[(tid >> 5) * 4 + 2] = a2;
    sh[(tid >> 5) * 4 + 3] = a3;
  }
  __syncthreads();
  const double invB = 1.0 / (double)B;
  float l_item = 0.f, l_user = 0.f, reg = 0.f;
  if (tid == 0) {
    a1 = a2 = a3 = 0;
    for (int k = 0; k < 8; ++k) {
      a1 += sh[k * 4 + 1];
      a2 += sh[k * 4 + 2];
      a3 += sh[k * 4 + 3];
    }
    l_item = (float)(a1 * invB);
    l_user = (float)(a2 * invB);
    reg = out.hp.decay * ((float)(a3 * 0.5) / (float)out.hp.batch_size_flag);
  }
  // the tile partials as they arrive
  double a0 = 0;
  if (tid < nparts) a0 = (double)take_partials(ws.losspart + tid, (nparts - tid + 255) / 256, 256);
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) a0 += __shfl_xor_sync(0xffffffffu, a0, o);
  if ((tid & 31) == 0) sh[(tid >> 5) * 4] = a0;
  __syncthreads();
  if (tid != 0) return;
  a0 = 0;
  for (int k = 0; k < 8; ++k) a0 += sh[k * 4];
  const float l_ori = (float)(-0.6931471805599453 * a0 * invB * invB);
  if (out.st == nullptr) {
    dst[0] = l_ori;
    dst[1] = l_item;
    dst[2] = l_user;
  } else {
    const float mf = l_ori + out.hp.alpha * l_item + out.hp.beta * l_user;
    dst[0] = mf + reg;
    dst[1] = mf;
    dst[2] = reg;
    dst[3] = l_ori;
  }
}

template <int TI, int TJ>
__device__ void grid_fold_band(int f, int B, float alpha, float beta, const GridWs &ws,
                               const GridOut &out, float *scratch /* >= 2*TI floats */) {
  const int tid = threadIdx.x, Bpad = ws.Bpad;
  const float invB = 1.0f / (float)B;
  const float invBB = invB * invB;
  if (f < ws.nblk_i) {  // row band f: d/d sp_i, sn_i, su_i
    const int i0 = f * TI;
    for (int t = tid; t < 2 * TI; t += 256) {
      const int which = t / TI, row = t - which * TI;
      float *src = (which ? ws.rowN : ws.rowP) + i0 + row;
      scratch[t] = take_partials(src, ws.ngrp_j, Bpad);
    }
    __syncthreads();
    for (int t = tid; t < TI; t += 256) {
      const int i = i0 + t;
      if (i >= B) continue;
      const float rp = scratch[t], rn = scratch[TI + t];
      const float a = ws.gA[i], an = ws.gAN[i], g = ws.gG[i];
      const float ea = a + kBceEps, ean = (1.0f - an) + kBceEps;
      const float eg = g + kBceEps, eg1 = (1.0f - g) + kBceEps;
      const float da = rp * invBB * g - alpha * invB / ea;
      const float dan = rn * invBB * g + alpha * invB / ean;
      const float dg = (rp * a + rn * an) * invBB + beta * invB * (1.0f / eg1 - 1.0f / eg);
      out.d_sp[i] = da * (a * (1.0f - a));
      out.d_sn[i] = dan * (an * (1.0f - an));
      out.d_su[i] = dg * (g * (1.0f - g));
    }
  } else if (f < ws.nblk_i + ws.nblk_j) {  // column band: d/d yp_j, d/d yn_j
    const int j0 = (f - ws.nblk_i) * TJ;
    for (int t = tid; t < 2 * TJ; t += 256) {
      const int which = t / TJ, colm = t - which * TJ;
      float *src = (which ? ws.colN : ws.colP) + j0 + colm;
      const float acc = take_partials(src, ws.nblk_i, Bpad);
      if (j0 + colm < B) (which ? out.d_yn : out.d_yp)[j0 + colm] = acc * invBB;
    }
  }
}

template <int RI, int RJ, int MINB, bool kMasked, bool kGrad>
__global__ void __launch_bounds__(256, MINB)
grid_bce_kernel(const float *__restrict__ yp, const float *__restrict__ yn, int B, float alpha,
                float beta, GridWs ws, GridOut out) {
  constexpr int TI = 16 * RI, TJ = 16 * RJ;
  constexpr int TMAX = TI > TJ ? TI : TJ;
  constexpr float kLog2e = 1.4426950408889634f;
  __shared__ float sAg[TI], sAng[TI];
  __shared__ float sYp[TJ], sYn[TJ];
  __shared__ float sRed[2][TMAX][17];
  __shared__ float sLoss[8];
  __shared__ float sMax[2][8];

  const int tid = threadIdx.x, tx = tid & 15, ty = tid >> 4;
  const int Bpad = ws.Bpad;
  pdl_wait();     // dots and gates of the batch (gather_dots / gate kernel) are complete
  pdl_trigger();  // row_grads' CTAs may queue behind this grid's last wave
  if ((int)blockIdx.y >= ws.nblk_i) {  // folders ride at the end of the grid
    const int f = (blockIdx.y - ws.nblk_i) * gridDim.x + blockIdx.x;
    if (f == 0)
      grid_fold_losses(B, ws, out, reinterpret_cast<double *>(&sRed[0][0][0]));
    else if (kGrad)
      grid_fold_band<TI, TJ>(f - 1, B, alpha, beta, ws, out, &sRed[0][0][0]);
    return;
  }
  const int i0 = blockIdx.y * TI;
  const int jt_begin = blockIdx.x * ws.grp_tiles;
  const int jt_end = min(ws.nblk_j, jt_begin + ws.grp_tiles);
  // gates sig(sp)*sig(su), sig(sn)*sig(su) of this row band, precomputed once per step by
  // gate_kernel / gather_dots
  for (int t = tid; t < TI; t += 256) {
    const int i = i0 + t;
    const bool ok = i < B;
    const float g = ok ? ws.gG[i] : 0.f;
    sAg[t] = ok ? ws.gA[i] * g : 0.f;
    sAng[t] = ok ? ws.gAN[i] * g : 0.f;
  }
  __syncthreads();
  float agl[RI], angl[RI], wr[RI];
  float mg = 0.f, mgn = 0.f;
#pragma unroll
  for (int r = 0; r < RI; ++r) {
    const int t = ty + 16 * r;
    const float ag = sAg[t], ang = sAng[t];
    mg = fmaxf(mg, ag);
    mgn = fmaxf(mgn, ang);
    agl[r] = -kLog2e * ag;  // exp(-P) = ex2(yp * (-c*log2e))
    angl[r] = -kLog2e * ang;
    wr[r] = (i0 + t < B) ? 1.f : 0.f;
  }
  // largest gate of the band -> can a tile skip the per-pair range test?
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    mg = fmaxf(mg, __shfl_xor_sync(0xffffffffu, mg, o));
    mgn = fmaxf(mgn, __shfl_xor_sync(0xffffffffu, mgn, o));
  }
  if ((tid & 31) == 0) {
    sMax[0][tid >> 5] = mg;
    sMax[1][tid >> 5] = mgn;
  }
  float sum_ang = 0.f;  // warp 0: sum_i ang_i of the band (rank-1 term of every tile)
  if (tid < 32) {
    for (int t = tid; t < TI; t += 32) sum_ang += sAng[t];
    sum_ang = warp_sum(sum_ang);
  }
  __syncthreads();
  mg = mgn = 0.f;
#pragma unroll
  for (int q = 0; q < 8; ++q) {
    mg = fmaxf(mg, sMax[0][q]);
    mgn = fmaxf(mgn, sMax[1][q]);
  }

  float rowP[RI], rowN[RI];
#pragma unroll
  for (int r = 0; r < RI; ++r) rowP[r] = rowN[r] = 0.f;
  float lgacc = 0.f;  // sum of log2(.) terms; loss = -ln2 * sum
  double rank1 = 0.0;  // thread 0: sum over the tiles of (sum_i -log2e*ang_i) * (sum_j yn_j)

  for (int jt = jt_begin; jt < jt_end; ++jt) {
    const int j0 = jt * TJ;
    for (int t = tid; t < TJ; t += 256) {
      const int j = j0 + t;
      const bool ok = j < B;
      sYp[t] = ok ? yp[j] : 0.f;
      sYn[t] = ok ? yn[j] : 0.f;
    }
    __syncthreads();
    float ypj[RJ], ynj[RJ], wc[RJ];
#pragma unroll
    for (int c = 0; c < RJ; ++c) {
      const int t = tx + 16 * c;
      ypj[c] = sYp[t];
      ynj[c] = sYn[t];
      wc[c] = (j0 + t < B) ? 1.f : 0.f;
    }
    bool in_range = true;
    if (tid < TJ) in_range = fabsf(sYp[tid]) * mg <= 6.2f && fabsf(sYn[tid]) * mgn <= 5.5f;
    const bool all_fast = __syncthreads_and(in_range);
    if (tid < 32) {  // rank-1 term of the tile: sum_ij xN_ij = (sum_i -log2e*ang_i) * (sum_j yn_j)
      float sy = 0.f;
      for (int t = tid; t < TJ; t += 32) sy += sYn[t];
      sy = warp_sum(sy);
      rank1 += -(double)kLog2e * (double)sum_ang * (double)sy;
    }
    float colP[RJ], colN[RJ];
#pragma unroll
    for (int c = 0; c < RJ; ++c) colP[c] = colN[c] = 0.f;

    if (all_fast && !kMasked) {
      static_assert(RJ % 4 == 0, "grid_quad takes four columns at a time");
#pragma unroll
      for (int r = 0; r < RI; ++r)
#pragma unroll
        for (int c = 0; c < RJ; c += 4)
          grid_quad<kGrad>(ypj + c, ynj + c, agl[r], angl[r], lgacc, colP + c, colN + c, rowP[r],
                           rowN[r]);
    } else if (all_fast) {
#pragma unroll
      for (int r = 0; r < RI; ++r)
#pragma unroll
        for (int c = 0; c < RJ; ++c)
          grid_pair<false, kMasked, kGrad>(ypj[c], ynj[c], agl[r], angl[r],
                                           kMasked ? wr[r] * wc[c] : 1.f, lgacc, colP[c], colN[c],
                                           rowP[r], rowN[r]);
    } else {
#pragma unroll
      for (int r = 0; r < RI; ++r)
#pragma unroll
        for (int c = 0; c < RJ; ++c)
          grid_pair<true, kMasked, kGrad>(ypj[c], ynj[c], agl[r], angl[r],
                                          kMasked ? wr[r] * wc[c] : 1.f, lgacc, colP[c], colN[c],
                                          rowP[r], rowN[r]);
    }
    if (kGrad) {  // column sums of this tile: one partial per (row band, column)
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
        float *dst = which ? ws.colN : ws.colP;
        __stcg(&dst[(size_t)blockIdx.y * Bpad + j0 + colm], not_sentinel(acc * (-1.0f / kLog2e)));
      }
    }
    __syncthreads();  // sYp / sYn / sRed are rewritten by the next tile
  }

  // ---- per-CTA reductions: row sums over the CTA's column tiles, loss partial ----------------
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
      float *dst = which ? ws.rowN : ws.rowP;
      __stcg(&dst[(size_t)blockIdx.x * Bpad + i0 + row], not_sentinel(acc));
    }
  }
  lgacc = warp_sum(lgacc);
  if ((tid & 31) == 0) sLoss[tid >> 5] = lgacc;
  __syncthreads();
  if (tid == 0) {
    float acc = 0.f;
#pragma unroll
    for (int q = 0; q < 8; ++q) acc += sLoss[q];
    __stcg(&ws.losspart[blockIdx.y * gridDim.x + blockIdx.x], not_sentinel((float)((double)acc + rank1)));
  }
}

// tile shapes: rows x cols = 16*RI x 16*RJ.  Row bands and column bands may have different widths.
static void grid_cfg(int B, int *ri, int *rj) {
  static int env_ri = -1, env_rj = -1;
  if (env_ri < 0) {
    const char *e = getenv("MACR_GRID_TILE");  // developer knob: "88", "48", "84", "44"
    env_ri = env_rj = 0;
    if (e && e[0] && e[1]) {
      env_ri = e[0] - '0';
      env_rj = e[1] - '0';
    }
  }
  if (env_ri > 0 && B >= 2048) {
    *ri = env_ri;
    *rj = env_rj;
    return;
  }
  *ri = *rj = (B >= 2048) ? 8 : 4;
}

GridWs grid_ws_layout(int B, void *base) {
  GridWs w;
  int ri, rj;
  grid_cfg(B, &ri, &rj);
  w.tile_i = 16 * ri;
  w.tile_j = 16 * rj;
  w.nblk_i = (B + w.tile_i - 1) / w.tile_i;
  w.nblk_j = (B + w.tile_j - 1) / w.tile_j;
  {
    // column tiles per CTA: the fewest tiles on the busiest CTA slot, preferring SHORT walks on a
    // tie.  Measured at B = 4096 (ncu, profiles/r2l_grid_tpc.txt): 26 us for 1, 2 or 4 tiles per
    // CTA -- the kernel is bound by MIO/MUFU latency at 4 warps per scheduler (XU pipe 45 %, issue
    // slots 50 % busy), not by its per-CTA prologue -- and inside the step graph short CTAs
    // interleave better with the sweep that runs beside them (62.2 vs 65.4 us per step).
    const int slots = sm_count() * (ri == 8 && rj == 8 ? 2 : ri * rj == 32 ? 3 : 4);
    int best_t = 1;
    long long best_cost = -1;
    for (int tpc = 1; tpc <= w.nblk_j; ++tpc) {
      const int grp = (w.nblk_j + tpc - 1) / tpc;
      const long long rounds = ((long long)w.nblk_i * grp + slots - 1) / slots;
      const long long cost = rounds * tpc;
      if (best_cost < 0 || cost < best_cost) {
        best_cost = cost;
        best_t = tpc;
      }
    }
    static int env_tpc = -1;
    if (env_tpc < 0) {
      const char *e = getenv("MACR_GRID_TPC");  // developer knob: column tiles per CTA
      env_tpc = e ? atoi(e) : 0;
    }
    if (env_tpc > 0) best_t = env_tpc < w.nblk_j ? env_tpc : w.nblk_j;
    w.grp_tiles = best_t;
    w.ngrp_j = (w.nblk_j + best_t - 1) / best_t;
  }
  const int tmax = w.tile_i > w.tile_j ? w.tile_i : w.tile_j;
  w.Bpad = (B + tmax - 1) / tmax * tmax;
  float *p = reinterpret_cast<float *>(base);
  const size_t band_r = (size_t)w.ngrp_j * w.Bpad, band_c = (size_t)w.nblk_i * w.Bpad;
  w.rowP = p;
  w.rowN = p + band_r;
  w.colP = p + 2 * band_r;
  w.colN = p + 2 * band_r + band_c;
  w.losspart = p + 2 * band_r + 2 * band_c;
  const size_t lp = ((size_t)w.nblk_i * w.ngrp_j + 3) & ~(size_t)3;
  w.litem = w.losspart + lp;
  w.luser = w.litem + w.Bpad;
  w.gA = w.luser + w.Bpad;
  w.gAN = w.gA + w.Bpad;
  w.gG = w.gAN + w.Bpad;
  w.item_gate_only = 0;
  w.part_bytes = (2 * band_r + 2 * band_c + lp) * sizeof(float);  // rowP..losspart: armed slots
  w.bytes = (2 * band_r + 2 * band_c + lp + 5 * (size_t)w.Bpad) * sizeof(float);
  return w;
}

// gates and branch losses of the batch from the three branch logits (model.py:204-205,213,215):
// a = sig(sp), an = sig(sn), g = sig(su); litem = -log(a+eps) - log(1-an+eps); luser likewise.
// (The trainers get the same values straight from gather_dots.)
__device__ __forceinline__ void gate_values(float sp, float sn, float su, float &a, float &an,
                                            float &g, float &litem, float &luser) {
  a = sigmoid_precise(sp);
  an = sigmoid_precise(sn);
  g = sigmoid_precise(su);
  const float ea = a + kBceEps, ean = (1.0f - an) + kBceEps;
  const float eg = g + kBceEps, eg1 = (1.0f - g) + kBceEps;
  litem = -logf(ea) - logf(ean);
  luser = -logf(eg) - logf(eg1);
}

__global__ void __launch_bounds__(256)
gate_kernel(const float *__restrict__ sp, const float *__restrict__ sn,
            const float *__restrict__ su, int B, GridWs ws) {
  const int b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= B) return;
  float a, an, g, li, lu;
  gate_values(sp[b], sn[b], su[b], a, an, g, li, lu);
  ws.gA[b] = a;
  ws.gAN[b] = an;
  ws.gG[b] = g;
  ws.litem[b] = li;
  ws.luser[b] = lu;
}

int launch_gates(const float *sp, const float *sn, const float *su, int B, const GridWs &ws,
                 cudaStream_t s) {
  gate_kernel<<<(B + 255) / 256, 256, 0, s>>>(sp, sn, su, B, ws);
  MACR_LAUNCH_CHECK();
  return MACR_OK;
}

template <int RI, int RJ, int MINB, bool kGrad>
static void launch_grid_t(const float *yp, const float *yn, int B, float alpha, float beta,
                          const GridWs &ws, const GridOut &out, cudaStream_t s, bool pdl) {
  // tiles, then the loss folder and (grad only) one folder CTA per row band and per column band
  const int folders = 1 + (kGrad ? ws.nblk_i + ws.nblk_j : 0);
  const int fold_rows = (folders + ws.ngrp_j - 1) / ws.ngrp_j;
  dim3 grid(ws.ngrp_j, ws.nblk_i + fold_rows);
  if (B % ws.tile_i == 0 && B % ws.tile_j == 0)
    launch_k(grid_bce_kernel<RI, RJ, MINB, false, kGrad>, grid, dim3(256), 0, s, pdl, yp, yn, B, alpha, beta, ws, out);
  else
    launch_k(grid_bce_kernel<RI, RJ, MINB, true, kGrad>, grid, dim3(256), 0, s, pdl, yp, yn, B, alpha, beta, ws, out);
}

// the partial-sum slots of `ws` (first ws.part_bytes bytes) must hold 0xff bytes before the first
// launch (the folders re-arm them);
// ws.gA / gAN / gG / litem / luser must hold the gates of the batch (launch_gates / gather_dots)
int launch_grid_bce(const float *yp, const float *yn, int B, const macr_hparams &hp,
                    const GridWs &ws, float *d_yp, float *d_yn, float *d_sp, float *d_sn,
                    float *d_su, int want_grad, const float *regsq, const StepState *st,
                    float *losses3, cudaStream_t s, bool pdl) {
  const float alpha = hp.alpha, beta = hp.beta;
  const GridOut out{d_yp, d_yn, d_sp, d_sn, d_su, regsq, hp, st, losses3};
  const int ri = ws.tile_i / 16, rj = ws.tile_j / 16;
#define MACR_GRID_CASE(RI_, RJ_, MINB_)                                                          \
  if (ri == RI_ && rj == RJ_) {                                                                   \
    if (want_grad) launch_grid_t<RI_, RJ_, MINB_, true>(yp, yn, B, alpha, beta, ws, out, s, pdl); \
    else launch_grid_t<RI_, RJ_, MINB_, false>(yp, yn, B, alpha, beta, ws, out, s, pdl);          \
  } else
  MACR_GRID_CASE(8, 8, 2)
  MACR_GRID_CASE(4, 8, 3)
  MACR_GRID_CASE(8, 4, 3)
  MACR_GRID_CASE(4, 4, 4)
  return fail(MACR_ERR_INVALID, "grid tile %dx%d not instantiated", ri, rj);
#undef MACR_GRID_CASE
  MACR_LAUNCH_CHECK();
  return MACR_OK;
}

// ---------------------------------------------------------------------------------------------
// `--train normalbce` (the README's baseline command; macr_mf/model.py:277-287): element-wise
//   mf = mean_b( -log(sig(yp_b) + 1e-9) - log(1 - sig(yn_b) + 1e-9) )
//   d mf / d yp_b = -s(1-s)/(s+1e-9)/B      d mf / d yn_b = t(1-t)/((1-t)+1e-9)/B
// One CTA: B scalars.  The gate gradients are zero in this graph (w, w_user are not part of it),
// so the row-gradient kernel runs unchanged.  Losses {loss, mf, reg, mf} go where the grid
// kernel's loss folder puts them.
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(1024)
plain_bce_kernel(const float *__restrict__ yp, const float *__restrict__ yn, int B,
                 macr_hparams hp, const float *__restrict__ regsq, const StepState *st,
                 float *losses_direct, float *__restrict__ d_yp, float *__restrict__ d_yn,
                 float *__restrict__ d_sp, float *__restrict__ d_sn, float *__restrict__ d_su) {
  __shared__ double sh[32][2];
  const float eps9 = 1e-9f, invB = 1.0f / (float)B;
  double a0 = 0, a1 = 0;
  for (int b = threadIdx.x; b < B; b += blockDim.x) {
    const float s = 1.0f / (1.0f + expf(-yp[b])), t = 1.0f / (1.0f + expf(-yn[b]));
    const float se = s + eps9, q = (1.0f - t) + eps9;
    a0 += (double)(-logf(se)) + (double)(-logf(q));
    a1 += regsq[b];
    d_yp[b] = -(s * (1.0f - s)) / se * invB;
    d_yn[b] = (t * (1.0f - t)) / q * invB;
    d_sp[b] = 0.f;
    d_sn[b] = 0.f;
    d_su[b] = 0.f;
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    a0 += __shfl_xor_sync(0xffffffffu, a0, o);
    a1 += __shfl_xor_sync(0xffffffffu, a1, o);
  }
  if ((threadIdx.x & 31) == 0) {
    sh[threadIdx.x >> 5][0] = a0;
    sh[threadIdx.x >> 5][1] = a1;
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    a0 = a1 = 0;
    for (int k = 0; k < (int)(blockDim.x >> 5); ++k) {
      a0 += sh[k][0];
      a1 += sh[k][1];
    }
    const float mf = (float)(a0 / (double)B);
    const float reg = hp.decay * ((float)(a1 * 0.5) / (float)hp.batch_size_flag);
    float *dst = st ? st->loss_base + st->step_idx * 4 : losses_direct;
    dst[0] = mf + reg;
    dst[1] = mf;
    dst[2] = reg;
    dst[3] = mf;
  }
}

int launch_plain_bce(const float *yp, const float *yn, int B, const macr_hparams &hp,
                     const float *regsq, const StepState *st, float *losses_direct, float *d_yp,
                     float *d_yn, float *d_sp, float *d_sn, float *d_su, cudaStream_t s) {
  plain_bce_kernel<<<1, 1024, 0, s>>>(yp, yn, B, hp, regsq, st, losses_direct, d_yp, d_yn, d_sp,
                                      d_sn, d_su);
  MACR_LAUNCH_CHECK();
  return MACR_OK;
}

// ---------------------------------------------------------------------------------------------
// K5a: batch plan -- group the batch positions that hit the same table row, deterministically.
// One CTA per table runs a stable LSD radix sort (8-bit digits) of the row ids in shared memory
// (n <= 16384 ids; a few microseconds, and it occupies 2 of the 148 SMs while the B x B grid
// runs on the rest).  A pass: every warp owns a contiguous chunk of the sequence and walks it 32
// keys at a time -- eight ballots group the lanes with the same digit, which gives each key its
// rank inside the warp's chunk and the warp's digit histogram in one sweep; a block scan over
// the (digit, warp) histogram turns ranks into destinations.  Stability keeps positions
// ascending inside a row -- the order TF's unsorted_segment_sum adds the duplicate slices in.
// A block scan over the segment heads then emits the compact plan:
//   uniq_rows[slot], seg_off[slot], seg_pos[k] (k = sorted index), *n_uniq, and one 16-byte
//   record per sorted index {rank << 16 | slot, row, position, first position of the segment}
//   -- everything the row-gradient kernel needs about an entry in one load.
// (array_ops.unique would list the rows in first-occurrence order; the order of the unique rows
// does not enter any result, only the order inside a segment does.)
// ---------------------------------------------------------------------------------------------
struct PlanTable {
  const int32_t *ids;
  int ids_off, n_ids, passes;  // passes = ceil(bits(table_rows - 1) / 8)
  PlanBufs out;
  uint32_t *bitmap;
};

constexpr int kPlanMaxRounds = 16;  // 16384 ids / 1024 threads

#ifdef MACR_PLAN_PROFILE  // developer instrumentation (dev/kbench.cu): phase time stamps of CTA 1
__device__ long long macr_plan_clk[32];
#define PLAN_STAMP(i) \
  do { if (threadIdx.x == 0 && blockIdx.x == gridDim.x - 1) macr_plan_clk[i] = clock64(); } while (0)
#else
#define PLAN_STAMP(i)
#endif

__global__ void __launch_bounds__(1024)
batch_plan_sort_kernel(PlanTable t0, PlanTable t1, const StepState *st, int B) {
  extern __shared__ __align__(16) unsigned char plan_smem[];
  __shared__ int s_scan[32];
  const PlanTable &t = blockIdx.x == 0 ? t0 : t1;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, n = t.n_ids;
  if (n == 0) return;
  PLAN_STAMP(0);
  const int32_t *ids = (st ? st->ids_base + st->step_idx * 3LL * B : t.ids) + t.ids_off;
  int npad = 1024;
  while (npad < n) npad <<= 1;
  uint32_t *keyA = reinterpret_cast<uint32_t *>(plan_smem), *keyB = keyA + npad;
  uint32_t(*hist)[33] = reinterpret_cast<uint32_t(*)[33]>(keyB + npad);  // [256][33]
  unsigned short *posA = reinterpret_cast<unsigned short *>(hist + 256), *posB = posA + npad;

  for (int e = tid; e < npad; e += 1024) {
    uint32_t r = 0xffffffffu;  // padding sorts last in every pass and stays last (stable)
    if (e < n) {
      r = (uint32_t)ids[e];
      if (t.bitmap) atomicOr(&t.bitmap[r >> 5], 1u << (r & 31));
    }
    keyA[e] = r;
    posA[e] = (unsigned short)e;
  }
  const int rounds = npad >> 10, chunk = npad >> 5;
  const unsigned lt_mask = (1u << lane) - 1u;
  for (int pass = 0; pass < t.passes; ++pass) {
    const int shift = 8 * pass;
    PLAN_STAMP(1 + 4 * pass);
    for (int c = tid; c < 256 * 33; c += 1024) (&hist[0][0])[c] = 0;
    __syncthreads();
    PLAN_STAMP(2 + 4 * pass);
    uint32_t key[kPlanMaxRounds];
    int rank[kPlanMaxRounds];
#pragma unroll
    for (int r = 0; r < kPlanMaxRounds; ++r) {
      if (r < rounds) {
        key[r] = keyA[warp * chunk + r * 32 + lane];
        const unsigned d = (key[r] >> shift) & 255u;
        // lanes holding the same digit: 8 ballots (MATCH.ANY is ~100x slower on this part)
        unsigned peers = 0xffffffffu;
#pragma unroll
        for (int bit = 0; bit < 8; ++bit) {
          const bool on = (d >> bit) & 1u;
          const unsigned vote = __ballot_sync(0xffffffffu, on);
          peers &= on ? vote : ~vote;
        }
        const uint32_t before = hist[d][warp];
        __syncwarp();
        if ((peers & lt_mask) == 0) hist[d][warp] = before + __popc(peers);
        __syncwarp();
        rank[r] = before + __popc(peers & lt_mask);
      }
    }
    __syncthreads();
    PLAN_STAMP(3 + 4 * pass);
    {  // exclusive scan of the 256 x 32 histogram in (digit, warp) order: 8 counters per thread
      const int d = tid >> 2, w0 = (tid & 3) * 8;
      uint32_t v[8], sum = 0;
#pragma unroll
      for (int q = 0; q < 8; ++q) {
        v[q] = hist[d][w0 + q];
        sum += v[q];
      }
      uint32_t incl = sum;
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) {
        const uint32_t x = __shfl_up_sync(0xffffffffu, incl, o);
        if (lane >= o) incl += x;
      }
      if (lane == 31) s_scan[warp] = (int)incl;
      __syncthreads();
      uint32_t base = 0;
#pragma unroll
      for (int w = 0; w < 32; ++w)
        if (w < warp) base += (uint32_t)s_scan[w];
      uint32_t run = base + incl - sum;
#pragma unroll
      for (int q = 0; q < 8; ++q) {
        hist[d][w0 + q] = run;
        run += v[q];
      }
    }
    __syncthreads();
    PLAN_STAMP(4 + 4 * pass);
#pragma unroll
    for (int r = 0; r < kPlanMaxRounds; ++r) {
      if (r < rounds) {
        const unsigned d = (key[r] >> shift) & 255u;
        const int dst = (int)hist[d][warp] + rank[r];
        keyB[dst] = key[r];
        posB[dst] = posA[warp * chunk + r * 32 + lane];
      }
    }
    __syncthreads();
    uint32_t *tk = keyA;
    keyA = keyB;
    keyB = tk;
    unsigned short *tp = posA;
    posA = posB;
    posB = tp;
  }
  PLAN_STAMP(20);
  // segment heads, walked in the same warp-striped order as the sort rounds (conflict-free
  // shared-memory reads, coalesced global stores): a ballot per round gives every sorted index
  // its segment number inside the warp's chunk and the start of its segment
  int run_heads = 0, run_last = -1;  // warp-uniform: heads so far in this chunk, last head index
  int lslot[kPlanMaxRounds], lstart[kPlanMaxRounds];
#pragma unroll
  for (int r = 0; r < kPlanMaxRounds; ++r) {
    if (r < rounds) {
      const int k = warp * chunk + r * 32 + lane;
      const bool head = k < n && (k == 0 || keyA[k] != keyA[k - 1]);
      const unsigned hb = __ballot_sync(0xffffffffu, head);
      const unsigned le = hb & (lt_mask | (1u << lane));
      lslot[r] = run_heads + __popc(le) - 1;  // -1: the segment started in an earlier chunk
      lstart[r] = le ? (k - lane + 31 - __clz(le)) : run_last;
      run_heads += __popc(hb);
      if (hb) run_last = k - lane + 31 - __clz(hb);
    }
  }
  __shared__ int s_max[32];
  __syncthreads();  // s_scan reuse
  if (lane == 0) {
    s_scan[warp] = run_heads;
    s_max[warp] = run_last;
  }
  __syncthreads();
  int base = 0, all = 0, bmax = -1;
#pragma unroll
  for (int w = 0; w < 32; ++w) {
    const int v = s_scan[w];
    if (w < warp) {
      base += v;
      bmax = max(bmax, s_max[w]);
    }
    all += v;
  }
  PLAN_STAMP(21);
#pragma unroll
  for (int r = 0; r < kPlanMaxRounds; ++r) {
    if (r < rounds) {
      const int k = warp * chunk + r * 32 + lane;
      if (k < n) {
        const int slot = base + lslot[r];
        const int start = lstart[r] >= 0 ? lstart[r] : bmax;
        t.out.seg_pos[k] = (int32_t)posA[k];
        t.out.rec[k] = make_int4((int)(((uint32_t)(k - start) << 16) | (uint32_t)slot), (int)keyA[k],
                                 (int)posA[k], (int)posA[start]);
        if (k == start) {
          t.out.uniq_rows[slot] = (int32_t)keyA[k];
          t.out.seg_off[slot] = k;
        }
      }
    }
  }
  if (tid == 0) {
    t.out.seg_off[all] = n;
    *t.out.n_uniq = all;
  }
  PLAN_STAMP(22);
}

// scratch behind PlanBufs: rec[n_ids] (int4), done[n_ids] (zero before the first use, self re-arming)
size_t plan_ws_bytes(int n_ids) { return sizeof(int32_t) * (5 * (size_t)n_ids + 16); }

static size_t plan_smem_bytes(int npad) { return (size_t)npad * 12 + 256 * 33 * 4; }

int plan_init() {  // opt in to the large dynamic shared memory once (outside any stream capture)
  static bool attr_set = false;
  if (!attr_set) {
    MACR_CUDA(cudaFuncSetAttribute(batch_plan_sort_kernel,
                                   cudaFuncAttributeMaxDynamicSharedMemorySize,
                                   (int)plan_smem_bytes(16384)));
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
  b.rec = reinterpret_cast<int4 *>(p);  // ws is 16-byte aligned (cudaMalloc / torch allocations)
  b.done = p + 4 * (size_t)n_ids;
  return b;
}

static int radix_passes(int64_t table_rows) {
  int bits = 1;
  while (bits < 32 && ((int64_t)1 << bits) < table_rows) ++bits;
  return (bits + 7) / 8;
}

int launch_batch_plan2(const int32_t *ids0, const StepState *st, int ids0_off, int n_ids0,
                       int64_t rows0, PlanBufs out0, uint32_t *bitmap0, const int32_t *ids1,
                       int ids1_off, int n_ids1, int64_t rows1, PlanBufs out1, uint32_t *bitmap1,
                       cudaStream_t s) {
  const int nmax = n_ids0 > n_ids1 ? n_ids0 : n_ids1;
  MACR_CHECK_ARG(nmax <= 16384, "batch plan supports at most 16384 ids per table (batch <= 8192)");
  int rci = plan_init();
  if (rci) return rci;
  PlanTable t0{ids0, ids0_off, n_ids0, radix_passes(rows0), out0, bitmap0};
  PlanTable t1{ids1, ids1_off, n_ids1, radix_passes(rows1), out1, bitmap1};
  int npad = 1024;
  while (npad < nmax) npad <<= 1;
  const int B = st ? n_ids0 : 0;  // trainer convention: table 0 = users (B ids)
  batch_plan_sort_kernel<<<n_ids1 > 0 ? 2 : 1, 1024, plan_smem_bytes(npad), s>>>(t0, t1, st, B);
  MACR_LAUNCH_CHECK();
  return MACR_OK;
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

// ---------------------------------------------------------------------------------------------
// K5b: Adam.  The sweep is the HBM-bound part of the step: 24 B per table element
// (read + write of var, m, v), 128-bit streaming accesses, UNROLL independent float4 triples in
// flight per thread.  Persistent grid: every CTA owns one contiguous, equally sized slice of the
// (table 0 | table 1) element space, so there is no wave-quantisation tail.  Rows touched by the
// batch are skipped here (bitmap) and handled by the row-gradient kernel once their gradient is
// known, so the sweep overlaps the B x B grid.
// ---------------------------------------------------------------------------------------------
struct SweepTable {
  float4 *var, *m, *v;
  long long n4;  // rows * 16
  const uint32_t *bitmap;
};

constexpr int kSweepThreads = 256;

template <int UNROLL>
__global__ void __launch_bounds__(kSweepThreads, 3)
adam_sweep_kernel(SweepTable t0, SweepTable t1, long long per_cta, float lr_or_lrt,
                  const StepState *st, float b1, float b2, float eps) {
  const float lr_t = step_lr_t(st, lr_or_lrt);
  const long long total = t0.n4 + t1.n4;
  const long long beg = (long long)blockIdx.x * per_cta;
  const long long end = min(total, beg + per_cta);
  for (long long base = beg + threadIdx.x; base < end; base += (long long)kSweepThreads * UNROLL) {
    float4 x[UNROLL], mm[UNROLL], vv[UNROLL];
    long long off[UNROLL];
    bool live[UNROLL], second[UNROLL];
#pragma unroll
    for (int k = 0; k < UNROLL; ++k) {
      long long e = base + (long long)k * kSweepThreads;
      live[k] = e < end;
      second[k] = e >= t0.n4;
      if (second[k]) e -= t0.n4;
      off[k] = e;
      if (live[k]) {
        const uint32_t *bm = second[k] ? t1.bitmap : t0.bitmap;
        const long long row = e >> 4;
        if (bm && ((bm[row >> 5] >> (row & 31)) & 1u)) live[k] = false;
      }
    }
#pragma unroll
    for (int k = 0; k < UNROLL; ++k)
      if (live[k]) {
        const SweepTable &t = second[k] ? t1 : t0;
        x[k] = ld_stream(t.var + off[k]);
        mm[k] = ld_stream(t.m + off[k]);
        vv[k] = ld_stream(t.v + off[k]);
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
        const SweepTable &t = second[k] ? t1 : t0;
        st_stream(t.var + off[k], x[k]);
        st_stream(t.m + off[k], mm[k]);
        st_stream(t.v + off[k], vv[k]);
      }
  }
}

// Resident sweep CTAs per SM.  3 fill the register file (85 regs x 256 threads x 3) and are the
// fastest for L2-sized tables (gowalla: a 21 us sweep beside a 25 us grid).  For tables of GBs the
// sweep runs for milliseconds and 3 CTAs/SM leave no room for the step's other kernels (gather,
// B x B grid, row gradients), which then queue behind it: measured on the 10M x 1M tables, B=8192,
// 3.54 ms/step with 3 CTAs/SM vs 3.02 ms with 2 (profiles/r2a_c5_probe.txt).
static int sweep_ctas_per_sm(long long total_f4) {
  static int v = -1;
  if (v < 0) {
    const char *e = getenv("MACR_SWEEP_CTAS_PER_SM");  // tuning knob, 0 / unset = by size
    v = e ? atoi(e) : 0;
    if (v < 0 || v > 8) v = 0;
  }
  if (v) return v;
  return total_f4 * 48 >= (256LL << 20) ? 2 : 3;  // >= 256 MiB of var + m + v
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
  long long ctas = (long long)sm_count() * sweep_ctas_per_sm(total);
  const long long min_per = (long long)kSweepThreads * UNROLL;
  if (ctas * min_per > total) ctas = (total + min_per - 1) / min_per;
  long long per = (total + ctas - 1) / ctas;
  per = (per + kSweepThreads - 1) / kSweepThreads * kSweepThreads;  // whole 4 KB lines per warp row
  ctas = (total + per - 1) / per;
  adam_sweep_kernel<UNROLL><<<(int)ctas, kSweepThreads, 0, s>>>(t0, t1, per, lr_t, st, b1, b2, eps);
  MACR_LAUNCH_CHECK();
  return MACR_OK;
}

// ---------------------------------------------------------------------------------------------
// Row gradients of the unique touched rows, summed per table row (the IndexedSlices dedup), and
// -- for MF -- TF's sparse Adam formula applied to each such row as soon as its sum is complete.
//   user row r :  sum_b  dyp_b*pe_b + dyn_b*ne_b + dsu_b*w_user (+ lam*ue_b)
//   item row r :  sum_q  dy_q*ue_b + ds_q*w (+ lam*ie_q),  q<B: pos role, q>=B: neg role
// Every embedding row is read from the step's snapshot [3][B][64] (users | pos | neg, written by
// gather_dots), indexed by batch position -- so updating a table row in place cannot disturb the
// gradient of another row, and no id indirection is left in this kernel.
// Popular items own segments hundreds of positions long, so the unit of work is <= 32 consecutive
// entries of one segment: the warp of sorted index k works only if (k - seg_off) % 32 == 0.  Its
// lanes fetch the 32 entries' metadata with one coalesced access each, then the row loads (lane l
// owns dims 2l, 2l+1) are issued 4 entries at a time.  Segments longer than one unit leave
// per-unit partials in `unit_part`; the last unit to arrive (per-segment ticket) adds them in
// unit order, so the sum is order-deterministic.
// Extra CTAs at the start of the grid reduce grad(w) = sum_b dsp_b*pe_b + dsn_b*ne_b and
// grad(w_user) = sum_b dsu_b*ue_b into per-CTA partials (fixed composition, fixed order).
// With tail.fused the last CTA of the whole grid (arrival ticket) also runs the step tail:
// ApplyAdam on w / w_user, the loss reduction and the step-state advance.
// ---------------------------------------------------------------------------------------------
constexpr int kRowWarps = 8;
constexpr int kUnit = 32;
constexpr int kWgradPerCta = 64;  // batch positions per w-gradient CTA

// ApplyAdam on w / w_user from the per-CTA gradient partials and the step-state advance
// (adam.py _finish: beta powers *= beta).  Runs in ONE CTA of NT threads.  (The loss scalars of
// the step are produced by the grid kernel's loss folder.)
template <int NT>
__device__ void step_tail_body(float *w, float *mw, float *vw, float *wu, float *mwu, float *vwu,
                               const float *gw_part, const float *gwu_part, int n_part,
                               const macr_hparams &hp, StepState *st, int train,
                               float (*shg)[kD]) {
  constexpr int GR = NT / (2 * kD);  // partial groups per vector
  const int tid = threadIdx.x;
  const float lr_t = step_lr_t(st, hp.lr);
  // bit 0: training step; bits 1..2: vectors without a gradient in this graph (minimize() skips
  // them: neither the variable nor its slots are written)
  const int frozen = train >> 1;
  train &= 1;
  if (train) {
    const int k = tid & 63, which = (tid >> 6) & 1, grp = tid / (2 * kD);
    const float *src = which ? gwu_part : gw_part;
    // partials q = grp, grp + GR, ... summed in that order; the loads go out 16 at a time (the
    // whole step waits for this CTA: one round trip per partial was ~5 us at B = 4096)
    float a = 0.f;
    for (int q0 = grp; q0 < n_part; q0 += 16 * GR) {
      float t[16];
#pragma unroll
      for (int j = 0; j < 16; ++j) {
        const int q = q0 + j * GR;
        t[j] = q < n_part ? __ldcg(src + (long long)q * kD + k) : 0.f;
      }
#pragma unroll
      for (int j = 0; j < 16; ++j)
        if (q0 + j * GR < n_part) a += t[j];
    }
    shg[which * GR + grp][k] = a;
  }
  __syncthreads();
  if (train && tid < 2 * kD && !((frozen >> (tid >> 6)) & 1)) {
    const int k = tid & 63, which = tid >> 6;
    float g = 0.f;
#pragma unroll
    for (int q = 0; q < GR; ++q) g += shg[which * GR + q][k];
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
    if (train) {
      st->b1p = __fmul_rn(st->b1p, hp.beta1);
      st->b2p = __fmul_rn(st->b2p, hp.beta2);
      st->t += 1;
    }
    st->step_idx += 1;
  }
}

__global__ void __launch_bounds__(kRowWarps * 32)
row_grads_kernel(const float *__restrict__ snap, const float *__restrict__ w,
                 const float *__restrict__ wu, int B, const float *__restrict__ d_yp,
                 const float *__restrict__ d_yn, const float *__restrict__ d_sp,
                 const float *__restrict__ d_sn, const float *__restrict__ d_su, float lam,
                 PlanBufs planU, PlanBufs planI, float *__restrict__ gU, float *__restrict__ gI,
                 float *unit_part, int pos_ctas, float *__restrict__ gw_part,
                 float *__restrict__ gwu_part, AdamTabs tabs, TailArgs tail) {
  __shared__ float2 sW[kRowWarps][32], sWU[kRowWarps][32];
  __shared__ float sTailG[2 * (kRowWarps * 32 / (2 * kD))][kD];
  __shared__ int sLast;
  const int lane = threadIdx.x & 31, wl = threadIdx.x >> 5;
  const float *snapU = snap, *snapP = snap + (long long)B * kD, *snapN = snap + 2LL * B * kD;
  pdl_wait();  // the grid's folders have written d_yp .. d_su (and everything before them is done)

  // the partial CTAs come FIRST in the grid: dispatched ahead of the row CTAs they are long done
  // when the last row CTA retires and the step tail sums their output
  const int w_ctas = (int)gridDim.x - pos_ctas;
  if ((int)blockIdx.x < w_ctas) {  // ---- grad(w), grad(w_user) partials ----
    const int cta = blockIdx.x;
    float2 aw = make_float2(0.f, 0.f), awu = make_float2(0.f, 0.f);
    const int b0 = cta * kWgradPerCta + wl * (kWgradPerCta / kRowWarps);
#pragma unroll 4
    for (int k = 0; k < kWgradPerCta / kRowWarps; ++k) {
      const int b = b0 + k;
      if (b < B) {
        const float dsp = d_sp[b], dsn = d_sn[b], dsu = d_su[b];
        const float2 pe = reinterpret_cast<const float2 *>(snapP + (long long)b * kD)[lane];
        const float2 ne = reinterpret_cast<const float2 *>(snapN + (long long)b * kD)[lane];
        const float2 ue = reinterpret_cast<const float2 *>(snapU + (long long)b * kD)[lane];
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
  } else {
    const int wid = ((int)blockIdx.x - w_ctas) * kRowWarps + wl;  // one warp per sorted index: users, then items
    const bool item = wid >= B;
    const int k = item ? wid - B : wid;
    const PlanBufs &pl = item ? planI : planU;
    const int n_ids = item ? 2 * B : B;
    // level 1: one 16-byte record per lane (entry k + lane), plus the record after the unit
    int4 rec = make_int4(0x10000, 0, 0, 0);  // rank 1: not a unit start
    if (wid < 3 * B && k + lane < n_ids) rec = pl.rec[k + lane];
    int next_info = -1;
    if (lane == 0 && wid < 3 * B && k + kUnit < n_ids) next_info = pl.rec[k + kUnit].x;
    const unsigned info = (unsigned)__shfl_sync(0xffffffffu, rec.x, 0);
    const int rk = (int)(info >> 16), slot = (int)(info & 0xffffu);
    if (rk % kUnit == 0) {  // this warp owns one unit of the segment
      const int s0 = k - rk;
      const bool mine = (rec.x & 0xffff) == slot && k + lane < n_ids;
      const int cnt = __popc(__ballot_sync(0xffffffffu, mine));  // sorted: a prefix of the lanes
      const long long r = __shfl_sync(0xffffffffu, rec.y, 0);
      const int q0 = __shfl_sync(0xffffffffu, rec.w, 0);
      next_info = __shfl_sync(0xffffffffu, next_info, 0);
      const bool multi = rk > 0 || (cnt == kUnit && next_info >= 0 && (next_info & 0xffff) == slot);
      const int total = multi ? pl.seg_off[slot + 1] - s0 : cnt;
      const int qq_spec = rec.z;
      const float2 bias = reinterpret_cast<const float2 *>(item ? w : wu)[lane];
      // metadata of this unit's entries, one entry per lane
      int ra = 0;  // snapshot row of the partner (item unit: user row b; user unit: position b)
      float ca = 0.f, cb = 0.f, cs = 0.f;
      if (lane < cnt) {
        const int qq = qq_spec;
        if (!item) {
          ra = qq;
          ca = d_yp[qq];
          cb = d_yn[qq];
          cs = d_su[qq];
        } else {
          const bool is_pos = qq < B;
          ra = is_pos ? qq : qq - B;
          ca = is_pos ? d_yp[ra] : d_yn[ra];
          cs = is_pos ? d_sp[ra] : d_sn[ra];
        }
      }
      // the row's Adam state is requested now and consumed after the gradient is complete
      float2 ax = make_float2(0.f, 0.f), am = ax, av = ax;
      float2 *pv = nullptr, *pm = nullptr, *pvv = nullptr;
      if (tabs.U != nullptr && total <= kUnit) {
        pv = reinterpret_cast<float2 *>((item ? tabs.I : tabs.U) + r * kD) + lane;
        pm = reinterpret_cast<float2 *>((item ? tabs.mI : tabs.mU) + r * kD) + lane;
        pvv = reinterpret_cast<float2 *>((item ? tabs.vI : tabs.vU) + r * kD) + lane;
        ax = *pv;
        am = *pm;
        av = *pvv;
      }
      float lx = 0.f, ly = 0.f;
      if (lam != 0.f) {  // L2 slice of the row itself: any position of the segment holds it
        const float2 raw = reinterpret_cast<const float2 *>(
            (item ? snapP : snapU) + (long long)q0 * kD)[lane];
        lx = lam * raw.x;
        ly = lam * raw.y;
      }
      float2 g = make_float2(0.f, 0.f);
      for (int e0 = 0; e0 < cnt; e0 += 4) {
        float2 va[4], vb[4];
        float fa[4], fb[4], fs[4];
        // partner rows first: their addresses need the plan record only, so these loads go out
        // together with the coefficient loads above instead of behind them (a shuffle of `ca`
        // waits for its load)
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          const int e = min(e0 + q, cnt - 1);
          const int ia = __shfl_sync(0xffffffffu, ra, e);
          if (item) {
            va[q] = reinterpret_cast<const float2 *>(snapU + (long long)ia * kD)[lane];
            vb[q] = make_float2(0.f, 0.f);
          } else {
            va[q] = reinterpret_cast<const float2 *>(snapP + (long long)ia * kD)[lane];
            vb[q] = reinterpret_cast<const float2 *>(snapN + (long long)ia * kD)[lane];
          }
        }
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          const int e = min(e0 + q, cnt - 1);
          fa[q] = __shfl_sync(0xffffffffu, ca, e);
          fb[q] = __shfl_sync(0xffffffffu, cb, e);
          fs[q] = __shfl_sync(0xffffffffu, cs, e);
        }
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          if (e0 + q < cnt) {
            float x = fa[q] * va[q].x + fb[q] * vb[q].x + fs[q] * bias.x;
            float y = fa[q] * va[q].y + fb[q] * vb[q].y + fs[q] * bias.y;
            if (lam != 0.f) {
              x += lx;
              y += ly;
            }
            g.x += x;
            g.y += y;
          }
        }
      }
      bool final_here = total <= kUnit;
      if (!final_here) {
        // multi-unit segment: publish the partial, the last unit to arrive folds them in order
        float *mypart = unit_part + ((long long)(item ? B : 0) + k) * kD;
        __stcg(reinterpret_cast<float2 *>(mypart) + lane, g);
        __syncwarp();
        int ticket = 0;
        if (lane == 0) {
          __threadfence();
          ticket = atomicAdd(&pl.done[slot], 1);
        }
        ticket = __shfl_sync(0xffffffffu, ticket, 0);
        const int n_units = (total + kUnit - 1) / kUnit;
        if (ticket == n_units - 1) {
          if (lane == 0) {
            __threadfence();
            pl.done[slot] = 0;  // re-armed for the next step
          }
          __syncwarp();
          g = make_float2(0.f, 0.f);
          for (int un = 0; un < n_units; ++un) {
            const float2 v = __ldcg(reinterpret_cast<const float2 *>(
                                        unit_part + ((long long)(item ? B : 0) + s0 + un * kUnit) * kD) + lane);
            g.x += v.x;
            g.y += v.y;
          }
          final_here = true;
          if (tabs.U != nullptr) {
            pv = reinterpret_cast<float2 *>((item ? tabs.I : tabs.U) + r * kD) + lane;
            pm = reinterpret_cast<float2 *>((item ? tabs.mI : tabs.mU) + r * kD) + lane;
            pvv = reinterpret_cast<float2 *>((item ? tabs.vI : tabs.vU) + r * kD) + lane;
            ax = *pv;
            am = *pm;
            av = *pvv;
          }
        }
      }
      if (final_here) {
        if (tabs.U == nullptr) {
          reinterpret_cast<float2 *>((item ? gI : gU) + (long long)slot * kD)[lane] = g;
        } else {  // adam.py _apply_sparse_shared on this row
          const float lr_t = step_lr_t(tabs.st, tabs.lr);
          const float omb1 = __fsub_rn(1.0f, tabs.b1), omb2 = __fsub_rn(1.0f, tabs.b2);
          adam_with_grad(ax.x, am.x, av.x, g.x, lr_t, tabs.b1, tabs.b2, omb1, omb2, tabs.eps);
          adam_with_grad(ax.y, am.y, av.y, g.y, lr_t, tabs.b1, tabs.b2, omb1, omb2, tabs.eps);
          *pv = ax;
          *pm = am;
          *pvv = av;
          uint32_t *bm = item ? tabs.bmI : tabs.bmU;
          if (bm && lane == 0) atomicAnd(&bm[r >> 5], ~(1u << (r & 31)));
        }
      }
    }
  }
  if (!tail.fused) return;
  // ---- last CTA of the grid: the step tail ----
  __syncthreads();
  if (threadIdx.x == 0) {
    __threadfence();  // cumulative over the CTA's writes ordered by the barrier above
    const bool last = atomicAdd(tail.ticket, 1u) == gridDim.x - 1;
    if (last) {
      __threadfence();
      *tail.ticket = 0;  // re-armed for the next step
    }
    sLast = last;
  }
  __syncthreads();
  if (!sLast) return;
  step_tail_body<kRowWarps * 32>(tail.w, tail.mw, tail.vw, tail.wu, tail.mwu, tail.vwu, gw_part,
                                 gwu_part, gridDim.x - pos_ctas, tail.hp, tail.st,
                                 1 | (tail.frozen << 1), sTailG);
}

int row_grads_max_parts(int B) { return (B + kWgradPerCta - 1) / kWgradPerCta; }

// tabs == nullptr: summed rows to gU / gI; tail == nullptr: no fused step tail
int launch_row_grads(const float *snap, const float *w, const float *wu, int B, const float *d_yp,
                     const float *d_yn, const float *d_sp, const float *d_sn, const float *d_su,
                     float lam, PlanBufs planU, PlanBufs planI, float *gU, float *gI,
                     float *unit_part, float *gw_part, float *gwu_part, int *n_part,
                     const AdamTabs *tabs, const TailArgs *tail, cudaStream_t s, bool pdl) {
  const int pos_ctas = (3 * B + kRowWarps - 1) / kRowWarps;
  const int w_ctas = row_grads_max_parts(B);
  AdamTabs tb{};
  if (tabs) tb = *tabs;
  TailArgs tl{};
  if (tail) tl = *tail;
  launch_k(row_grads_kernel, dim3(pos_ctas + w_ctas), dim3(kRowWarps * 32), 0, s, pdl,
           snap, w, wu, B, d_yp, d_yn, d_sp, d_sn, d_su, lam, planU, planI, gU, gI, unit_part, pos_ctas,
           gw_part, gwu_part, tb, tl);
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

// ---------------------------------------------------------------------------------------------
// stand-alone step tail (one CTA): LightGCN steps (the dense Adam kernels read the step state, so
// the tail cannot ride on the row-gradient kernel) and loss-only steps.
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(1024)
step_tail_kernel(float *w, float *mw, float *vw, float *wu, float *mwu, float *vwu,
                 const float *__restrict__ gw_part, const float *__restrict__ gwu_part, int n_part,
                 macr_hparams hp, StepState *st, int train) {
  __shared__ float shg[16][kD];
  step_tail_body<1024>(w, mw, vw, wu, mwu, vwu, gw_part, gwu_part, n_part, hp, st, train, shg);
}

int launch_step_tail(float *w, float *mw, float *vw, float *wu, float *mwu, float *vwu,
                     const float *gw_part, const float *gwu_part, int n_part,
                     const macr_hparams &hp, StepState *st, int train, cudaStream_t s) {
  step_tail_kernel<<<1, 1024, 0, s>>>(w, mw, vw, wu, mwu, vwu, gw_part, gwu_part, n_part, hp, st,
                                      train);
  MACR_LAUNCH_CHECK();
  return MACR_OK;
}

}  // namespace macr
