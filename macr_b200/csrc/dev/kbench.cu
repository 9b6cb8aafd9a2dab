// kbench.cu -- developer micro-harness (not part of the library): times single kernels of the
// training step in isolation with CUDA events, L2-warm and L2-cold, on bench-shaped inputs.
//   nvcc -O3 -std=c++17 -lineinfo -gencode arch=compute_100a,code=sm_100a -o kbench dev/kbench.cu \
//        api.o train_kernels.o   (see Makefile target `kbench`)
#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <random>
#include <vector>

#include "../train_kernels.cuh"
#ifdef MACR_PLAN_PROFILE
namespace macr { extern __device__ long long macr_plan_clk[32]; }
#endif

using namespace macr;

#define CK(x)                                                                      \
  do {                                                                             \
    cudaError_t e = (x);                                                           \
    if (e != cudaSuccess) {                                                        \
      printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e), __FILE__, __LINE__); \
      exit(1);                                                                     \
    }                                                                              \
  } while (0)

template <class F>
static float time_us(F f, int iters, void *flush, size_t flush_bytes) {
  cudaEvent_t a, b;
  CK(cudaEventCreate(&a));
  CK(cudaEventCreate(&b));
  for (int i = 0; i < 3; ++i) f();
  CK(cudaDeviceSynchronize());
  float total = 0;
  for (int i = 0; i < iters; ++i) {
    if (flush) CK(cudaMemsetAsync(flush, 0, flush_bytes));
    CK(cudaEventRecord(a));
    f();
    CK(cudaEventRecord(b));
    CK(cudaEventSynchronize(b));
    float ms;
    CK(cudaEventElapsedTime(&ms, a, b));
    total += ms;
  }
  return 1e3f * total / iters;
}

int main(int argc, char **argv) {
  const int B = argc > 1 ? atoi(argv[1]) : 4096;
  const int64_t NU = 29858, NI = 40981;
  std::mt19937 rng(12345);
  std::vector<int32_t> ids(3 * B);
  {
    std::vector<int32_t> perm(NU);
    for (int i = 0; i < NU; ++i) perm[i] = i;
    std::shuffle(perm.begin(), perm.end(), rng);
    for (int b = 0; b < B; ++b) ids[b] = perm[b % NU];
    std::vector<double> cdf(NI);
    double s = 0;
    for (int i = 0; i < NI; ++i) cdf[i] = (s += 1.0 / (i + 1));
    std::uniform_real_distribution<double> U01(0, 1);
    for (int b = 0; b < B; ++b) {
      ids[B + b] = (int32_t)(std::lower_bound(cdf.begin(), cdf.end(), U01(rng) * s) - cdf.begin());
      ids[2 * B + b] = (int32_t)(rng() % NI);
    }
  }
  int32_t *d_ids;
  CK(cudaMalloc(&d_ids, sizeof(int32_t) * 3 * B));
  CK(cudaMemcpy(d_ids, ids.data(), sizeof(int32_t) * 3 * B, cudaMemcpyHostToDevice));
  void *flush;
  const size_t fb = 256u << 20;
  CK(cudaMalloc(&flush, fb));

  // ---- plan ----
  int32_t *pm;
  const size_t pwords = (size_t)B * 9 + 64;
  CK(cudaMalloc(&pm, sizeof(int32_t) * pwords + plan_ws_bytes(B) + plan_ws_bytes(2 * B)));
  CK(cudaMemset(pm, 0, sizeof(int32_t) * pwords + plan_ws_bytes(B) + plan_ws_bytes(2 * B)));
  int32_t *p = pm;
  int32_t *uU = p; p += B; int32_t *oU = p; p += B + 1; int32_t *sU = p; p += B; int32_t *nU = p; p += 1;
  int32_t *uI = p; p += 2 * B; int32_t *oI = p; p += 2 * B + 1; int32_t *sI = p; p += 2 * B; int32_t *nI = p; p += 1;
  p = pm + pwords;
  PlanBufs planU = plan_carve(uU, oU, sU, nU, p, B);
  PlanBufs planI = plan_carve(uI, oI, sI, nI, (char *)p + plan_ws_bytes(B), 2 * B);
  auto plan = [&]() {
    launch_batch_plan2(d_ids, nullptr, 0, B, NU, planU, nullptr, d_ids, B, 2 * B, NI, planI, nullptr, 0);
  };
  printf("B=%d\n", B);
  printf("plan(users+items)   warm %7.2f us   cold %7.2f us\n", time_us(plan, 20, nullptr, 0),
         time_us(plan, 20, flush, fb));
  auto plan_u = [&]() {
    PlanBufs none{};
    launch_batch_plan2(d_ids, nullptr, 0, B, NU, planU, nullptr, nullptr, 0, 0, 1, none, nullptr, 0);
  };
  printf("plan(users only)    warm %7.2f us\n", time_us(plan_u, 20, nullptr, 0));
#ifdef MACR_PLAN_PROFILE
  {
    plan();
    CK(cudaDeviceSynchronize());
    long long clk[32];
    CK(cudaMemcpyFromSymbol(clk, macr::macr_plan_clk, sizeof(clk)));
    printf("plan phases (cycles since start, items CTA): ");
    for (int i : {1, 2, 3, 4, 5, 6, 7, 8, 20, 21, 22}) printf("[%d]=%lld ", i, clk[i] - clk[0]);
    printf("\n");
  }
#endif

  // ---- tables ----
  const size_t eU = (size_t)NU * kD, eI = (size_t)NI * kD;
  float *U, *mU, *vU, *I, *mI, *vI, *w, *wu;
  std::vector<float> h(eU + eI);
  std::uniform_real_distribution<float> Ur(-0.05f, 0.05f);
  for (auto &x : h) x = Ur(rng);
  CK(cudaMalloc(&U, 4 * eU)); CK(cudaMalloc(&mU, 4 * eU)); CK(cudaMalloc(&vU, 4 * eU));
  CK(cudaMalloc(&I, 4 * eI)); CK(cudaMalloc(&mI, 4 * eI)); CK(cudaMalloc(&vI, 4 * eI));
  CK(cudaMalloc(&w, 4 * kD)); CK(cudaMalloc(&wu, 4 * kD));
  CK(cudaMemcpy(U, h.data(), 4 * eU, cudaMemcpyHostToDevice));
  CK(cudaMemcpy(I, h.data() + eU, 4 * eI, cudaMemcpyHostToDevice));
  CK(cudaMemcpy(mU, h.data(), 4 * eU, cudaMemcpyHostToDevice));
  CK(cudaMemcpy(mI, h.data(), 4 * eI, cudaMemcpyHostToDevice));
  CK(cudaMemset(vU, 0x30, 4 * eU)); CK(cudaMemset(vI, 0x30, 4 * eI));  // tiny positive floats
  CK(cudaMemcpy(w, h.data(), 4 * kD, cudaMemcpyHostToDevice));
  CK(cudaMemcpy(wu, h.data() + 64, 4 * kD, cudaMemcpyHostToDevice));
  float *sc, *snap;
  CK(cudaMalloc(&sc, 4 * 11 * (size_t)B));
  CK(cudaMalloc(&snap, 4 * 3 * (size_t)B * kD));
  float *yp = sc, *yn = sc + B, *sp = sc + 2 * B, *sn = sc + 3 * B, *su = sc + 4 * B, *rq = sc + 5 * B,
        *dyp = sc + 6 * B, *dyn = sc + 7 * B, *dsp = sc + 8 * B, *dsn = sc + 9 * B, *dsu = sc + 10 * B;
  void *gws;
  GridWs g0 = grid_ws_layout(B, nullptr);
  CK(cudaMalloc(&gws, g0.bytes));
  CK(cudaMemset(gws, 0xff, g0.bytes));
  GridWs g = grid_ws_layout(B, gws);
  auto gather = [&]() {
    launch_gather_dots(U, I, U, I, w, wu, d_ids, d_ids + B, d_ids + 2 * B, nullptr, B, yp, yn, sp, sn,
                       su, rq, snap, &g, 0);
  };
  printf("gather_dots         warm %7.2f us   cold %7.2f us\n", time_us(gather, 20, nullptr, 0),
         time_us(gather, 20, flush, fb));
  // ---- grid ----
  macr_hparams hpk{};
  hpk.alpha = 1e-2f; hpk.beta = 1e-3f; hpk.batch_size_flag = 1;
  float *l3;
  CK(cudaMalloc(&l3, 64));
  auto grid = [&]() { launch_grid_bce(yp, yn, B, hpk, g, dyp, dyn, dsp, dsn, dsu, 1, nullptr, nullptr, l3, 0); };
  printf("grid_bce (grad)     warm %7.2f us   cold %7.2f us\n", time_us(grid, 20, nullptr, 0),
         time_us(grid, 20, flush, fb));
  auto grid0 = [&]() { launch_grid_bce(yp, yn, B, hpk, g, dyp, dyn, dsp, dsn, dsu, 0, nullptr, nullptr, l3, 0); };
  printf("grid_bce (loss)     warm %7.2f us\n", time_us(grid0, 20, nullptr, 0));
  // ---- sweep ----
  auto sweep = [&]() {
    launch_adam_sweep2(U, mU, vU, NU, nullptr, I, mI, vI, NI, nullptr, 1e-4f, nullptr, 0.9f, 0.999f, 1e-8f, 0);
  };
  printf("adam_sweep          warm %7.2f us   cold %7.2f us  (%.1f MB algorithmic)\n",
         time_us(sweep, 20, nullptr, 0), time_us(sweep, 20, flush, fb), 24.0 * (eU + eI) / 1e6);
  // ---- row grads + Adam (no tail) ----
  float *gU, *gI, *unit_part, *gwp, *gwup;
  CK(cudaMalloc(&gU, 4 * (size_t)B * kD)); CK(cudaMalloc(&gI, 8 * (size_t)B * kD));
  CK(cudaMalloc(&unit_part, 12 * (size_t)B * kD));
  CK(cudaMalloc(&gwp, 4 * kD * 256)); CK(cudaMalloc(&gwup, 4 * kD * 256));
  plan();
  AdamTabs tabs{U, mU, vU, I, mI, vI, nullptr, nullptr, 0.9f, 0.999f, 1e-8f, 1e-4f, nullptr};
  auto rowg = [&]() {
    launch_row_grads(snap, w, wu, B, dyp, dyn, dsp, dsn, dsu, 1e-9f, planU, planI, gU, gI, unit_part, gwp,
                     gwup, nullptr, &tabs, nullptr, 0);
  };
  printf("row_grads+adam      warm %7.2f us   cold %7.2f us\n", time_us(rowg, 20, nullptr, 0),
         time_us(rowg, 20, flush, fb));
  auto rowg2 = [&]() {
    launch_row_grads(snap, w, wu, B, dyp, dyn, dsp, dsn, dsu, 1e-9f, planU, planI, gU, gI, unit_part, gwp,
                     gwup, nullptr, nullptr, nullptr, 0);
  };
  printf("row_grads (no adam) warm %7.2f us\n", time_us(rowg2, 20, nullptr, 0));
  auto empty = [&]() { launch_mark_touched(nullptr, d_ids, 1, (uint32_t *)gU, (uint32_t *)gI, 0); };
  printf("tiny kernel         warm %7.2f us\n", time_us(empty, 20, nullptr, 0));
  CK(cudaDeviceSynchronize());
  return 0;
}
