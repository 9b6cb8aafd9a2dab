// common.cuh -- shared helpers for libmacr_b200 (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <string.h>

#include "../../include/macr_b200.h"

namespace macr {

// thread-local last-error message behind macr_last_error()
char *err_buf();
int fail(int code, const char *fmt, ...);

#define MACR_CHECK_ARG(cond, ...)                                   \
  do {                                                              \
    if (!(cond)) return ::macr::fail(MACR_ERR_INVALID, __VA_ARGS__); \
  } while (0)

#define MACR_CUDA(call)                                                                   \
  do {                                                                                    \
    cudaError_t e__ = (call);                                                             \
    if (e__ != cudaSuccess)                                                               \
      return ::macr::fail(MACR_ERR_CUDA, "%s:%d %s: %s", __FILE__, __LINE__, #call,       \
                          cudaGetErrorString(e__));                                       \
  } while (0)

#define MACR_LAUNCH_CHECK() MACR_CUDA(cudaGetLastError())

inline cudaStream_t as_stream(macr_stream_t s) { return reinterpret_cast<cudaStream_t>(s); }

int sm_count();

// Programmatic dependent launch (MACR_PDL=1): a kernel launched with launch_k(..., pdl = true)
// may become resident while its predecessor in the stream still runs; it must execute
// pdl_wait() before it touches anything the predecessor reads or writes (the wait returns once
// every prerequisite grid has completed and flushed).  The predecessor opens the window with
// pdl_trigger() (all of its CTAs must have issued it or exited).  Both are no-ops otherwise.
bool pdl_enabled();
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void pdl_trigger() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }

template <class... KArgs, class... Args>
inline cudaError_t launch_k(void (*kern)(KArgs...), dim3 grid, dim3 block, size_t smem,
                            cudaStream_t s, bool pdl, Args... args) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = grid;
  cfg.blockDim = block;
  cfg.dynamicSmemBytes = smem;
  cfg.stream = s;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = (pdl && pdl_enabled()) ? 1 : 0;
  return cudaLaunchKernelEx(&cfg, kern, KArgs(args)...);
}

constexpr int kD = MACR_EMBED_DIM;  // embedding width every kernel is specialised for
constexpr float kBceEps = 1e-10f;   // "+1e-10" of macr_mf/model.py:211

// ---- device helpers --------------------------------------------------------------------
__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// streaming 128-bit accesses: the Adam sweep touches every element exactly once per step
__device__ __forceinline__ float4 ld_stream(const float4 *p) {
  float4 r;
  asm volatile("ld.global.L1::no_allocate.v4.f32 {%0,%1,%2,%3}, [%4];"
               : "=f"(r.x), "=f"(r.y), "=f"(r.z), "=f"(r.w)
               : "l"(p));
  return r;
}
__device__ __forceinline__ void st_stream(float4 *p, const float4 &v) {
  asm volatile("st.global.L1::no_allocate.v4.f32 [%0], {%1,%2,%3,%4};" ::"l"(p), "f"(v.x),
               "f"(v.y), "f"(v.z), "f"(v.w)
               : "memory");
}

// MUFU-based transcendental pieces (ex2 / rcp / lg2 .approx) -- the grid kernel is bound by
// these, see DESIGN.md section 4.
__device__ __forceinline__ float fast_sigmoid(float x) {
  // 1/(1+exp(-x)); __expf = ex2.approx(x*log2e), __frcp_rn replaced by rcp.approx via __fdividef
  return __fdividef(1.0f, 1.0f + __expf(-x));
}

// Adam element updates with every operation individually rounded (no FMA contraction), in the
// order TF-1.14 evaluates them, so results are bit-identical to the CPU oracle.
// adam.py _apply_sparse_shared:  m = m*b1 (+ g*(1-b1));  v = v*b2 (+ (g*g)*(1-b2));
//                                var -= (lr_t*m)/(sqrt(v)+eps)
__device__ __forceinline__ void adam_decay_only(float &var, float &m, float &v, float lr_t,
                                                float b1, float b2, float eps) {
  m = __fmul_rn(m, b1);
  v = __fmul_rn(v, b2);
  var = __fsub_rn(var, __fdiv_rn(__fmul_rn(lr_t, m), __fadd_rn(__fsqrt_rn(v), eps)));
}
__device__ __forceinline__ void adam_with_grad(float &var, float &m, float &v, float g,
                                               float lr_t, float b1, float b2, float omb1,
                                               float omb2, float eps) {
  m = __fadd_rn(__fmul_rn(m, b1), __fmul_rn(g, omb1));
  v = __fadd_rn(__fmul_rn(v, b2), __fmul_rn(__fmul_rn(g, g), omb2));
  var = __fsub_rn(var, __fdiv_rn(__fmul_rn(lr_t, m), __fadd_rn(__fsqrt_rn(v), eps)));
}

}  // namespace macr
