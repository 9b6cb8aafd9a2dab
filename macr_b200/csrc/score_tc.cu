// score_tc.cu -- K7+K8 on the 5th-generation tensor cores: TMA-fed tcgen05 tiles (kind::f16, bf16
// operands, fp32 accumulators in TMEM) whose accumulator IS the approximate counterfactual score
// (y - c) * sig_i, followed by an EXACT fp32 re-rank, so the emitted ids and scores are
// bit-identical to the fp32 path of score.cu (and to the CPU oracle).
//   replaces sess.run(model.rubi_ratings_both, ...) + host top-K
//   (macr_mf/train.py:249-251,89-104; macr_lightgcn/utility/batch_test.py:85-134; model.py:45,199).
//
// Why a sampled threshold + ONE full pass.  A running per-row top-K in the epilogue costs
// ~K(1+ln(n/K)) list insertions per row, serialised across the 32 rows a warp owns: far more than
// the pass itself (half an instruction per score).  Instead:
//   prep         items: bf16(sig_i * I_i) and the three bf16 pieces of -c*sig_i as an augmented K
//                step (rows padded to whole tiles with -inf there); users: bf16(U), |u|.
//   pass MAX     over a SAMPLE of the item tiles (every `stride`-th tile, stride <= 4 by catalogue size): per row and
//                item chunk, 32 running maxima of the approximate score -- one per (32-column batch
//                position, column residue mod 4) -- kept in registers, train items of the row left
//                out.  The groups are disjoint item sets, so nothing but 128 bytes per (row, chunk)
//                is written: no per-batch maxima matrix.
//   threshold    m_K = K-th largest of the row's n_chunks*32 group maxima; thr = m_K - 2 eps - slack.
//                K distinct unmasked items have an approximate score >= m_K, so the exact K-th best
//                score is >= m_K - eps and every item of the true top-K scores >= thr in the next pass.
//   pass FILTER  ONE pass over the whole catalogue: every batch's maximum is compared with thr
//                (the same FMNMX chains as the maxima pass); scores >= thr that are not train items
//                of the row (a trained model ranks exactly those on top: a per-row cursor into the
//                sorted train list marks them tile by tile) are appended with their approximate
//                score to the row's candidate list (~K*stride entries; one atomic per append).
//   re-rank      one warp per row: a_K = K-th largest APPROXIMATE score of the list (radix select);
//                K items score >= a_K, so the exact K-th best is >= a_K - eps and only candidates
//                with approximate score >= a_K - 2 eps (~25) can be in the top-K: those are
//                re-scored with the exact fp32 FMA chain (the arithmetic of score.cu) and ranked by
//                counting under the order (score desc, lower id first).
//   fallback     a row whose candidate list overflowed (degenerate score distributions), a row
//                without K finite group maxima, or a row whose train list exceeds 1/16 of the
//                catalogue is done by the exact fp32 kernel (score.cu) -- never silently wrong.
// eps is a rigorous bound of |approximate - exact| (see row_threshold_kernel).  Operands are
// rounded once to bf16: measured on B200, a 128-row SS-mode MMA instruction takes the same time
// for tf32 (K=8) and bf16 (K=16), so bf16 halves the tensor time and the TMA bytes; a looser
// bound only means more candidates, never a wrong id.
//
// Kernel anatomy (one CTA per SM, persistent over (256-user tile, item chunk) work items):
//   warp 0      TMA producer   cp.async.bulk.tensor.2d: 128B-swizzled K-major boxes of 128-byte
//                              rows, 32B-swizzled boxes for the augmented step; two 128-row user
//                              tiles stay resident, 256-item tiles stream through a 4-stage ring
//                              (each item tile is read from L2 once per 256 users)
//   warp 1      MMA issuer     tcgen05.mma.cta_group::1.kind::f16, M=128 N=256 K=16, D in TMEM:
//                              2 accumulators x 256 columns (one per user tile); the tensor pipe
//                              works on one user tile while the other is drained; also owns
//                              TMEM alloc/dealloc
//   warps 2..3  idle           (complete the control warpgroup; they only give up registers)
//   warps 4..11 epilogue       two warpgroups, warpgroup g owns user tile g and accumulator g:
//                              tcgen05.ld.32x32b.x64 (thread = user row), batch maxima (FMNMX3)
//                              or threshold filter
// mbarrier pipelines: user tiles full/empty, item stages full/empty, TMEM full/empty.
// setmaxnreg re-splits the register file once the roles part (control 56, epilogue 224).
#include <cuda.h>
#include <math.h>

#include <type_traits>

#include "common.cuh"
#include "score.cuh"

namespace macr {
namespace tc {

constexpr int BM = 128;              // user rows per tile (UMMA M, TMEM lanes)
constexpr int BN = 256;              // items per tile (UMMA N, TMEM columns per accumulator)
constexpr int NB = BN / 32;          // 32-column batches per tile
typedef unsigned short oper_t;       // operands rounded to bf16
constexpr int KB = 64;               // bf16 per 128-byte swizzle row = the whole embedding row
static_assert(KB == kD && KB * sizeof(oper_t) == 128, "one swizzle row holds one embedding row");
constexpr int A_TILE_BYTES = BM * 128;   // a 128 x 64 bf16 user tile, 128B-swizzled
constexpr int B_TILE_BYTES = BN * 128;   // a 256 x 64 bf16 item tile
// The "- c*sig_i" term rides on one extra K=16 step: user side [1,1,1,0..0], item side the three
// bf16 pieces of -c*sig_i (exact to 24 bits), 32-byte rows, 32B-swizzled.
constexpr int KA = 16;
constexpr int A_AUG_BYTES = BM * 32;
constexpr int B_AUG_BYTES = BN * 32;
constexpr int B_BYTES = B_TILE_BYTES + B_AUG_BYTES;  // one ring stage (40 KiB)
constexpr int kCap = 256;            // candidate slots per row (approximate score, id)
constexpr int kKeep = 128;           // candidates re-scored exactly per row (4 x 32 lanes of the re-rank warp)
constexpr int kMaxStride = 4;        // the maxima pass samples every stride-th item tile
constexpr int kMaxChunksS = 16;      // item chunks of the maxima pass (32 group maxima each)
// 3 warpgroups: {TMA producer, MMA issuer, 2 idle warps} + 2 epilogue warpgroups.  The register
// file is re-split with setmaxnreg once the roles part: 384 x 168 = 128 x 56 + 256 x 224.
constexpr int kThreads = 384;
constexpr int kCtrlRegs = 56, kEpiRegs = 224;
constexpr int kTmemCols = 512;       // 2 accumulators (one per user tile) x 256 columns
constexpr int UT = 2;                // user tiles per CTA (one per epilogue warpgroup)
constexpr int STAGES = 4;            // item-tile ring (160 KiB)
constexpr int A_BYTES = UT * (A_TILE_BYTES + A_AUG_BYTES);
constexpr int SMEM_BYTES = 1024 /*align*/ + A_BYTES + STAGES * B_BYTES;
constexpr int MODE_MAX = 0, MODE_FILTER = 1;

// ---- PTX wrappers ----------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void *p) {
  return (uint32_t)__cvta_generic_to_shared(p);
}
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes)
               : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
// spin on a phase parity; a 4 s watchdog turns a protocol bug into a trap instead of a hang
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  uint32_t done = 0;
  long long t0 = 0;
  for (uint32_t it = 0;; ++it) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(done)
        : "r"(bar), "r"(parity)
        : "memory");
    if (done) return;
    if ((it & 0xfff) == 0xfff) {
      const long long now = clock64();
      if (t0 == 0) t0 = now;
      else if (now - t0 > 8000000000LL) __trap();
    }
  }
}
__device__ __forceinline__ void tma_load_2d(uint32_t dst, const CUtensorMap *map, uint32_t bar,
                                            int c0, int c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
      ::"r"(dst), "l"(reinterpret_cast<uint64_t>(map)), "r"(bar), "r"(c0), "r"(c1)
      : "memory");
}
__device__ __forceinline__ void tc_fence_before() {
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
}
__device__ __forceinline__ void tc_fence_after() {
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
}
__device__ __forceinline__ void tc_commit(uint32_t bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar)
               : "memory");
}
// D[tmem] (+)= A[smem] * B[smem]^T, tf32 inputs, fp32 accumulate
__device__ __forceinline__ void tc_mma(uint32_t d_tmem, uint64_t adesc, uint64_t bdesc,
                                       uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
      ::"r"(d_tmem), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}
// K-major, 128-byte swizzle: 8-row groups 1024 B apart (SBO), version 1 (sm_100), layout type 2
__device__ __forceinline__ uint64_t umma_desc(uint32_t smem_addr) {
  uint64_t d = 0;
  d |= (uint64_t)((smem_addr >> 4) & 0x3FFF);
  d |= (uint64_t)1 << 16;            // leading byte offset (ignored for swizzled K-major)
  d |= (uint64_t)(1024 >> 4) << 32;  // stride byte offset
  d |= (uint64_t)1 << 46;            // descriptor version
  d |= (uint64_t)2 << 61;            // SWIZZLE_128B
  return d;
}
// K-major, 32-byte swizzle (rows of 32 B): 8-row groups 256 B apart, layout type 6
__device__ __forceinline__ uint64_t umma_desc32(uint32_t smem_addr) {
  uint64_t d = 0;
  d |= (uint64_t)((smem_addr >> 4) & 0x3FFF);
  d |= (uint64_t)1 << 16;
  d |= (uint64_t)(256 >> 4) << 32;
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)6 << 61;  // SWIZZLE_32B
  return d;
}
// fp32 accumulate (bit 4), A/B format bf16 (1 at bits 7-9 / 10-12), both K-major,
// N >> 3 at bit 17, M >> 4 at bit 24
constexpr uint32_t kIdesc = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(BN >> 3) << 17) |
                            ((uint32_t)(BM >> 4) << 24);

#define MACR_R32(v)                                                                              \
  "+r"(v[0]), "+r"(v[1]), "+r"(v[2]), "+r"(v[3]), "+r"(v[4]), "+r"(v[5]), "+r"(v[6]), "+r"(v[7]), \
      "+r"(v[8]), "+r"(v[9]), "+r"(v[10]), "+r"(v[11]), "+r"(v[12]), "+r"(v[13]), "+r"(v[14]),    \
      "+r"(v[15]), "+r"(v[16]), "+r"(v[17]), "+r"(v[18]), "+r"(v[19]), "+r"(v[20]), "+r"(v[21]),  \
      "+r"(v[22]), "+r"(v[23]), "+r"(v[24]), "+r"(v[25]), "+r"(v[26]), "+r"(v[27]), "+r"(v[28]),  \
      "+r"(v[29]), "+r"(v[30]), "+r"(v[31])
#define MACR_W32(v)                                                                              \
  "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), \
      "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]),    \
      "=r"(v[15]), "=r"(v[16]), "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]),  \
      "=r"(v[22]), "=r"(v[23]), "=r"(v[24]), "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]),  \
      "=r"(v[29]), "=r"(v[30]), "=r"(v[31])

// 32 consecutive accumulator columns of this thread's TMEM lane (= user row)
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t (&v)[32]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : MACR_W32(v)
      : "r"(taddr)
      : "memory");
}
// the registers are operands so no use of them can be scheduled above the wait
__device__ __forceinline__ void tmem_ld_wait(uint32_t (&v)[32]) {
  asm volatile("tcgen05.wait::ld.sync.aligned;" : MACR_R32(v)::"memory");
}

// 64 consecutive accumulator columns of this thread's TMEM lane: fewer, longer loads hide the
// TMEM latency behind two batches of arithmetic
__device__ __forceinline__ void tmem_ld64(uint32_t taddr, uint32_t (&v)[64]) {
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x64.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31, %32, %33, %34, %35, %36, %37, %38, %39, %40, %41, %42, %43, %44, %45, %46, %47, %48, %49, %50, %51, %52, %53, %54, %55, %56, %57, %58, %59, %60, %61, %62, %63}, [%64];"
               : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]), "=r"(v[16]), "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]), "=r"(v[24]), "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31]), "=r"(v[32]), "=r"(v[33]), "=r"(v[34]), "=r"(v[35]), "=r"(v[36]), "=r"(v[37]), "=r"(v[38]), "=r"(v[39]), "=r"(v[40]), "=r"(v[41]), "=r"(v[42]), "=r"(v[43]), "=r"(v[44]), "=r"(v[45]), "=r"(v[46]), "=r"(v[47]), "=r"(v[48]), "=r"(v[49]), "=r"(v[50]), "=r"(v[51]), "=r"(v[52]), "=r"(v[53]), "=r"(v[54]), "=r"(v[55]), "=r"(v[56]), "=r"(v[57]), "=r"(v[58]), "=r"(v[59]), "=r"(v[60]), "=r"(v[61]), "=r"(v[62]), "=r"(v[63])
               : "r"(taddr)
               : "memory");
}
__device__ __forceinline__ void tmem_ld_wait64(uint32_t (&v)[64]) {
  asm volatile("tcgen05.wait::ld.sync.aligned;" : "+r"(v[0]), "+r"(v[1]), "+r"(v[2]), "+r"(v[3]), "+r"(v[4]), "+r"(v[5]), "+r"(v[6]), "+r"(v[7]), "+r"(v[8]), "+r"(v[9]), "+r"(v[10]), "+r"(v[11]), "+r"(v[12]), "+r"(v[13]), "+r"(v[14]), "+r"(v[15]), "+r"(v[16]), "+r"(v[17]), "+r"(v[18]), "+r"(v[19]), "+r"(v[20]), "+r"(v[21]), "+r"(v[22]), "+r"(v[23]), "+r"(v[24]), "+r"(v[25]), "+r"(v[26]), "+r"(v[27]), "+r"(v[28]), "+r"(v[29]), "+r"(v[30]), "+r"(v[31]), "+r"(v[32]), "+r"(v[33]), "+r"(v[34]), "+r"(v[35]), "+r"(v[36]), "+r"(v[37]), "+r"(v[38]), "+r"(v[39]), "+r"(v[40]), "+r"(v[41]), "+r"(v[42]), "+r"(v[43]), "+r"(v[44]), "+r"(v[45]), "+r"(v[46]), "+r"(v[47]), "+r"(v[48]), "+r"(v[49]), "+r"(v[50]), "+r"(v[51]), "+r"(v[52]), "+r"(v[53]), "+r"(v[54]), "+r"(v[55]), "+r"(v[56]), "+r"(v[57]), "+r"(v[58]), "+r"(v[59]), "+r"(v[60]), "+r"(v[61]), "+r"(v[62]), "+r"(v[63])::"memory");
}

struct TileParams {
  int T;            // query rows in this row block
  int n_items;      // items in this shard
  int n_utiles;     // 256-row user tile pairs
  int n_itiles, n_chunks, tiles_per_chunk;
  int id_off;       // global id of local item 0 (candidates carry global ids)
  int tile_stride;  // item tile visited for tile index t: t * tile_stride (maxima pass: the sample)
  int ld_g;         // maxima pass: row pitch of the group maxima (floats) = 32 * n_chunks
  int dbg;          // developer timing experiment: bit2 skips the MMAs (results are then meaningless)
};

// fold the 32 accumulator columns v[OFF .. OFF+32) into the 4 running maxima of their column
// residues mod 4 (4 independent FMNMX chains: the cost of a plain batch maximum)
template <int OFF, int N>
__device__ __forceinline__ void fold4(const uint32_t (&v)[N], float (&g)[4]) {
#pragma unroll
  for (int j4 = 0; j4 < 8; ++j4) {
    g[0] = fmaxf(g[0], __uint_as_float(v[OFF + 4 * j4 + 0]));
    g[1] = fmaxf(g[1], __uint_as_float(v[OFF + 4 * j4 + 1]));
    g[2] = fmaxf(g[2], __uint_as_float(v[OFF + 4 * j4 + 2]));
    g[3] = fmaxf(g[3], __uint_as_float(v[OFF + 4 * j4 + 3]));
  }
}
// The same for rows with train items among these 32 columns (`skip`: one bit per column).  The
// maxima only have to be a LOWER bound built from unmasked items, so a residue group that holds a
// train item simply does not take this tile's 8 columns (the 7 clean ones are given up: the
// threshold gets looser by a hair, never wrong).  Every lane of a warp with such a row runs the
// same ~30 instructions -- a lane with train items used to drag its warp through a 32-way
// select (half of all batches at the gowalla shape: maxima pass 86 -> 70 k cycles per CTA).
template <int OFF, int N>
__device__ __forceinline__ void fold4_skip(const uint32_t (&v)[N], float (&g)[4], uint32_t skip) {
  float m[4] = {-INFINITY, -INFINITY, -INFINITY, -INFINITY};
  fold4<OFF>(v, m);
#pragma unroll
  for (int r = 0; r < 4; ++r) g[r] = fmaxf(g[r], (skip & (0x11111111u << r)) ? -INFINITY : m[r]);
}

// ---------------------------------------------------------------------------------------------
// The item operand is pre-scaled by its gate (rows sig_i * I_i) and one extra K step adds
// -c*sig_i, so the accumulator IS the approximate score (y - c) * sig_i: the epilogue is a bare
// maximum / compare, half an instruction per score, with no shared-memory traffic.
//   MODE_MAX     visits item tiles t * tile_stride; per (row, chunk) 32 running group maxima
//                gmax[row][32*chunk + 4*b + r] over the columns of batch position b with residue r
//                (train items of the row excluded when MASKED), written once per work item
//   MODE_FILTER  every batch whose maximum reaches thr.x is scanned; scores >= thr.x that are not
//                train items of the row (MASKED) are appended with their approximate score
// Pipeline per CTA: the MMA warp alternates between the two user tiles (A0 x B -> TMEM[0:256),
// A1 x B -> TMEM[256:512)); while warpgroup 0 drains its accumulator the tensor pipe computes
// warpgroup 1's, so with epilogue <= MMA time the tensor pipe never idles.
// ---------------------------------------------------------------------------------------------
template <int MODE, bool MASKED>
__global__ void __launch_bounds__(kThreads, 1)
score_tc_kernel(const __grid_constant__ CUtensorMap tmU, const __grid_constant__ CUtensorMap tmI,
                const __grid_constant__ CUtensorMap tmUa, const __grid_constant__ CUtensorMap tmIa,
                const TileParams P, const int32_t *__restrict__ mask_rowptr,
                const int32_t *__restrict__ mask_col, float *__restrict__ gmax, const float2 *__restrict__ thr,
                uint2 *__restrict__ cand, int *__restrict__ cand_cnt,
                long long *__restrict__ prof /* developer cycle counters of CTA 0, nullable */) {
  extern __shared__ unsigned char smem_raw[];
  // 1024-byte alignment: the 128B swizzle pattern repeats every 8 rows x 128 B
  unsigned char *smem = reinterpret_cast<unsigned char *>(
      (reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  unsigned char *sA = smem;              // [UT][128 rows x 128 B] then [UT][128 rows x 32 B]
  unsigned char *sAaug = smem + UT * A_TILE_BYTES;
  unsigned char *sB = smem + A_BYTES;    // [STAGES]{[256 rows x 128 B], [256 rows x 32 B]}
  __shared__ __align__(8) uint64_t bars[32];   // see indices below
  __shared__ uint32_t tmem_slot[1];

  enum { A_FULL = 0, A_EMPTY = 1, TM_FULL = 2, TM_EMPTY = 4, B_FULL = 6, B_EMPTY = 6 + STAGES };
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t bar0 = smem_u32(bars);
  auto BAR = [&](int i) { return bar0 + 8u * (uint32_t)i; };

  if (threadIdx.x == 0) {
    mbar_init(BAR(A_FULL), 1);
    mbar_init(BAR(A_EMPTY), 1);
    for (int b = 0; b < 2; ++b) {
      mbar_init(BAR(TM_FULL + b), 1);
      mbar_init(BAR(TM_EMPTY + b), 4);  // one arrive per warp of the owning epilogue warpgroup
    }
    for (int s = 0; s < STAGES; ++s) {
      mbar_init(BAR(B_FULL + s), 1);
      mbar_init(BAR(B_EMPTY + s), 1);
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1) {  // TMEM: all 512 columns (one CTA per SM)
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(
                     smem_u32(tmem_slot)),
                 "r"(kTmemCols)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = tmem_slot[0];
  const int n_work = P.n_utiles * P.n_chunks;
  const bool profiling = prof != nullptr && blockIdx.x == 0;
  long long pc[6] = {0, 0, 0, 0, 0, 0};
  auto timed_wait = [&](uint32_t bar, uint32_t parity, int slot) {
    if (profiling) {
      const long long t0 = clock64();
      mbar_wait(bar, parity);
      pc[slot] += clock64() - t0;
    } else {
      mbar_wait(bar, parity);
    }
  };
  const long long t_start = profiling ? clock64() : 0;

  if (warp < 4) {
    asm volatile("setmaxnreg.dec.sync.aligned.u32 %0;" ::"n"(kCtrlRegs));
  if (warp == 0) {
    // =============================== TMA producer ===============================
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0, a_phase = 0;
      for (int w = blockIdx.x; w < n_work; w += gridDim.x) {
        const int ut = w / P.n_chunks, ch = w - ut * P.n_chunks;
        const int t_begin = ch * P.tiles_per_chunk;
        const int t_end = min(P.n_itiles, t_begin + P.tiles_per_chunk);
        timed_wait(BAR(A_EMPTY), a_phase ^ 1, 0);
        mbar_expect_tx(BAR(A_FULL), A_BYTES);
        for (int a = 0; a < UT; ++a) {
          tma_load_2d(smem_u32(sA + a * A_TILE_BYTES), &tmU, BAR(A_FULL), 0, (ut * UT + a) * BM);
          tma_load_2d(smem_u32(sAaug + a * A_AUG_BYTES), &tmUa, BAR(A_FULL), 0, 0);
        }
        a_phase ^= 1;
        for (int t = t_begin; t < t_end; ++t) {
          timed_wait(BAR(B_EMPTY + stage), phase ^ 1, 1);
          mbar_expect_tx(BAR(B_FULL + stage), B_BYTES);
          unsigned char *dst = sB + stage * B_BYTES;
          const int row0 = t * P.tile_stride * BN;
          tma_load_2d(smem_u32(dst), &tmI, BAR(B_FULL + stage), 0, row0);
          tma_load_2d(smem_u32(dst + B_TILE_BYTES), &tmIa, BAR(B_FULL + stage), 0, row0);
          if (++stage == STAGES) {
            stage = 0;
            phase ^= 1;
          }
        }
      }
      if (profiling) {
        prof[0] = pc[0], prof[1] = pc[1], prof[2] = clock64() - t_start;
      }
    }
  } else if (warp == 1) {
    // =============================== MMA issuer ===============================
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0, a_phase = 0;
      uint32_t n = 0;  // running tile counter -> TMEM barrier parity
      const uint32_t a_addr = smem_u32(sA), aaug_addr = smem_u32(sAaug);
      for (int w = blockIdx.x; w < n_work; w += gridDim.x) {
        const int ut = w / P.n_chunks, ch = w - ut * P.n_chunks;
        const int t_begin = ch * P.tiles_per_chunk;
        const int t_end = min(P.n_itiles, t_begin + P.tiles_per_chunk);
        timed_wait(BAR(A_FULL), a_phase, 0);
        a_phase ^= 1;
        for (int t = t_begin; t < t_end; ++t, ++n) {
          timed_wait(BAR(B_FULL + stage), phase, 1);
          const uint32_t b_addr = smem_u32(sB + stage * B_BYTES);
#pragma unroll
          for (int a = 0; a < UT; ++a) {
            timed_wait(BAR(TM_EMPTY + a), (n & 1) ^ 1, 2 + a);
            tc_fence_after();
            const uint32_t d_tmem = tmem_base + a * BN;
#pragma unroll
            for (int ks = 0; ks < 4; ++ks) {  // 32-byte K steps inside the 128-byte row
              const uint64_t ad = umma_desc(a_addr + a * A_TILE_BYTES + ks * 32);
              const uint64_t bd = umma_desc(b_addr + ks * 32);
              if (!(P.dbg & 4)) tc_mma(d_tmem, ad, bd, kIdesc, ks ? 1u : 0u);
            }
            // + 1 * (-c * sig_i)
            if (!(P.dbg & 4))
              tc_mma(d_tmem, umma_desc32(aaug_addr + a * A_AUG_BYTES),
                     umma_desc32(b_addr + B_TILE_BYTES), kIdesc, 1u);
            tc_commit(BAR(TM_FULL + a));  // accumulator ready for warpgroup a
          }
          tc_commit(BAR(B_EMPTY + stage));  // item stage may be refilled
          if (t == t_end - 1) tc_commit(BAR(A_EMPTY));
          if (++stage == STAGES) {
            stage = 0;
            phase ^= 1;
          }
        }
      }
      if (profiling) {
        prof[4] = pc[0], prof[5] = pc[1], prof[6] = pc[2], prof[7] = pc[3];
        prof[8] = clock64() - t_start, prof[9] = n;
      }
    }
  }
  } else {
    asm volatile("setmaxnreg.inc.sync.aligned.u32 %0;" ::"n"(kEpiRegs));
    // ===== epilogue: 2 warpgroups of 128 threads, warpgroup g owns user tile g of the pair =====
    const int g = (warp - 4) >> 2;          // warpgroup
    const int q = warp & 3;                 // TMEM lane quarter this warp may read
    const int r_in = q * 32 + lane;         // row inside the user tile
    const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + g * BN;
    uint32_t n = 0;
    for (int w = blockIdx.x; w < n_work; w += gridDim.x) {
      const int ut = w / P.n_chunks, ch = w - ut * P.n_chunks;
      const int t_begin = ch * P.tiles_per_chunk;
      const int t_end = min(P.n_itiles, t_begin + P.tiles_per_chunk);
      const int row = (ut * UT + g) * BM + r_in;
      const bool valid = row < P.T;
      float2 th = make_float2(INFINITY, INFINITY);
      uint2 *my_cand = nullptr;
      // MASKED: cursor into the row's sorted train-item list, positioned at the chunk's first
      // item and advanced tile by tile (kAhead entries are kept in flight).  Rows that go straight
      // to the exact kernel (filter pass: thr = +inf; maxima pass: the same train-list-length test
      // as row_threshold_kernel) append nothing and walk nothing.
      constexpr int kAhead = 4;
      int mptr = 0, mend = 0, nxt[kAhead];
#pragma unroll
      for (int d = 0; d < kAhead; ++d) nxt[d] = 0x7fffffff;
      if (valid) {
        bool walk = MASKED;
        if (MODE == MODE_FILTER) {
          th = thr[row];
          my_cand = cand + (size_t)row * kCap;
          walk = walk && th.x < INFINITY;
        }
        if (MASKED && walk) {
          mptr = mask_rowptr[row];
          mend = mask_rowptr[row + 1];
          if (MODE == MODE_MAX && mend - mptr > (P.n_items >> 4) + 64) mend = mptr;
          const int first = P.id_off + t_begin * P.tile_stride * BN;
          int lo = mptr, hi = mend;
          while (lo < hi) {
            const int mid = (lo + hi) >> 1;
            if (mask_col[mid] < first) lo = mid + 1;
            else hi = mid;
          }
          mptr = lo;
#pragma unroll
          for (int d = 0; d < kAhead; ++d)
            nxt[d] = mptr + d < mend ? mask_col[mptr + d] : 0x7fffffff;
        }
      }
      float gm[NB][4];  // MODE_MAX: running group maxima of this (row, chunk)
#pragma unroll
      for (int i = 0; i < NB; ++i) gm[i][0] = gm[i][1] = gm[i][2] = gm[i][3] = -INFINITY;

      for (int t = t_begin; t < t_end; ++t, ++n) {
        const uint32_t full_parity = n & 1;
        // the row's train items of this tile -> 8 mask words (a well-trained model ranks exactly
        // those items on top: unfiltered they would flood the candidate list / spoil the maxima)
        uint32_t mw[NB];
#pragma unroll
        for (int i = 0; i < NB; ++i) mw[i] = 0u;
        if (MASKED) {
          const int g0 = P.id_off + t * P.tile_stride * BN, g1 = g0 + BN;
          while (nxt[0] < g1) {
            const int b = nxt[0] - g0;
            if (b >= 0) {
              const uint32_t bit = 1u << (b & 31);
              const int ws = b >> 5;
#pragma unroll
              for (int i = 0; i < NB; ++i) mw[i] |= ws == i ? bit : 0u;
            }
#pragma unroll
            for (int d = 0; d + 1 < kAhead; ++d) nxt[d] = nxt[d + 1];
            ++mptr;
            nxt[kAhead - 1] = mptr + kAhead - 1 < mend ? mask_col[mptr + kAhead - 1] : 0x7fffffff;
          }
        }
        uint32_t va[64], vb[64];
        timed_wait(BAR(TM_FULL + g), full_parity, 1);
        tc_fence_after();
        __syncwarp();
        if (MODE == MODE_MAX) {
          // 64-column loads, one in flight while the previous 64 columns are folded
          auto fold = [&](auto off_tag, const uint32_t (&v)[64], int cb) {
            constexpr int OFF = decltype(off_tag)::value;
            // warp-uniform choice: large catalogues rarely have a train item in a 32-column batch
            if (MASKED && __any_sync(0xffffffffu, mw[cb] != 0u)) fold4_skip<OFF>(v, gm[cb], mw[cb]);
            else fold4<OFF>(v, gm[cb]);
          };
          tmem_ld64(taddr, va);
          tmem_ld_wait64(va);
#pragma unroll
          for (int c4 = 0; c4 < NB / 4; ++c4) {
            const int cb = 4 * c4;  // va holds batches cb, cb+1
            tmem_ld64(taddr + (cb + 2) * 32, vb);
            fold(std::integral_constant<int, 0>{}, va, cb);
            fold(std::integral_constant<int, 32>{}, va, cb + 1);
            tmem_ld_wait64(vb);
            if (cb + 4 < NB) {
              tmem_ld64(taddr + (cb + 4) * 32, va);
            } else {
              // all TMEM reads of this accumulator are complete: hand it back to the MMA warp
              tc_fence_before();
              __syncwarp();
              if (lane == 0) mbar_arrive(BAR(TM_EMPTY + g));
            }
            fold(std::integral_constant<int, 0>{}, vb, cb + 2);
            fold(std::integral_constant<int, 32>{}, vb, cb + 3);
            if (cb + 4 < NB) tmem_ld_wait64(va);
          }
        } else {
          // Same 64-column load pipeline.  A batch is looked at element by element only if its
          // maximum reached the threshold for some row of the warp, and then only by the lanes
          // concerned.  Padded columns score -inf.
          auto scan = [&](auto off_tag, const uint32_t (&v)[64], int cb) {
            constexpr int OFF = decltype(off_tag)::value;
            float m4[4] = {-INFINITY, -INFINITY, -INFINITY, -INFINITY};
            fold4<OFF>(v, m4);
            const bool need = fmaxf(fmaxf(m4[0], m4[1]), fmaxf(m4[2], m4[3])) >= th.x;
            if (__any_sync(0xffffffffu, need)) {
              if (need) {
                const int gbase = P.id_off + t * BN + cb * 32;
#pragma unroll
                for (int j4 = 0; j4 < 8; ++j4) {
                  const float s0 = __uint_as_float(v[OFF + 4 * j4 + 0]);
                  const float s1 = __uint_as_float(v[OFF + 4 * j4 + 1]);
                  const float s2 = __uint_as_float(v[OFF + 4 * j4 + 2]);
                  const float s3 = __uint_as_float(v[OFF + 4 * j4 + 3]);
                  if (fmaxf(fmaxf(s0, s1), fmaxf(s2, s3)) >= th.x) {
                    const float ss[4] = {s0, s1, s2, s3};
#pragma unroll
                    for (int e = 0; e < 4; ++e) {
                      if (ss[e] >= th.x && !(MASKED && ((mw[cb] >> (4 * j4 + e)) & 1u))) {
                        // ~K * stride appends per row over the whole catalogue: the atomic is rare
                        const int pos = atomicAdd(cand_cnt + row, 1);
                        if (pos < kCap)
                          my_cand[pos] =
                              make_uint2(__float_as_uint(ss[e]), (uint32_t)(gbase + 4 * j4 + e));
                      }
                    }
                  }
                }
              }
              __syncwarp();
            }
          };
          tmem_ld64(taddr, va);
          tmem_ld_wait64(va);
#pragma unroll
          for (int c4 = 0; c4 < NB / 4; ++c4) {
            const int cb = 4 * c4;  // va holds batches cb, cb+1
            tmem_ld64(taddr + (cb + 2) * 32, vb);
            scan(std::integral_constant<int, 0>{}, va, cb);
            scan(std::integral_constant<int, 32>{}, va, cb + 1);
            tmem_ld_wait64(vb);
            if (cb + 4 < NB) {
              tmem_ld64(taddr + (cb + 4) * 32, va);
            } else {
              tc_fence_before();
              __syncwarp();
              if (lane == 0) mbar_arrive(BAR(TM_EMPTY + g));
            }
            scan(std::integral_constant<int, 0>{}, vb, cb + 2);
            scan(std::integral_constant<int, 32>{}, vb, cb + 3);
            if (cb + 4 < NB) tmem_ld_wait64(va);
          }
        }
      }
      if (MODE == MODE_MAX && valid) {
        float4 *gp = reinterpret_cast<float4 *>(gmax + (size_t)row * P.ld_g + 32 * ch);
#pragma unroll
        for (int i = 0; i < NB; ++i) gp[i] = make_float4(gm[i][0], gm[i][1], gm[i][2], gm[i][3]);
      }
    }
    if (profiling && lane == 0 && q == 0) {  // one thread per warpgroup
      prof[12 + 4 * g] = pc[0], prof[13 + 4 * g] = pc[1], prof[14 + 4 * g] = clock64() - t_start;
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base),
                 "r"(kTmemCols)
                 : "memory");
  }
}

// ---------------------------------------------------------------------------------------------
// operand preparation: x' = x * scale[row] (items: the gate sig_i; users: 1), out = bf16(x')
// (round to nearest); row norm of x'.  Items also get their augmented-K row: the three bf16
// pieces of -c*sig_i, and rows n .. n_pad-1 (padding to a whole tile) are zeros with the
// augmented entry -inf, so padded columns score -inf.
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ unsigned short to_bf16(float x) {
  unsigned short r;
  asm("cvt.rn.bf16.f32 %0, %1;" : "=h"(r) : "f"(x));
  return r;
}
__device__ __forceinline__ float from_bf16(unsigned short h) {
  return __uint_as_float((uint32_t)h << 16);
}

// one half-warp per row (float4 per lane); norm_out[row] = ||row||_2 rounded up a little;
// norm_max (nullable): max over rows via atomicMax on the bit pattern (non-negative floats)
__global__ void __launch_bounds__(256)
split_rows_kernel(const float *__restrict__ X, long long n, long long n_pad,
                  const float *__restrict__ scale, float c, oper_t *__restrict__ out,
                  oper_t *__restrict__ aug, float *__restrict__ norm_out,
                  unsigned int *__restrict__ norm_max) {
  const long long r = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 4;
  const int hl = threadIdx.x & 15;
  float ss = 0.f;
  if (r < n) {
    float4 v = reinterpret_cast<const float4 *>(X + r * kD)[hl];
    if (scale) {
      const float sc = scale[r];
      v.x = __fmul_rn(v.x, sc), v.y = __fmul_rn(v.y, sc), v.z = __fmul_rn(v.z, sc),
      v.w = __fmul_rn(v.w, sc);
      if (hl == 0) {
        const float t0 = -__fmul_rn(c, sc);
        const unsigned short p1 = to_bf16(t0);
        const float t1 = t0 - from_bf16(p1);  // exact
        const unsigned short p2 = to_bf16(t1);
        const unsigned short p3 = to_bf16(t1 - from_bf16(p2));
        uint4 lo = make_uint4((uint32_t)p1 | ((uint32_t)p2 << 16), (uint32_t)p3, 0u, 0u);
        reinterpret_cast<uint4 *>(aug + r * KA)[0] = lo;
        reinterpret_cast<uint4 *>(aug + r * KA)[1] = make_uint4(0u, 0u, 0u, 0u);
      }
    }
    ushort4 h;
    h.x = to_bf16(v.x), h.y = to_bf16(v.y), h.z = to_bf16(v.z), h.w = to_bf16(v.w);
    reinterpret_cast<ushort4 *>(out + r * kD)[hl] = h;
    ss = v.x * v.x + v.y * v.y + v.z * v.z + v.w * v.w;
  } else if (r < n_pad) {
    reinterpret_cast<ushort4 *>(out + r * kD)[hl] = make_ushort4(0, 0, 0, 0);
    if (hl == 0 && aug) {
      reinterpret_cast<uint4 *>(aug + r * KA)[0] = make_uint4(0xFF80u /* -inf */, 0u, 0u, 0u);
      reinterpret_cast<uint4 *>(aug + r * KA)[1] = make_uint4(0u, 0u, 0u, 0u);
    }
  }
#pragma unroll
  for (int o = 8; o > 0; o >>= 1) ss += __shfl_xor_sync(0xffffffffu, ss, o);
  if (r < n && hl == 0) {
    const float nrm = sqrtf(ss) * 1.0001f;  // the sum of squares itself is rounded
    if (norm_out) norm_out[r] = nrm;
    if (norm_max) atomicMax(norm_max, __float_as_uint(nrm));
  }
}

// user-side augmented rows: [1, 1, 1, 0 .. 0] x 128 (the same tile for every user tile)
__global__ void fill_user_aug_kernel(oper_t *__restrict__ aug) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < BM * KA) aug[i] = (i % KA) < 3 ? (oper_t)0x3F80 /* 1.0 */ : (oper_t)0;
}

// orderable key of a float: a larger float has a larger key; key 0 is below every float (-inf
// included) and marks entries to ignore
__device__ __forceinline__ uint32_t fkey(float x) {
  const uint32_t u = __float_as_uint(x);
  return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
}
__device__ __forceinline__ float fkey_inv(uint32_t k) {
  return __uint_as_float((k & 0x80000000u) ? (k & 0x7fffffffu) : ~k);
}
// K-th largest (K >= 1) of the warp's NV x 32 keys by radix select: 32 rounds of (NV compares +
// one warp reduction).  Returns 0 when fewer than K keys are non-zero.
template <int NV>
__device__ __forceinline__ uint32_t warp_kth_largest(const uint32_t (&key)[NV], int K) {
  // bits on which all non-zero keys agree need no round: scores above one threshold share their
  // sign, exponent and leading mantissa bits
  uint32_t all_and = 0xffffffffu, all_or = 0u;
  int n_valid = 0;
#pragma unroll
  for (int r = 0; r < NV; ++r) {
    all_and &= key[r] ? key[r] : 0xffffffffu;
    all_or |= key[r];
    n_valid += key[r] != 0u;
  }
  all_and = __reduce_and_sync(0xffffffffu, all_and);
  all_or = __reduce_or_sync(0xffffffffu, all_or);
  n_valid = __reduce_add_sync(0xffffffffu, n_valid);
  if (n_valid < K) return 0u;
  const uint32_t differ = all_and ^ all_or;
  if (differ == 0u) return all_or;  // all valid keys are equal
  const int top = 31 - __clz(differ);
  uint32_t prefix = top == 31 ? 0u : (all_or >> (top + 1)) << (top + 1);  // the common leading bits
  int remaining = K;
#pragma unroll 1
  for (int bit = top; bit >= 0; --bit) {
    const uint32_t want = (prefix | (1u << bit)) >> bit;
    int cnt = 0;
#pragma unroll
    for (int r = 0; r < NV; ++r) cnt += (key[r] >> bit) == want ? 1 : 0;
    cnt = __reduce_add_sync(0xffffffffu, cnt);
    if (cnt >= remaining) prefix |= 1u << bit;
    else remaining -= cnt;
  }
  return prefix;
}

// thr[row] = {filter threshold, eps}, one warp per row.
// The maxima pass left n_groups = 32 * n_chunks group maxima per row; the groups are disjoint sets
// of unmasked sampled items, so K distinct unmasked items score at least the K-th largest group
// maximum m_K in the approximate arithmetic: the exact K-th best score of the row is >= m_K - eps1.
// With y~ the tensor-core dot product of the gate-scaled item row and y the fp32 FMA chain:
//   |y~ - sig*y| <= kappa * |u| * |sig*i|
//   kappa = 1.5 * 2^-8  (two bf16 roundings 2^-9 each -> 2^-8 (1 + 2^-10) per product, products
//   exact in fp32, plus fp32 accumulation of 64 terms on both sides and the gate pre-scale)
// and the roundings of (sig*y - c*sig), with c*sig carried as three bf16 pieces (24 bits), versus
// ((y - c) * sig) add at most 2^-22 * (|c| + |u||i|).  eps = 2 * eps1 covers (a) the sampled
// maximum vs the exact score plus (b) the exact score vs the filter pass's approximate score.
__global__ void __launch_bounds__(256, 4)
row_threshold_kernel(const float *__restrict__ gmax, int T, int n_groups, int ld_g, int K,
                     const int32_t *__restrict__ mask_rowptr, int n_items,
                     const float *__restrict__ unorm, const unsigned int *__restrict__ inorm_max,
                     float c, float kappa_sum, float2 *__restrict__ thr, int *__restrict__ cand_cnt) {
  const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
  const int t = blockIdx.x * 8 + wib;
  if (t >= T) return;
  // 32 groups per chunk; groups g, g + 128, g + 256, ... are merged (a union of disjoint item sets
  // is one) so the select always runs on 128 values, 4 per lane
  constexpr int NV = 4;
  uint32_t key[NV];
  {
    // all (<= kMaxChunksS) loads of the lane issued before any is used: a latency-bound kernel
    float v[NV][kMaxChunksS / NV];
#pragma unroll
    for (int r = 0; r < NV; ++r)
#pragma unroll
      for (int q = 0; q < kMaxChunksS / NV; ++q) {
        const int idx = r * 32 + lane + 32 * NV * q;
        v[r][q] = idx < n_groups ? gmax[(size_t)t * ld_g + idx] : -INFINITY;
      }
#pragma unroll
    for (int r = 0; r < NV; ++r) {
      float m = v[r][0];
#pragma unroll
      for (int q = 1; q < kMaxChunksS / NV; ++q) m = fmaxf(m, v[r][q]);
      key[r] = r * 32 + lane < n_groups ? fkey(m) : 0u;
    }
  }
  const uint32_t kk = warp_kth_largest<NV>(key, K);
  const float mk = kk ? fkey_inv(kk) : -INFINITY;
  if (lane == 0) {
    // Straight to the exact kernel (which strikes train items out tile by tile): rows without K
    // finite group maxima -- no usable threshold, every item would be a candidate -- and rows whose
    // train list is dense enough (> 1/16 of the catalogue) that walking it costs the tensor-core
    // passes more than the exact kernel costs the row.  +inf keeps the filter pass off the row;
    // the over-full count makes the re-rank kernel queue it.
    const int mask_len = mask_rowptr ? mask_rowptr[t + 1] - mask_rowptr[t] : 0;
    float2 out = make_float2(INFINITY, 0.f);
    if (!(mk > -INFINITY) || mask_len > (n_items >> 4) + 64) {
      cand_cnt[t] = kCap + 1;
    } else {
      const float yb = unorm[t] * __uint_as_float(*inorm_max);
      const float eps = kappa_sum * yb + 9.5367431640625e-7f /*2^-20*/ * (fabsf(c) + yb);
      out.x = mk - eps - 9.5367431640625e-7f * fabsf(mk);
      out.y = eps;
    }
    thr[t] = out;
  }
}

// Candidates of one row (one warp per row).  (1) a_K = K-th largest approximate score of the list;
// K unmasked items score >= a_K approximately, hence >= a_K - eps1 exactly, so a member of the true
// top-K has an approximate score >= a_K - 2 eps1 = a_K - eps: everything below is dropped without
// being touched.  (2) each lane fetches and re-scores one survivor per round with the fp32 FMA
// chain of score.cu; ranks come from counting.  (Train items never get here: the filter pass marks
// every entry of the row's sorted train list that falls into a tile.)  Rows whose list overflowed
// (or with more than kKeep survivors) are queued for the exact fp32 kernel.
// 5 resident CTAs/SM (48 registers, no spills): A/B on one box 0.290 -> 0.281 ms per gowalla-shape call
// (3 CTAs/SM at 80 registers before; 4 -> 0.285, 6 -> 0.282)
__global__ void __launch_bounds__(256, 5)
rerank_kernel(const float *__restrict__ Uq, int T, const float *__restrict__ It,
              const float *__restrict__ sig_i, const float *__restrict__ sig_u, float c, int id_off,
              const uint2 *__restrict__ cand, const int *__restrict__ cand_cnt,
              const float2 *__restrict__ thr, int K,
              int32_t *__restrict__ out_ids, float *__restrict__ out_scores,
              int32_t *__restrict__ fb_rows, int *__restrict__ fb_count,
              unsigned long long *__restrict__ cand_total) {
  __shared__ float su[8][kD];
  __shared__ int s_keep[8][kKeep];
  const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
  const int t = blockIdx.x * 8 + wib;
  if (t >= T) return;
  constexpr int kRounds = kKeep / 32;  // survivors live in registers, 32 per round
  constexpr int kLoad = kCap / 32;
  const int total = cand_cnt[t];
  if (total > kCap) {
    if (lane == 0) fb_rows[atomicAdd(fb_count, 1)] = t;
    return;
  }
  // ---- (1) cut on the approximate scores ----
  uint32_t key[kLoad];
  int gidv[kLoad];
  const int nload = (total + 31) >> 5;  // warp-uniform: lists are ~K * stride long, 3 of the 8 rounds
#pragma unroll
  for (int r = 0; r < kLoad; ++r) {
    key[r] = 0u;
    gidv[r] = 0;
    if (r < nload) {
      const int e = r * 32 + lane;
      if (e < total) {
        const uint2 cv = cand[(size_t)t * kCap + e];
        key[r] = fkey(__uint_as_float(cv.x));
        gidv[r] = (int)cv.y;
      }
    }
  }
  uint32_t cut = 1u;  // fewer than K candidates (fewer than K unmasked items exist): keep all
  if (total > K) {
    // lists are ~K * stride long: select over as many 32-entry rounds as the list fills
    const uint32_t kk = total <= 64    ? warp_kth_largest<2>(reinterpret_cast<const uint32_t (&)[2]>(key), K)
                        : total <= 128 ? warp_kth_largest<4>(reinterpret_cast<const uint32_t (&)[4]>(key), K)
                                       : warp_kth_largest<kLoad>(key, K);
    const float aK = fkey_inv(kk);
    cut = fkey(aK - thr[t].y - 9.5367431640625e-7f * fabsf(aK));
  }
  int n_keep = 0;
#pragma unroll
  for (int r = 0; r < kLoad; ++r) {
    if (r >= nload) break;  // warp-uniform
    const bool keep = key[r] >= cut && key[r] != 0u;
    const unsigned bal = __ballot_sync(0xffffffffu, keep);
    const int pos = n_keep + __popc(bal & ((1u << lane) - 1u));
    if (keep && pos < kKeep) s_keep[wib][pos] = gidv[r];
    n_keep += __popc(bal);
  }
  if (n_keep > kKeep) {
    if (lane == 0) fb_rows[atomicAdd(fb_count, 1)] = t;
    return;
  }
  if (cand_total && lane == 0) atomicAdd(cand_total, (unsigned long long)n_keep);
  su[wib][lane] = Uq[(size_t)t * kD + lane];
  su[wib][lane + 32] = Uq[(size_t)t * kD + lane + 32];
  __syncwarp();
  // ---- (2) exact scores of the survivors ----
  const float sgu = sig_u[t];
  float cs_[kRounds];
  int cg_[kRounds];
#pragma unroll
  for (int r = 0; r < kRounds; ++r) {
    cs_[r] = -INFINITY;
    cg_[r] = 0x7fffffff;
    const int e = r * 32 + lane;
    if (e < n_keep) {
      const int gid = s_keep[wib][e];
      const float4 *ip = reinterpret_cast<const float4 *>(It + (size_t)(gid - id_off) * kD);
      float acc = 0.f;
#pragma unroll
      for (int q4 = 0; q4 < kD / 4; ++q4) {
        const float4 v = ip[q4];
        acc = fmaf(su[wib][4 * q4 + 0], v.x, acc);
        acc = fmaf(su[wib][4 * q4 + 1], v.y, acc);
        acc = fmaf(su[wib][4 * q4 + 2], v.z, acc);
        acc = fmaf(su[wib][4 * q4 + 3], v.w, acc);
      }
      cs_[r] = __fmul_rn(__fmul_rn(__fsub_rn(acc, c), sig_i[gid - id_off]), sgu);
      cg_[r] = gid;
    }
  }
  // rank by counting: the order (score desc, lower id first) is strict among valid candidates
  const int nr = (n_keep + 31) >> 5;
  int rank[kRounds];
#pragma unroll
  for (int r = 0; r < kRounds; ++r) rank[r] = 0;
#pragma unroll
  for (int r2 = 0; r2 < kRounds; ++r2) {
    if (r2 < nr) {
      for (int l2 = 0; l2 < 32; ++l2) {
        const float os = __shfl_sync(0xffffffffu, cs_[r2], l2);
        const int og = __shfl_sync(0xffffffffu, cg_[r2], l2);
#pragma unroll
        for (int r = 0; r < kRounds; ++r)
          if (r < nr) rank[r] += score_better(os, og, cs_[r], cg_[r]) ? 1 : 0;  // warp-uniform
      }
    }
  }
  if (lane < K) {
    out_ids[(size_t)t * K + lane] = -1;
    out_scores[(size_t)t * K + lane] = -INFINITY;
  }
  __syncwarp();
#pragma unroll
  for (int r = 0; r < kRounds; ++r) {
    if (cg_[r] != 0x7fffffff && rank[r] < K) {
      out_ids[(size_t)t * K + rank[r]] = cg_[r];
      out_scores[(size_t)t * K + rank[r]] = cs_[r];
    }
  }
}

__global__ void accumulate_stats_kernel(int64_t *stats, const int *fb_count,
                                        const unsigned long long *cand_total) {
  stats[0] += *fb_count;
  stats[1] += (int64_t)*cand_total;
}
static void accumulate_stats(int64_t *stats, const int *fb_count,
                             const unsigned long long *cand_total, cudaStream_t s) {
  accumulate_stats_kernel<<<1, 1, 0, s>>>(stats, fb_count, cand_total);
}

// ---------------------------------------------------------------------------------------------
// host side
// ---------------------------------------------------------------------------------------------
typedef CUresult (*EncodeTiledFn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *,
                                  const cuuint64_t *, const cuuint64_t *, const cuuint32_t *,
                                  const cuuint32_t *, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn encode_fn() {
  static EncodeTiledFn fn = nullptr;
  if (!fn) {
    void *p = nullptr;
    cudaDriverEntryPointQueryResult qr;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qr) ==
            cudaSuccess &&
        qr == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<EncodeTiledFn>(p);
  }
  return fn;
}

// [rows][64] operands row-major -> boxes of 128 rows x 128 B, 128-byte swizzle;
// rows beyond `rows` read as zeros
// cols = 64 (128-byte rows, 128B swizzle) or 16 (32-byte augmented rows, 32B swizzle)
static int make_map(CUtensorMap *m, const oper_t *base, long long rows, int box_rows, int cols) {
  EncodeTiledFn fn = encode_fn();
  if (!fn) return fail(MACR_ERR_CUDA, "cuTensorMapEncodeTiled is not available from the driver");
  cuuint64_t dims[2] = {(cuuint64_t)cols, (cuuint64_t)rows};
  cuuint64_t strides[1] = {(cuuint64_t)cols * sizeof(oper_t)};
  cuuint32_t box[2] = {(cuuint32_t)cols, (cuuint32_t)box_rows};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = fn(m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<oper_t *>(base), dims, strides,
                  box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                  cols == kD ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_32B,
                  CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) return fail(MACR_ERR_CUDA, "cuTensorMapEncodeTiled failed (%d)", (int)r);
  return MACR_OK;
}

struct Plan {
  int TB;        // query rows per row block
  int n_itiles, n_chunks, tiles_per_chunk;          // filter pass: the whole catalogue
  int stride, n_stiles, n_chunks_s, tiles_per_chunk_s, ld_g;  // maxima pass: the sampled tiles
  size_t scratch, off_items;  // bytes of per-call scratch; where macr_score_topk_tc keeps its items
  size_t off_uhi, off_unorm, off_misc, off_gmax, off_thr, off_cand,
      off_cnt, off_fbrows, off_exact, total;
  size_t exact_bytes;
};

static size_t align_up(size_t x, size_t a) { return (x + a - 1) / a * a; }

// item chunks: enough work items to balance the SMs, every chunk non-empty; -> tiles per chunk
static int pick_tiles_per_chunk(int utiles, int n_tiles, int nc_min, int nc_max) {
  const int sms = sm_count();
  int best = nc_min;
  double best_cost = 1e30;
  for (int nc = nc_min; nc <= nc_max; ++nc) {
    const int tpc = (n_tiles + nc - 1) / nc;
    const int ncr = (n_tiles + tpc - 1) / tpc;
    const long long work = (long long)utiles * ncr;
    const long long rounds = (work + sms - 1) / sms;
    const double cost = (double)rounds * tpc;  // tiles on the busiest SM
    if (cost < best_cost * 0.97) {
      best_cost = cost;
      best = ncr;
    }
  }
  return (n_tiles + best - 1) / best;
}

// prepared item operands: bf16 rows scaled by sig_i, augmented rows (three bf16 pieces of
// -c*sig_i; -inf for the padding rows), the user-side augmented tile, and the largest item norm
struct ItemsLayout {
  size_t off_ihi, off_iaug, off_uaug, off_norm, total;
  long long n_pad;
};
static ItemsLayout items_layout(long long n_items) {
  ItemsLayout L;
  L.n_pad = (n_items + BN - 1) / BN * BN;  // items padded to whole tiles
  size_t o = 0;
  auto take = [&](size_t bytes) {
    const size_t at = o;
    o = align_up(o + bytes, 1024);
    return at;
  };
  L.off_ihi = take((size_t)L.n_pad * kD * sizeof(oper_t));
  L.off_iaug = take((size_t)L.n_pad * KA * sizeof(oper_t));
  L.off_uaug = take((size_t)BM * KA * sizeof(oper_t));
  L.off_norm = take(64);  // [0] item norm max (uint bits)
  L.total = o;
  return L;
}

static Plan make_plan(int T, long long n_items, int K) {
  Plan p;
  p.n_itiles = (int)((n_items + BN - 1) / BN);
  // the maxima pass visits every stride-th tile: the threshold then sits near rank K * stride of
  // the row instead of K (that many candidates reach the re-rank's approximate cut), for 1/stride
  // of a pass; small catalogues keep enough sampled batches for tight group maxima
  // measured (profiles/r2s_stride.txt): 15 424 x 40 981 (161 tiles): 0.351 / 0.317 / 0.303 / 0.320 ms at stride
  // 1 / 2 / 3 / 4 (a looser threshold makes more warps take the filter pass's append path);
  // 262 144 x 1 M (3907 tiles): 52.2 ms at stride 2, 43.9 ms at stride 4
  p.stride = p.n_itiles >= 256 ? kMaxStride : p.n_itiles >= 96 ? 3 : p.n_itiles >= 64 ? 2 : 1;
  {
    static int env_stride = -1;
    if (env_stride < 0) {
      const char *e = getenv("MACR_TC_STRIDE");  // developer knob
      env_stride = e ? atoi(e) : 0;
    }
    if (env_stride >= 1 && env_stride <= kMaxStride) p.stride = env_stride;
  }
  p.n_stiles = (p.n_itiles + p.stride - 1) / p.stride;
  // row blocks: group maxima + candidate lists at most ~2 GiB per block, and blocks of equal size
  // (a short trailing block would leave most SMs idle for a whole pass over the catalogue)
  long long tb = (2048LL << 20) / (4LL * 32 * kMaxChunksS + 8LL * kCap + 64);
  tb = tb / (UT * BM) * (UT * BM);
  if (tb >= T) {
    tb = T;
  } else {
    const long long n_blocks = (T + tb - 1) / tb;
    tb = ((T + n_blocks - 1) / n_blocks + UT * BM - 1) / (UT * BM) * (UT * BM);
  }
  p.TB = (int)tb;
  const int utiles = (p.TB + UT * BM - 1) / (UT * BM);
  int nc_max = p.n_itiles / 2 < 16 ? p.n_itiles / 2 : 16;
  p.tiles_per_chunk = pick_tiles_per_chunk(utiles, p.n_itiles, 1, nc_max < 1 ? 1 : nc_max);
  p.n_chunks = (p.n_itiles + p.tiles_per_chunk - 1) / p.tiles_per_chunk;
  // >= 4 chunks = 128 disjoint groups per row (K <= 32), more when few user tiles must fill the SMs
  const int ncs_max = p.n_stiles < kMaxChunksS ? p.n_stiles : kMaxChunksS;
  p.tiles_per_chunk_s = pick_tiles_per_chunk(utiles, p.n_stiles, ncs_max < 4 ? ncs_max : 4, ncs_max);
  p.n_chunks_s = (p.n_stiles + p.tiles_per_chunk_s - 1) / p.tiles_per_chunk_s;
  p.ld_g = 32 * p.n_chunks_s;
  size_t o = 0;
  auto take = [&](size_t bytes) {
    const size_t at = o;
    o = align_up(o + bytes, 1024);
    return at;
  };
  p.off_uhi = take((size_t)p.TB * kD * sizeof(oper_t));
  p.off_unorm = take((size_t)p.TB * 4);
  p.off_misc = take(64);  // [1] fb_count [2..3] cand_total (u64)
  p.off_gmax = take((size_t)p.TB * p.ld_g * 4);
  p.off_thr = take((size_t)p.TB * 8);
  p.off_cand = take((size_t)p.TB * kCap * 8);
  p.off_cnt = take((size_t)p.TB * 4);
  p.off_fbrows = take((size_t)p.TB * 4);
  p.exact_bytes = score_exact_workspace_bytes(p.TB, n_items, K);
  p.off_exact = take(p.exact_bytes);
  // the item-side operands depend on (items, sig_i, c) alone: a caller that scores several query
  // blocks against one model prepares them once (macr_score_tc_prepare_items) into its own
  // buffer; macr_score_topk_tc keeps one behind its scratch
  p.scratch = o;
  p.off_items = o;
  p.total = o + items_layout(n_items).total + 1024;
  return p;
}

static int g_dbg = 0;
static long long *g_prof = nullptr;  // device int64[64]: cycle counters of CTA 0, per pass

template <int MODE, bool MASKED>
static int launch_pass(const CUtensorMap &mu, const CUtensorMap &mi, const CUtensorMap &mua,
                       const CUtensorMap &mia, const TileParams &P, const int32_t *mrp,
                       const int32_t *mcol, float *gmax, const float2 *thr, uint2 *cand, int *cnt,
                       cudaStream_t s) {
  static bool opted = false;
  long long *prof = g_prof ? g_prof + 32 * MODE : nullptr;
  if (!opted) {
    MACR_CUDA(cudaFuncSetAttribute(score_tc_kernel<MODE, MASKED>,
                                   cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_BYTES));
    opted = true;
  }
  const int n_work = P.n_utiles * P.n_chunks;
  const int grid = n_work < sm_count() ? n_work : sm_count();
  score_tc_kernel<MODE, MASKED><<<grid, kThreads, SMEM_BYTES, s>>>(mu, mi, mua, mia, P, mrp, mcol,
                                                                   gmax, thr, cand, cnt, prof);
  MACR_LAUNCH_CHECK();
  return MACR_OK;
}

}  // namespace tc
}  // namespace macr

using namespace macr;

// developer hook (not in the public header): timing experiments of the maxima pass
extern "C" int macr_score_tc_debug(int dbg, long long *prof_dev) {
  tc::g_dbg = dbg;
  tc::g_prof = prof_dev;
  return MACR_OK;
}

extern "C" size_t macr_score_topk_tc_workspace_bytes(int T, int64_t n_items, int K) {
  if (T <= 0 || n_items <= 0 || K <= 0) return 1024;
  return tc::make_plan(T, n_items, K).total;
}

static int prepare_items(const float *It, long long n_items, const float *sig_i, float c,
                         unsigned char *iw, cudaStream_t s) {
  using namespace tc;
  const ItemsLayout L = items_layout(n_items);
  oper_t *ihi = reinterpret_cast<oper_t *>(iw + L.off_ihi);
  oper_t *iaug = reinterpret_cast<oper_t *>(iw + L.off_iaug);
  oper_t *uaug = reinterpret_cast<oper_t *>(iw + L.off_uaug);
  unsigned int *norm = reinterpret_cast<unsigned int *>(iw + L.off_norm);
  MACR_CUDA(cudaMemsetAsync(norm, 0, 64, s));
  split_rows_kernel<<<(unsigned)((L.n_pad * 16 + 255) / 256), 256, 0, s>>>(
      It, n_items, L.n_pad, sig_i, c, ihi, iaug, nullptr, norm);
  MACR_LAUNCH_CHECK();
  fill_user_aug_kernel<<<(BM * KA + 255) / 256, 256, 0, s>>>(uaug);
  MACR_LAUNCH_CHECK();
  return MACR_OK;
}

static int check_tc_shape(const char *who, int d, int K, int T, int64_t n_items) {
  using namespace tc;
  MACR_CHECK_ARG(d == kD, "%s: d must be %d (got %d)", who, kD, d);
  MACR_CHECK_ARG(K >= 1 && K <= 32, "%s: K must be in [1,32] (got %d)", who, K);
  MACR_CHECK_ARG(T >= 0 && n_items >= 0, "%s: negative size", who);
  MACR_CHECK_ARG(T == 0 || (n_items >= 2048 && n_items < (1LL << 31) - BN),
                 "%s: needs at least 2048 items (got %lld): use macr_score_topk", who,
                 (long long)n_items);
  return MACR_OK;
}

extern "C" size_t macr_score_tc_items_bytes(int64_t n_items) {
  return n_items > 0 ? tc::items_layout(n_items).total : 1024;
}

extern "C" int macr_score_tc_prepare_items(const float *It, int64_t n_items, int d,
                                           const float *sig_i, float c, void *items_ws,
                                           size_t items_bytes, macr_stream_t stream) {
  int rc = check_tc_shape("macr_score_tc_prepare_items", d, 1, 1, n_items);
  if (rc) return rc;
  MACR_CHECK_ARG(It && sig_i && items_ws, "macr_score_tc_prepare_items: null pointer");
  MACR_CHECK_ARG((reinterpret_cast<uintptr_t>(items_ws) & 1023) == 0,
                 "macr_score_tc_prepare_items: buffer must be 1024-byte aligned");
  if (items_bytes < tc::items_layout(n_items).total)
    return fail(MACR_ERR_WORKSPACE, "macr_score_tc_prepare_items: buffer %zu < %zu bytes",
                items_bytes, tc::items_layout(n_items).total);
  return prepare_items(It, n_items, sig_i, c, reinterpret_cast<unsigned char *>(items_ws),
                       as_stream(stream));
}

// the passes, thresholds and re-rank of one call; `iw` holds the prepared item operands
static int score_topk_tc_run(const float *Uq, int T, const float *It, int64_t n_items,
                             const float *sig_i, const float *sig_u, float c,
                             const int32_t *mask_rowptr, const int32_t *mask_col, int K,
                             int32_t item_id_offset, int32_t *out_ids, float *out_scores,
                             const unsigned char *iw, unsigned char *w, const tc::Plan &p,
                             int64_t *stats, cudaStream_t s) {
  using namespace tc;
  const ItemsLayout L = items_layout(n_items);
  oper_t *uhi = reinterpret_cast<oper_t *>(w + p.off_uhi);
  const oper_t *ihi = reinterpret_cast<const oper_t *>(iw + L.off_ihi);
  float *unorm = reinterpret_cast<float *>(w + p.off_unorm);
  unsigned int *misc = reinterpret_cast<unsigned int *>(w + p.off_misc);
  const unsigned int *inorm_max = reinterpret_cast<const unsigned int *>(iw + L.off_norm);
  float *gmax = reinterpret_cast<float *>(w + p.off_gmax);
  float2 *thr = reinterpret_cast<float2 *>(w + p.off_thr);
  const oper_t *iaug = reinterpret_cast<const oper_t *>(iw + L.off_iaug);
  const oper_t *uaug = reinterpret_cast<const oper_t *>(iw + L.off_uaug);
  uint2 *cand = reinterpret_cast<uint2 *>(w + p.off_cand);
  int *cnt = reinterpret_cast<int *>(w + p.off_cnt);
  int32_t *fb_rows = reinterpret_cast<int32_t *>(w + p.off_fbrows);
  int *fb_count = reinterpret_cast<int *>(misc + 1);
  unsigned long long *cand_total = reinterpret_cast<unsigned long long *>(misc + 2);

  MACR_CUDA(cudaMemsetAsync(misc, 0, 64, s));
  const long long n_pad = L.n_pad;
  CUtensorMap mih, mia, mua;
  int rc = make_map(&mih, ihi, n_pad, BN, kD);
  if (rc) return rc;
  rc = make_map(&mia, iaug, n_pad, BN, KA);
  if (rc) return rc;
  rc = make_map(&mua, uaug, BM, BM, KA);
  if (rc) return rc;
  // eps of row_threshold_kernel: the sampled maximum vs exact + exact vs the filter pass's score
  const float kappa_sum = 2.f * 1.5f * 3.90625e-3f;

  for (int t0 = 0; t0 < T; t0 += p.TB) {
    const int nb = T - t0 < p.TB ? T - t0 : p.TB;
    const float *Ub = Uq + (size_t)t0 * kD;
    const int32_t *mrp = mask_rowptr ? mask_rowptr + t0 : nullptr;
    split_rows_kernel<<<(unsigned)(((long long)nb * 16 + 255) / 256), 256, 0, s>>>(
        Ub, nb, nb, nullptr, 0.f, uhi, nullptr, unorm, nullptr);
    MACR_LAUNCH_CHECK();
    CUtensorMap muh;
    rc = make_map(&muh, uhi, nb, BM, kD);
    if (rc) return rc;
    TileParams P;
    P.T = nb;
    P.n_items = (int)n_items;
    P.n_utiles = (nb + UT * BM - 1) / (UT * BM);
    P.id_off = item_id_offset;
    P.ld_g = p.ld_g;
    P.dbg = g_dbg;
    // maxima pass over the sampled tiles (train items of the row excluded inside the pass)
    P.n_itiles = p.n_stiles;
    P.n_chunks = p.n_chunks_s;
    P.tiles_per_chunk = p.tiles_per_chunk_s;
    P.tile_stride = p.stride;
    rc = mrp ? launch_pass<MODE_MAX, true>(muh, mih, mua, mia, P, mrp, mask_col, gmax, nullptr,
                                           nullptr, nullptr, s)
             : launch_pass<MODE_MAX, false>(muh, mih, mua, mia, P, nullptr, nullptr, gmax, nullptr,
                                            nullptr, nullptr, s);
    if (rc) return rc;
    MACR_CUDA(cudaMemsetAsync(cnt, 0, (size_t)nb * sizeof(int), s));
    row_threshold_kernel<<<(nb + 7) / 8, 256, 0, s>>>(gmax, nb, p.ld_g, p.ld_g, K, mrp, (int)n_items,
                                                      unorm, inorm_max, c, kappa_sum, thr, cnt);
    MACR_LAUNCH_CHECK();
    // ONE pass over the whole catalogue; train items are filtered inside it whenever a mask is given
    P.n_itiles = p.n_itiles;
    P.n_chunks = p.n_chunks;
    P.tiles_per_chunk = p.tiles_per_chunk;
    P.tile_stride = 1;
    rc = mrp ? launch_pass<MODE_FILTER, true>(muh, mih, mua, mia, P, mrp, mask_col, nullptr, thr,
                                              cand, cnt, s)
             : launch_pass<MODE_FILTER, false>(muh, mih, mua, mia, P, nullptr, nullptr, nullptr, thr,
                                               cand, cnt, s);
    if (rc) return rc;
    MACR_CUDA(cudaMemsetAsync(fb_count, 0, sizeof(int), s));
    rerank_kernel<<<(nb + 7) / 8, 256, 0, s>>>(Ub, nb, It, sig_i, sig_u + t0, c, item_id_offset,
                                               cand, cnt, thr, K,
                                               out_ids + (size_t)t0 * K, out_scores + (size_t)t0 * K,
                                               fb_rows, fb_count, cand_total);
    MACR_LAUNCH_CHECK();
    // rows whose candidate lists overflowed: exact fp32 kernel on the queued rows (device-side
    // count, so nothing is read back; the launch is a no-op when the queue is empty)
    rc = score_exact_rows(Ub, nb, It, n_items, sig_i, sig_u + t0, c, mrp, mask_col, K,
                          item_id_offset, fb_rows, fb_count, out_ids + (size_t)t0 * K,
                          out_scores + (size_t)t0 * K, w + p.off_exact, p.exact_bytes, s);
    if (rc) return rc;
    if (stats) {
      // stats[0] += rows sent to the exact kernel, stats[1] += candidates re-ranked
      accumulate_stats(stats, fb_count, cand_total, s);
      MACR_CUDA(cudaMemsetAsync(cand_total, 0, sizeof(unsigned long long), s));
    }
  }
  return MACR_OK;
}

extern "C" int macr_score_topk_tc(const float *Uq, int T, const float *It, int64_t n_items, int d,
                                  const float *sig_i, const float *sig_u, float c,
                                  const int32_t *mask_rowptr, const int32_t *mask_col, int K,
                                  int32_t item_id_offset, int32_t *out_ids, float *out_scores,
                                  void *ws, size_t ws_bytes, int64_t *stats,
                                  macr_stream_t stream) {
  int rc = check_tc_shape("macr_score_topk_tc", d, K, T, n_items);
  if (rc) return rc;
  if (T == 0) return MACR_OK;
  MACR_CHECK_ARG(Uq && It && sig_i && sig_u && out_ids && out_scores && ws,
                 "macr_score_topk_tc: null pointer");
  MACR_CHECK_ARG((reinterpret_cast<uintptr_t>(ws) & 1023) == 0,
                 "macr_score_topk_tc: workspace must be 1024-byte aligned");
  const tc::Plan p = tc::make_plan(T, n_items, K);
  if (ws_bytes < p.total)
    return fail(MACR_ERR_WORKSPACE, "macr_score_topk_tc: workspace %zu < %zu bytes", ws_bytes,
                p.total);
  unsigned char *w = reinterpret_cast<unsigned char *>(ws);
  rc = prepare_items(It, n_items, sig_i, c, w + p.off_items, as_stream(stream));
  if (rc) return rc;
  return score_topk_tc_run(Uq, T, It, n_items, sig_i, sig_u, c, mask_rowptr, mask_col, K,
                           item_id_offset, out_ids, out_scores, w + p.off_items, w, p, stats,
                           as_stream(stream));
}

extern "C" size_t macr_score_topk_tc_prepared_workspace_bytes(int T, int64_t n_items, int K) {
  if (T <= 0 || n_items <= 0 || K <= 0) return 1024;
  return tc::make_plan(T, n_items, K).scratch + 1024;
}

extern "C" int macr_score_topk_tc_prepared(const float *Uq, int T, const float *It,
                                           int64_t n_items, int d, const float *sig_i,
                                           const float *sig_u, float c,
                                           const int32_t *mask_rowptr, const int32_t *mask_col,
                                           int K, int32_t item_id_offset, int32_t *out_ids,
                                           float *out_scores, const void *items_ws, void *ws,
                                           size_t ws_bytes, int64_t *stats, macr_stream_t stream) {
  int rc = check_tc_shape("macr_score_topk_tc_prepared", d, K, T, n_items);
  if (rc) return rc;
  if (T == 0) return MACR_OK;
  MACR_CHECK_ARG(Uq && It && sig_i && sig_u && out_ids && out_scores && ws && items_ws,
                 "macr_score_topk_tc_prepared: null pointer");
  MACR_CHECK_ARG((reinterpret_cast<uintptr_t>(ws) & 1023) == 0 &&
                     (reinterpret_cast<uintptr_t>(items_ws) & 1023) == 0,
                 "macr_score_topk_tc_prepared: buffers must be 1024-byte aligned");
  const tc::Plan p = tc::make_plan(T, n_items, K);
  if (ws_bytes < p.scratch)
    return fail(MACR_ERR_WORKSPACE, "macr_score_topk_tc_prepared: workspace %zu < %zu bytes",
                ws_bytes, p.scratch);
  return score_topk_tc_run(Uq, T, It, n_items, sig_i, sig_u, c, mask_rowptr, mask_col, K,
                           item_id_offset, out_ids, out_scores,
                           reinterpret_cast<const unsigned char *>(items_ws),
                           reinterpret_cast<unsigned char *>(ws), p, stats, as_stream(stream));
}
