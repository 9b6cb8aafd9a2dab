// score.cu -- K7+K8: full-catalogue counterfactual score, train-item mask and top-K, fused.
//   S[t,i] = ((u_t . i_i) - c) * sig(i_i . w) * sig(u_t . w_user)
// replaces sess.run(model.rubi_ratings_both, ...) + host top-K (macr_mf/train.py:249-251,89-104;
// macr_lightgcn/utility/batch_test.py:85-134; model.py:45,199; tools.h:13-33).
//
// This file is the EXACT fp32 path: every dot product is the fp32 FMA chain k = 0..63, so the CPU
// oracle reproduces the scores bit for bit and top-K ids are bit-exact (ties -> lower id).
// CTA tile: 128 query users x 128 items, K-dim 64 staged k-major in shared memory, 256 threads
// with an 8x8 register micro-tile; the score tile goes to shared memory (never to HBM) and each
// warp folds the 16 user rows it owns into register-resident sorted top-K lists (one rank per
// lane, K <= 32).  Item chunks (blockIdx.y) give enough CTAs to fill 148 SMs; their partial
// lists are merged by topk_merge_kernel, the same kernel that merges per-GPU shards.
#include <math.h>

#include "common.cuh"
#include "score.cuh"

namespace macr {

constexpr int kTU = 128, kTN = 128;
constexpr int kMaxKFast = 32;

// Train-item cursor of one row: narrows [lo, hi) of the row's ascending train list until at most
// 32 entries below `key` remain in front (the tile loop consumes those without marking).
__device__ __forceinline__ int mask_seek(const int32_t *__restrict__ mask_col, int lo, int hi,
                                         int key) {
  while (hi - lo > 32) {
    const int mid = (lo + hi) >> 1;
    if (mask_col[mid] < key) lo = mid + 1;
    else hi = mid;
  }
  return lo;
}

// Shape of a launch over the overflow queue of the tcgen05 path.  The host sizes the grid for all
// T rows (the queue length lives on the device); the few rows actually queued re-divide the same
// CTAs into more item chunks so the whole GPU works on them.  chunks * pitch never exceeds the
// host's chunks_full * roundup(T, 128), which is what the partial-list workspace holds.
struct QueueShape {
  int rows, utiles, chunks, pitch;
};
__device__ __forceinline__ QueueShape queue_shape(int T, int chunks_full, long long itiles,
                                                  const int *__restrict__ n_rows_dev) {
  QueueShape q;
  q.rows = min(T, *n_rows_dev);
  q.utiles = (q.rows + kTU - 1) / kTU;
  const long long ctas = (long long)((T + kTU - 1) / kTU) * chunks_full;
  q.chunks = q.utiles ? (int)min(min(ctas / q.utiles, itiles), 64LL) : 0;
  q.pitch = q.utiles * kTU;
  return q;
}

struct ScoreSmem {
  float sU[kD][kTU];       // k-major query-user tile
  float sI[kD][kTN];       // k-major item tile
  float sS[kTU][kTN + 4];  // score tile
  float sigI[kTN];
  float sigU[kTU];
};

template <bool kTopK>
__global__ void __launch_bounds__(256, 1)
score_kernel(const float *__restrict__ Uq, int T, const float *__restrict__ It, long long n_items,
             const float *__restrict__ sig_i, const float *__restrict__ sig_u, float c,
             const int32_t *__restrict__ mask_rowptr, const int32_t *__restrict__ mask_col, int K,
             int id_off, long long chunk_items, int32_t *__restrict__ part_ids,
             float *__restrict__ part_scores, float *__restrict__ out_matrix,
             const int32_t *__restrict__ row_map, const int *__restrict__ n_rows_dev) {
  // row_map != nullptr: slot t of this launch is query row row_map[t], and only the first
  // *n_rows_dev slots exist (overflow queue of the tcgen05 path); outputs stay in slot order
  int Tpitch = T, ut_idx = blockIdx.x, chunk_idx = blockIdx.y;
  if (row_map) {
    const QueueShape q = queue_shape(T, gridDim.y, (n_items + kTN - 1) / kTN, n_rows_dev);
    const int lin = blockIdx.y * gridDim.x + blockIdx.x;
    if (lin >= q.utiles * q.chunks) return;
    ut_idx = lin % q.utiles;
    chunk_idx = lin / q.utiles;
    T = q.rows;
    Tpitch = q.pitch;
    chunk_items = (((n_items + kTN - 1) / kTN + q.chunks - 1) / q.chunks) * kTN;
  }
  if (ut_idx * kTU >= T) return;
  extern __shared__ __align__(16) unsigned char smem_raw[];
  ScoreSmem &sm = *reinterpret_cast<ScoreSmem *>(smem_raw);
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int tx = tid & 15, ty = tid >> 4;
  const int u0 = ut_idx * kTU;
  const long long i_begin = (long long)chunk_idx * chunk_items;
  const long long i_end = min(n_items, i_begin + chunk_items);

  // stage the user tile once (transposing: consecutive threads -> consecutive rows)
  {
    const int r = tid & (kTU - 1), kq0 = tid >> 7;
    const int t = u0 + r;
    for (int kq = kq0; kq < kD / 4; kq += 2) {
      float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
      if (t < T) v = reinterpret_cast<const float4 *>(Uq + (long long)(row_map ? row_map[t] : t) * kD)[kq];
      sm.sU[4 * kq + 0][r] = v.x;
      sm.sU[4 * kq + 1][r] = v.y;
      sm.sU[4 * kq + 2][r] = v.z;
      sm.sU[4 * kq + 3][r] = v.w;
    }
    if (tid < kTU)
      sm.sigU[tid] = (u0 + tid < T) ? sig_u[row_map ? row_map[u0 + tid] : u0 + tid] : 0.f;
  }

  // per-warp top-K lists for users warp*16 .. warp*16+15 (rank = lane)
  float ls[16];
  int li[16];
  int mlo[16], mhi[16];
  const unsigned kmask = (K >= 32) ? 0xffffffffu : ((1u << K) - 1u);
  if (kTopK) {
#pragma unroll
    for (int j = 0; j < 16; ++j) {
      ls[j] = -INFINITY;
      li[j] = 0x7fffffff;
      const int t = u0 + warp * 16 + j;
      const int tr = (row_map && t < T) ? row_map[t] : t;
      mlo[j] = (mask_rowptr && t < T) ? mask_rowptr[tr] : 0;
      mhi[j] = (mask_rowptr && t < T) ? mask_rowptr[tr + 1] : 0;
      if (id_off + i_begin > 0) mlo[j] = mask_seek(mask_col, mlo[j], mhi[j], id_off + (int)i_begin);
    }
  }

  for (long long it0 = i_begin; it0 < i_end; it0 += kTN) {
    __syncthreads();  // previous tile fully consumed
    {
      const int r = tid & (kTN - 1), kq0 = tid >> 7;
      const long long i = it0 + r;
      for (int kq = kq0; kq < kD / 4; kq += 2) {
        float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
        if (i < i_end) v = reinterpret_cast<const float4 *>(It + i * kD)[kq];
        sm.sI[4 * kq + 0][r] = v.x;
        sm.sI[4 * kq + 1][r] = v.y;
        sm.sI[4 * kq + 2][r] = v.z;
        sm.sI[4 * kq + 3][r] = v.w;
      }
      if (tid < kTN) sm.sigI[tid] = (it0 + tid < i_end) ? sig_i[it0 + tid] : 0.f;
    }
    __syncthreads();

    // ---- phase 1: 8x8 micro-tile, fp32 FMA chain over k ascending --------------------------
    float acc[8][8];
#pragma unroll
    for (int a = 0; a < 8; ++a)
#pragma unroll
      for (int b = 0; b < 8; ++b) acc[a][b] = 0.f;
#pragma unroll 4
    for (int k = 0; k < kD; ++k) {
      const float4 ua = *reinterpret_cast<const float4 *>(&sm.sU[k][ty * 4]);
      const float4 ub = *reinterpret_cast<const float4 *>(&sm.sU[k][64 + ty * 4]);
      const float4 ia = *reinterpret_cast<const float4 *>(&sm.sI[k][tx * 4]);
      const float4 ib = *reinterpret_cast<const float4 *>(&sm.sI[k][64 + tx * 4]);
      const float uu[8] = {ua.x, ua.y, ua.z, ua.w, ub.x, ub.y, ub.z, ub.w};
      const float ii[8] = {ia.x, ia.y, ia.z, ia.w, ib.x, ib.y, ib.z, ib.w};
#pragma unroll
      for (int a = 0; a < 8; ++a)
#pragma unroll
        for (int b = 0; b < 8; ++b) acc[a][b] = fmaf(uu[a], ii[b], acc[a][b]);
    }
    // epilogue: ((y - c) * sig_i) * sig_u, left to right as in model.py:199
#pragma unroll
    for (int a = 0; a < 8; ++a) {
      const int ur = (a < 4) ? ty * 4 + a : 64 + ty * 4 + (a - 4);
      const float su = sm.sigU[ur];
#pragma unroll
      for (int h = 0; h < 2; ++h) {
        const int ic = h * 64 + tx * 4;
        float4 o;
        o.x = __fmul_rn(__fmul_rn(__fsub_rn(acc[a][h * 4 + 0], c), sm.sigI[ic + 0]), su);
        o.y = __fmul_rn(__fmul_rn(__fsub_rn(acc[a][h * 4 + 1], c), sm.sigI[ic + 1]), su);
        o.z = __fmul_rn(__fmul_rn(__fsub_rn(acc[a][h * 4 + 2], c), sm.sigI[ic + 2]), su);
        o.w = __fmul_rn(__fmul_rn(__fsub_rn(acc[a][h * 4 + 3], c), sm.sigI[ic + 3]), su);
        if (kTopK) {
          *reinterpret_cast<float4 *>(&sm.sS[ur][ic]) = o;
        } else {
          const int t = u0 + ur;
          const long long i = it0 + ic;
          if (t < T) {
            float *dst = out_matrix + (long long)t * n_items + i;
            if (i + 3 < i_end && ((n_items & 3) == 0)) {
              *reinterpret_cast<float4 *>(dst) = o;
            } else {
              if (i + 0 < i_end) dst[0] = o.x;
              if (i + 1 < i_end) dst[1] = o.y;
              if (i + 2 < i_end) dst[2] = o.z;
              if (i + 3 < i_end) dst[3] = o.w;
            }
          }
        }
      }
    }
    if (!kTopK) continue;
    __syncthreads();

    // ---- train items of this tile never compete: a well-trained model ranks exactly those on
    // top, so they are struck out of the score tile (NaN compares false everywhere) before the
    // fold instead of being looked up one candidate at a time.  mlo[j] is a cursor into row j's
    // ascending train list; 32 entries per row are fetched at once, all 16 rows in flight.
    if (mask_rowptr) {
      const int tile_end = id_off + (int)min(it0 + kTN, i_end);  // first global id past the tile
      int mv[16];
#pragma unroll
      for (int j = 0; j < 16; ++j)
        mv[j] = (mlo[j] + lane < mhi[j]) ? mask_col[mlo[j] + lane] : 0x7fffffff;
#pragma unroll
      for (int j = 0; j < 16; ++j) {
        int v = mv[j];
        for (;;) {
          const bool take = v < tile_end;
          const int loc = v - id_off - (int)it0;
          if (take && loc >= 0) sm.sS[warp * 16 + j][loc] = __int_as_float(0x7fc00000);
          const int cnt = __popc(__ballot_sync(0xffffffffu, take));
          mlo[j] += cnt;
          if (cnt < 32) break;
          v = (mlo[j] + lane < mhi[j]) ? mask_col[mlo[j] + lane] : 0x7fffffff;
        }
      }
      __syncwarp();
    }

    // ---- phase 2: each warp folds its 16 user rows into the register lists ------------------
#pragma unroll
    for (int j = 0; j < 16; ++j) {
      const int ur = warp * 16 + j;
      const int t = u0 + ur;
      if (t >= T) continue;
      float ws = __shfl_sync(0xffffffffu, ls[j], K - 1);
      int wi = __shfl_sync(0xffffffffu, li[j], K - 1);
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        const int colm = lane + 32 * e;
        const float s = sm.sS[ur][colm];
        const int gid = id_off + (int)(it0 + colm);
        const bool pass = (it0 + colm < i_end) && score_better(s, gid, ws, wi);
        unsigned m = __ballot_sync(0xffffffffu, pass);
        while (m) {
          const int src = __ffs(m) - 1;
          m &= m - 1;
          const float cs = __shfl_sync(0xffffffffu, s, src);
          const int cid = id_off + (int)(it0 + src + 32 * e);
          if (!score_better(cs, cid, ws, wi)) continue;
          score_list_insert(ls[j], li[j], cs, cid, lane, kmask, K);
          ws = __shfl_sync(0xffffffffu, ls[j], K - 1);
          wi = __shfl_sync(0xffffffffu, li[j], K - 1);
        }
      }
    }
  }

  if (kTopK) {
#pragma unroll
    for (int j = 0; j < 16; ++j) {
      const int t = u0 + warp * 16 + j;
      if (t < T && lane < K) {
        const long long o = ((long long)chunk_idx * Tpitch + t) * K + lane;
        const bool empty = li[j] == 0x7fffffff;
        part_ids[o] = empty ? -1 : li[j];
        part_scores[o] = empty ? -INFINITY : ls[j];
      }
    }
  }
}

// merge G sorted candidate lists per row -> top-K (score desc, lower id first); one warp per row
__global__ void __launch_bounds__(256)
topk_merge_kernel(const int32_t *__restrict__ ids, const float *__restrict__ scores, int T, int K,
                  int G, int32_t *__restrict__ out_ids, float *__restrict__ out_scores,
                  const int32_t *__restrict__ row_map, const int *__restrict__ n_rows_dev,
                  long long itiles) {
  const int lane = threadIdx.x & 31;
  const int t = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (row_map) {  // overflow queue: the lists were written in the queue's own shape
    const QueueShape q = queue_shape(T, G, itiles, n_rows_dev);
    if (t >= q.rows) return;
    G = q.chunks;
    T = q.pitch;
  }
  if (t >= T) return;
  const int to = row_map ? row_map[t] : t;  // output row
  const unsigned kmask = (K >= 32) ? 0xffffffffu : ((1u << K) - 1u);
  float ls = -INFINITY;
  int li = 0x7fffffff;
  for (int g = 0; g < G; ++g) {
    const long long base = ((long long)g * T + t) * K;
    const float s = lane < K ? scores[base + lane] : -INFINITY;
    const int id = lane < K ? ids[base + lane] : -1;
    for (int k = 0; k < K; ++k) {
      const float cs = __shfl_sync(0xffffffffu, s, k);
      const int cid = __shfl_sync(0xffffffffu, id, k);
      if (cid < 0) break;  // lists are padded at the tail
      const float ws = __shfl_sync(0xffffffffu, ls, K - 1);
      const int wi = __shfl_sync(0xffffffffu, li, K - 1);
      if (!score_better(cs, cid, ws, wi)) break;  // sorted input: the rest of this list loses too
      score_list_insert(ls, li, cs, cid, lane, kmask, K);
    }
  }
  if (lane < K) {
    const bool empty = li == 0x7fffffff;
    out_ids[(long long)to * K + lane] = empty ? -1 : li;
    out_scores[(long long)to * K + lane] = empty ? -INFINITY : ls;
  }
}

// top-K column indices of each row of a caller-supplied score matrix (tools.h:24-33 contract)
__global__ void __launch_bounds__(256)
topk_rows_kernel(const float *__restrict__ scores, int cols, int rows, int K,
                 int32_t *__restrict__ rankings) {
  const int lane = threadIdx.x & 31;
  const int r = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (r >= rows) return;
  const unsigned kmask = (K >= 32) ? 0xffffffffu : ((1u << K) - 1u);
  const float *row = scores + (long long)r * cols;
  float ls = -INFINITY;
  int li = 0x7fffffff;
  float ws = -INFINITY;
  int wi = 0x7fffffff;
  for (int c0 = 0; c0 < cols; c0 += 32) {
    const int cidx = c0 + lane;
    const float s = cidx < cols ? row[cidx] : -INFINITY;
    unsigned m = __ballot_sync(0xffffffffu, cidx < cols && score_better(s, cidx, ws, wi));
    while (m) {
      const int src = __ffs(m) - 1;
      m &= m - 1;
      const float cs = __shfl_sync(0xffffffffu, s, src);
      const int cid = c0 + src;
      if (!score_better(cs, cid, ws, wi)) continue;
      score_list_insert(ls, li, cs, cid, lane, kmask, K);
      ws = __shfl_sync(0xffffffffu, ls, K - 1);
      wi = __shfl_sync(0xffffffffu, li, K - 1);
    }
  }
  if (lane < K) rankings[(long long)r * K + lane] = (li == 0x7fffffff) ? -1 : li;
}

// sig[r] = sigmoid(rows[r] . w): fp32 FMA chain, sigmoid in fp64 rounded once to fp32
__global__ void __launch_bounds__(256)
score_gates_kernel(const float *__restrict__ rows, long long n, const float *__restrict__ wvec,
                   float *__restrict__ sig) {
  __shared__ float sw[kD];
  if (threadIdx.x < kD) sw[threadIdx.x] = wvec[threadIdx.x];
  __syncthreads();
  const long long r = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (r >= n) return;
  const float4 *p = reinterpret_cast<const float4 *>(rows + r * kD);
  float acc = 0.f;
#pragma unroll
  for (int q = 0; q < kD / 4; ++q) {
    const float4 v = p[q];
    acc = fmaf(v.x, sw[4 * q + 0], acc);
    acc = fmaf(v.y, sw[4 * q + 1], acc);
    acc = fmaf(v.z, sw[4 * q + 2], acc);
    acc = fmaf(v.w, sw[4 * q + 3], acc);
  }
  sig[r] = (float)(1.0 / (1.0 + exp(-(double)acc)));
}

__global__ void __launch_bounds__(256)
gather_rows_kernel(const float *__restrict__ table, const int32_t *__restrict__ ids, int n,
                   float *__restrict__ out) {
  const int q = (blockIdx.x * blockDim.x + threadIdx.x) >> 4, hl = threadIdx.x & 15;
  if (q >= n) return;
  reinterpret_cast<float4 *>(out + (long long)q * kD)[hl] =
      reinterpret_cast<const float4 *>(table + (long long)ids[q] * kD)[hl];
}

static int pick_chunks(int T, long long n_items) {
  const int utiles = (T + kTU - 1) / kTU;
  const long long itiles = (n_items + kTN - 1) / kTN;
  long long want = (2LL * sm_count() + utiles - 1) / utiles;  // >= 2 CTAs per SM overall
  if (want < 1) want = 1;
  if (want > itiles) want = itiles > 0 ? itiles : 1;
  if (want > 64) want = 64;
  return (int)want;
}

}  // namespace macr

using namespace macr;

extern "C" int macr_score_gates(const float *rows, int64_t n, int d, const float *wvec,
                                float *sig_out, macr_stream_t stream) {
  MACR_CHECK_ARG(d == kD, "macr_score_gates: d must be %d (got %d)", kD, d);
  MACR_CHECK_ARG(n >= 0, "macr_score_gates: negative n");
  if (n == 0) return MACR_OK;
  MACR_CHECK_ARG(rows && wvec && sig_out, "macr_score_gates: null pointer");
  score_gates_kernel<<<(unsigned)((n + 255) / 256), 256, 0, as_stream(stream)>>>(rows, n, wvec,
                                                                                 sig_out);
  MACR_LAUNCH_CHECK();
  return MACR_OK;
}

extern "C" int macr_gather_rows(const float *table, const int32_t *ids, int n, int d, float *out,
                                macr_stream_t stream) {
  MACR_CHECK_ARG(d == kD, "macr_gather_rows: d must be %d (got %d)", kD, d);
  MACR_CHECK_ARG(n >= 0, "macr_gather_rows: negative n");
  if (n == 0) return MACR_OK;
  MACR_CHECK_ARG(table && ids && out, "macr_gather_rows: null pointer");
  gather_rows_kernel<<<(unsigned)(((long long)n * 16 + 255) / 256), 256, 0, as_stream(stream)>>>(
      table, ids, n, out);
  MACR_LAUNCH_CHECK();
  return MACR_OK;
}

extern "C" size_t macr_score_topk_workspace_bytes(int T, int64_t n_items, int K) {
  if (T <= 0 || n_items < 0 || K <= 0) return 16;
  const int chunks = pick_chunks(T, n_items);
  const size_t Tpad = ((size_t)T + kTU - 1) / kTU * kTU;  // the overflow-queue shape may use the pad
  return (size_t)chunks * Tpad * K * (sizeof(int32_t) + sizeof(float)) + 256;
}

static int score_smem_opt_in() {
  static bool done = false;
  if (!done) {
    MACR_CUDA(cudaFuncSetAttribute(score_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                   (int)sizeof(ScoreSmem)));
    MACR_CUDA(cudaFuncSetAttribute(score_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                   (int)sizeof(ScoreSmem)));
    done = true;
  }
  return MACR_OK;
}

extern "C" int macr_score_topk(const float *Uq, int T, const float *It, int64_t n_items, int d,
                               const float *sig_i, const float *sig_u, float c,
                               const int32_t *mask_rowptr, const int32_t *mask_col, int K,
                               int32_t item_id_offset, int32_t *out_ids, float *out_scores,
                               void *ws, size_t ws_bytes, macr_stream_t stream) {
  MACR_CHECK_ARG(d == kD, "macr_score_topk: d must be %d (got %d)", kD, d);
  MACR_CHECK_ARG(K >= 1 && K <= kMaxKFast, "macr_score_topk: K must be in [1,%d] (got %d)",
                 kMaxKFast, K);
  MACR_CHECK_ARG(T >= 0 && n_items >= 0, "macr_score_topk: negative size");
  if (T == 0) return MACR_OK;
  MACR_CHECK_ARG(out_ids && out_scores, "macr_score_topk: null output");
  cudaStream_t s = as_stream(stream);
  // an empty shard (n_items == 0) runs zero tiles and emits padded lists (-1 / -inf)
  MACR_CHECK_ARG(Uq && sig_u && ws && (n_items == 0 || (It && sig_i)),
                 "macr_score_topk: null pointer");
  const size_t need = macr_score_topk_workspace_bytes(T, n_items, K);
  if (ws_bytes < need)
    return fail(MACR_ERR_WORKSPACE, "macr_score_topk: workspace %zu < %zu bytes", ws_bytes, need);
  int rc = score_smem_opt_in();
  if (rc) return rc;
  const int chunks = pick_chunks(T, n_items);
  const long long itiles = (n_items + kTN - 1) / kTN;
  const long long chunk_items = ((itiles + chunks - 1) / chunks) * kTN;
  int32_t *pids = reinterpret_cast<int32_t *>(ws);
  float *psc = reinterpret_cast<float *>(pids + (size_t)chunks * T * K);
  dim3 grid((T + kTU - 1) / kTU, chunks);
  score_kernel<true><<<grid, 256, sizeof(ScoreSmem), s>>>(Uq, T, It, n_items, sig_i, sig_u, c,
                                                          mask_rowptr, mask_col, K, item_id_offset,
                                                          chunk_items, pids, psc, nullptr,
                                                          nullptr, nullptr);
  MACR_LAUNCH_CHECK();
  topk_merge_kernel<<<(T + 7) / 8, 256, 0, s>>>(pids, psc, T, K, chunks, out_ids, out_scores,
                                                nullptr, nullptr, 0);
  MACR_LAUNCH_CHECK();
  return MACR_OK;
}

namespace macr {
size_t score_exact_workspace_bytes(int T, long long n_items, int K) {
  return macr_score_topk_workspace_bytes(T, n_items, K);
}

int score_exact_rows(const float *Uq, int T, const float *It, long long n_items,
                     const float *sig_i, const float *sig_u, float c, const int32_t *mask_rowptr,
                     const int32_t *mask_col, int K, int id_off, const int32_t *row_map,
                     const int *n_rows_dev, int32_t *out_ids, float *out_scores, void *ws,
                     size_t ws_bytes, cudaStream_t s) {
  const size_t need = score_exact_workspace_bytes(T, n_items, K);
  if (ws_bytes < need)
    return fail(MACR_ERR_WORKSPACE, "score_exact_rows: workspace %zu < %zu bytes", ws_bytes, need);
  int rc = score_smem_opt_in();
  if (rc) return rc;
  const int chunks = pick_chunks(T, n_items);
  const long long itiles = (n_items + kTN - 1) / kTN;
  const long long chunk_items = ((itiles + chunks - 1) / chunks) * kTN;
  int32_t *pids = reinterpret_cast<int32_t *>(ws);
  const size_t Tpad = ((size_t)T + kTU - 1) / kTU * kTU;
  float *psc = reinterpret_cast<float *>(pids + (size_t)chunks * Tpad * K);
  dim3 grid((T + kTU - 1) / kTU, chunks);
  // partial lists are indexed [chunk][slot][K]; with a row queue the pitch and the number of
  // chunks are the queue's own (queue_shape), within chunks * Tpad lists
  score_kernel<true><<<grid, 256, sizeof(ScoreSmem), s>>>(Uq, T, It, n_items, sig_i, sig_u, c,
                                                          mask_rowptr, mask_col, K, id_off,
                                                          chunk_items, pids, psc, nullptr, row_map,
                                                          n_rows_dev);
  MACR_LAUNCH_CHECK();
  topk_merge_kernel<<<(T + 7) / 8, 256, 0, s>>>(pids, psc, T, K, chunks, out_ids, out_scores,
                                                row_map, n_rows_dev, itiles);
  MACR_LAUNCH_CHECK();
  return MACR_OK;
}
}  // namespace macr

extern "C" int macr_score_matrix(const float *Uq, int T, const float *It, int64_t n_items, int d,
                                 const float *sig_i, const float *sig_u, float c, float *out,
                                 macr_stream_t stream) {
  MACR_CHECK_ARG(d == kD, "macr_score_matrix: d must be %d (got %d)", kD, d);
  MACR_CHECK_ARG(T >= 0 && n_items >= 0, "macr_score_matrix: negative size");
  if (T == 0 || n_items == 0) return MACR_OK;
  MACR_CHECK_ARG(Uq && It && sig_i && sig_u && out, "macr_score_matrix: null pointer");
  int rc = score_smem_opt_in();
  if (rc) return rc;
  const int chunks = pick_chunks(T, n_items);
  const long long itiles = (n_items + kTN - 1) / kTN;
  const long long chunk_items = ((itiles + chunks - 1) / chunks) * kTN;
  dim3 grid((T + kTU - 1) / kTU, chunks);
  score_kernel<false><<<grid, 256, sizeof(ScoreSmem), as_stream(stream)>>>(
      Uq, T, It, n_items, sig_i, sig_u, c, nullptr, nullptr, 1, 0, chunk_items, nullptr, nullptr,
      out, nullptr, nullptr);
  MACR_LAUNCH_CHECK();
  return MACR_OK;
}

extern "C" int macr_topk_merge(const int32_t *ids, const float *scores, int T, int K, int G,
                               int32_t *out_ids, float *out_scores, macr_stream_t stream) {
  MACR_CHECK_ARG(K >= 1 && K <= kMaxKFast, "macr_topk_merge: K must be in [1,%d] (got %d)",
                 kMaxKFast, K);
  MACR_CHECK_ARG(T >= 0 && G >= 1, "macr_topk_merge: bad T/G");
  if (T == 0) return MACR_OK;
  MACR_CHECK_ARG(ids && scores && out_ids && out_scores, "macr_topk_merge: null pointer");
  topk_merge_kernel<<<(T + 7) / 8, 256, 0, as_stream(stream)>>>(ids, scores, T, K, G, out_ids,
                                                                out_scores, nullptr, nullptr, 0);
  MACR_LAUNCH_CHECK();
  return MACR_OK;
}

extern "C" int macr_topk_rows(const float *scores, int columns_num, int rows_num, int top_k,
                              int32_t *rankings, macr_stream_t stream) {
  MACR_CHECK_ARG(top_k >= 1 && top_k <= kMaxKFast, "macr_topk_rows: top_k must be in [1,%d] (got %d)",
                 kMaxKFast, top_k);
  MACR_CHECK_ARG(columns_num >= 0 && rows_num >= 0, "macr_topk_rows: negative size");
  if (rows_num == 0) return MACR_OK;
  MACR_CHECK_ARG(scores && rankings, "macr_topk_rows: null pointer");
  topk_rows_kernel<<<(rows_num + 7) / 8, 256, 0, as_stream(stream)>>>(scores, columns_num, rows_num,
                                                                      top_k, rankings);
  MACR_LAUNCH_CHECK();
  return MACR_OK;
}

// ---------------------------------------------------------------------------------------------
// K9: fold-out metric curves (evaluate_foldout.h:16-113), one thread per user
// ---------------------------------------------------------------------------------------------
namespace macr {
__global__ void __launch_bounds__(128)
foldout_metrics_kernel(const int32_t *__restrict__ topk, int T, int K,
                       const int32_t *__restrict__ truth_rowptr,
                       const int32_t *__restrict__ truth_col, const double *__restrict__ inv_log2,
                       float *__restrict__ out) {
  const int t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= T) return;
  const int32_t *rank = topk + (long long)t * K;
  const int lo = truth_rowptr[t], tl = truth_rowptr[t + 1] - lo;
  float *o = out + (long long)t * 5 * K;
  int hits = 0;
  bool found = false;
  float sum_pre = 0.f, DCG = 0.f, iDCG = 0.f, rr = 0.f;
  for (int i = 0; i < K; ++i) {
    const int id = rank[i];
    bool hit = false;
    for (int q = 0; q < tl; ++q) hit |= (truth_col[lo + q] == id);
    if (hit) {
      hits += 1;
      const float pre = (float)(1.0 * hits / (double)(i + 1));
      sum_pre = __fadd_rn(sum_pre, pre);
      DCG = (float)((double)DCG + inv_log2[i]);
      if (!found) {
        found = true;
        rr = (float)(1.0 / (double)(i + 1));
      }
    }
    if (i < tl) iDCG = (float)((double)iDCG + inv_log2[i]);
    o[0 * K + i] = (float)(1.0 * hits / (double)(i + 1));
    o[1 * K + i] = (float)(1.0 * hits / (double)tl);
    o[2 * K + i] = __fdiv_rn(sum_pre, (float)tl);
    o[3 * K + i] = __fdiv_rn(DCG, iDCG);
    o[4 * K + i] = rr;
  }
}
}  // namespace macr

extern "C" int macr_foldout_metrics(const int32_t *topk_ids, int T, int K,
                                    const int32_t *truth_rowptr, const int32_t *truth_col,
                                    const double *inv_log2, float *out, macr_stream_t stream) {
  MACR_CHECK_ARG(T >= 0 && K >= 1, "macr_foldout_metrics: bad T/K");
  if (T == 0) return MACR_OK;
  MACR_CHECK_ARG(topk_ids && truth_rowptr && truth_col && inv_log2 && out,
                 "macr_foldout_metrics: null pointer");
  foldout_metrics_kernel<<<(T + 127) / 128, 128, 0, as_stream(stream)>>>(
      topk_ids, T, K, truth_rowptr, truth_col, inv_log2, out);
  MACR_LAUNCH_CHECK();
  return MACR_OK;
}
