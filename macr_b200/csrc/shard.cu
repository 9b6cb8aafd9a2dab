// shard.cu -- row-partitioned tables over the GPUs of one box: the exchange of the
// batch's rows between the ranks, fused with the id renumbering.
//
// Every rank owns a contiguous id range of both embedding tables (and of their Adam slots).  A
// training step needs the 3B rows of the batch on every rank; every row has exactly one owner.
// Behind the owned rows each local table carries GHOST rows, two parities of them:
//
//     U_local : [ n_local_u owned rows | parity 0: maxB ghosts  | parity 1: maxB ghosts  ]
//     I_local : [ n_local_i owned rows | parity 0: 2maxB ghosts | parity 1: 2maxB ghosts ]
//
// Batch position b of the user column lives in ghost `b`, of the pos column in ghost `b`, of the
// neg column in ghost `B + b`.  The step graph then runs on local ids unchanged.
//
// Two transports:
//   * push (NVLink peer stores): ONE kernel renumbers the ids and stores each owned row straight
//     into the ghost slot of every peer's table (pointers from cudaIpcOpenMemHandle), then a
//     one-warp flag barrier over peer memory (st.release.sys / ld.acquire.sys).  No staging
//     buffer, no reduction, no copy: gather + all-gather in one pass over the owned rows.  Ghost
//     parities alternate per step, so a fast peer may already push step t+1 while this rank still
//     runs step t (it cannot reach t+2 before this rank has passed barrier t+1).
//   * pack (for an NCCL all-reduce by the caller): owned rows -> ex[3B][64], zeros elsewhere;
//     after the sum macr_shard_unpack copies ex into the ghost slots.
#include "score.cuh"
#include "shard.cuh"
#include "train_kernels.cuh"

namespace macr {

typedef PeerGhosts PeerTabs;  // ghost base (row n_local of rank r's table) as mapped in THIS process
struct PeerFlags {
  unsigned long long *p[kMaxRanks];  // rank r's flag array [world]
};

template <bool PUSH>
__global__ void __launch_bounds__(256)
shard_exchange_kernel(const float *__restrict__ U, const float *__restrict__ I, macr_shard_desc a,
                      const int32_t *__restrict__ ids3, int B, int parity,
                      int32_t *__restrict__ local3, float *__restrict__ ex, PeerTabs peers,
                      const StepState *__restrict__ st) {
  const int q = (blockIdx.x * blockDim.x + threadIdx.x) >> 4, hl = threadIdx.x & 15;
  if (q >= 3 * B) return;
  if (st) {  // inside the step graph: this step's slot of the staged epoch, parity from the Adam step
    ids3 = st->gids_base + st->step_idx * 3LL * B;
    local3 = const_cast<int32_t *>(st->ids_base) + st->step_idx * 3LL * B;
    parity = (int)(st->t & 1);
  }
  const bool is_user = q < B;
  const long long id = ids3[q];
  const long long lo = is_user ? a.u_lo : a.i_lo, hi = is_user ? a.u_hi : a.i_hi;
  const bool own = id >= lo && id < hi;
  const long long ghost = is_user ? (long long)parity * a.max_batch + q
                                  : (long long)parity * 2 * a.max_batch + (q - B);
  if (hl == 0) local3[q] = (int32_t)(own ? id - lo : (hi - lo) + ghost);
  const float4 *src = reinterpret_cast<const float4 *>((is_user ? U : I) + (id - lo) * kD) + hl;
  if (PUSH) {
    if (!own) return;
    const float4 v = *src;
#pragma unroll 1
    for (int r = 0; r < a.world; ++r) {
      if (r == a.rank) continue;
      float *dst = is_user ? peers.u[r] : peers.i[r];
      st_stream(reinterpret_cast<float4 *>(dst + ghost * kD) + hl, v);
    }
    // the peer stores of this thread are performed system-wide before it retires: the flag of
    // the barrier kernel that follows in the stream is then never seen ahead of the rows
    __threadfence_system();
  } else {
    reinterpret_cast<float4 *>(ex + (long long)q * kD)[hl] = own ? *src : make_float4(0.f, 0.f, 0.f, 0.f);
  }
}

// one warp: lane r signals peer r (flags_r[me] = epoch) and waits for peer r (flags_me[r] >= epoch).
// A peer that never arrives raises *err instead of hanging the GPU.
__global__ void shard_barrier_kernel(PeerFlags f, int rank, int world, unsigned long long epoch,
                                     long long spin_limit, int *err) {
  const int r = threadIdx.x;
  if (r >= world || r == rank) return;
  __threadfence_system();
  asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(f.p[r] + rank), "l"(epoch) : "memory");
  const unsigned long long *mine = f.p[rank] + r;
  const long long t0 = clock64();
  for (;;) {
    unsigned long long v;
    asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(mine) : "memory");
    if (v >= epoch) break;
    if (clock64() - t0 > spin_limit) {
      atomicExch(err, 1 + r);
      break;
    }
  }
}

int launch_shard_push_st(const float *U_local, const float *I_local, const macr_shard_desc &desc,
                         const StepState *st, int B, const PeerGhosts &peers, cudaStream_t s) {
  shard_exchange_kernel<true><<<(unsigned)((3LL * B * 16 + 255) / 256), 256, 0, s>>>(
      U_local, I_local, desc, nullptr, B, 0, nullptr, nullptr, peers, st);
  MACR_LAUNCH_CHECK();
  return MACR_OK;
}

// ---- all-gather of row ranges by peer stores (row-partitioned LightGCN) --------
__global__ void __launch_bounds__(256)
peer_push_kernel(PeerPush p) {
  const long long n4a = p.rows[0] * (kD / 4), n4 = n4a + p.rows[1] * (kD / 4);
  for (long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x; e < n4;
       e += (long long)gridDim.x * blockDim.x) {
    const int k = e >= n4a;
    const long long off = k ? e - n4a : e;
    const float4 v = ld_stream(reinterpret_cast<const float4 *>(p.src[k]) + off);
#pragma unroll 1
    for (int r = 0; r < p.world; ++r)
      if (r != p.rank) st_stream(reinterpret_cast<float4 *>(p.dst[k][r]) + off, v);
  }
  __threadfence_system();  // see shard_exchange_kernel
}

int launch_peer_push(const PeerPush &p, cudaStream_t s) {
  const long long n4 = (p.rows[0] + p.rows[1]) * (kD / 4);
  if (n4 == 0 || p.world <= 1) return MACR_OK;
  long long ctas = (n4 + 255) / 256;
  const long long cap = (long long)sm_count() * 8;
  if (ctas > cap) ctas = cap;
  peer_push_kernel<<<(unsigned)ctas, 256, 0, s>>>(p);
  MACR_LAUNCH_CHECK();
  return MACR_OK;
}

__global__ void __launch_bounds__(256)
peer_push_batch_rows_kernel(const float *__restrict__ buf, PeerBufs peers, macr_shard_desc d,
                            long long n_users, const StepState *__restrict__ st, int B) {
  const int q = (blockIdx.x * blockDim.x + threadIdx.x) >> 4, hl = threadIdx.x & 15;
  if (q >= 3 * B) return;
  const long long id = (st->ids_base + st->step_idx * 3LL * B)[q];
  const bool is_user = q < B;
  const bool own = is_user ? (id >= d.u_lo && id < d.u_hi) : (id >= d.i_lo && id < d.i_hi);
  if (!own) return;
  const long long off = (id + (is_user ? 0 : n_users)) * kD;
  const float4 v = reinterpret_cast<const float4 *>(buf + off)[hl];
#pragma unroll 1
  for (int r = 0; r < d.world; ++r)
    if (r != d.rank) st_stream(reinterpret_cast<float4 *>(peers.p[r] + off) + hl, v);
  __threadfence_system();  // see shard_exchange_kernel
}

int launch_peer_push_batch_rows(const float *buf, const PeerBufs &peers, const macr_shard_desc &desc,
                                long long n_users, const StepState *st, int B, cudaStream_t s) {
  if (desc.world <= 1) return MACR_OK;
  peer_push_batch_rows_kernel<<<(unsigned)((3LL * B * 16 + 255) / 256), 256, 0, s>>>(buf, peers, desc,
                                                                                      n_users, st, B);
  MACR_LAUNCH_CHECK();
  return MACR_OK;
}

__global__ void peer_barrier_dev_kernel(PeerFlagsDev f, unsigned long long *epoch_ctr, int rank,
                                        int world, long long spin_limit, int *err) {
  const int r = threadIdx.x;
  unsigned long long epoch = 0;
  if (r == 0) epoch = *epoch_ctr + 1;
  epoch = __shfl_sync(0xffffffffu, epoch, 0);
  if (r < world && r != rank) {
    __threadfence_system();
    asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(f.p[r] + rank), "l"(epoch) : "memory");
    const unsigned long long *mine = f.p[rank] + r;
    const long long t0 = clock64();
    for (;;) {
      unsigned long long v;
      asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(mine) : "memory");
      if (v >= epoch) break;
      if (clock64() - t0 > spin_limit) {
        atomicExch(err, 1 + r);
        break;
      }
    }
  }
  __syncwarp();
  if (r == 0) *epoch_ctr = epoch;
}

int launch_peer_barrier_dev(const PeerFlagsDev &f, unsigned long long *epoch_ctr, int *err, int rank,
                            int world, cudaStream_t s) {
  if (world <= 1) return MACR_OK;
  peer_barrier_dev_kernel<<<1, 32, 0, s>>>(f, epoch_ctr, rank, world, 20000000000LL, err);
  MACR_LAUNCH_CHECK();
  return MACR_OK;
}

// ---- item-partitioned scoring: exchange + merge of the shards' candidates in ONE kernel ----------
// Rank j owns the query rows [row0, row0 + rows).  For each of them a warp reads the K candidates of
// every shard straight from that shard's buffer (peer loads over NVLink, shard order = rank order,
// the order rule of topk_merge_kernel), merges them and stores the merged row into the result
// buffer of EVERY rank (peer stores): all-to-all, K-way merge and all-gather without a staging
// buffer.  The caller brackets it with two flag barriers (candidates complete / results landed).
struct MergePeers {
  const int32_t *ids[kMaxRanks];
  const float *sc[kMaxRanks];
  int32_t *out_ids[kMaxRanks];
  float *out_sc[kMaxRanks];
  int world;
};

__global__ void __launch_bounds__(256)
topk_merge_peers_kernel(MergePeers p, int row0, int rows, int K) {
  const int lane = threadIdx.x & 31;
  const int w = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (w >= rows) return;
  const long long base = (long long)(row0 + w) * K;
  const unsigned kmask = (K >= 32) ? 0xffffffffu : ((1u << K) - 1u);
  float ls = -INFINITY;
  int li = 0x7fffffff;
  float s[kMaxRanks];
  int id[kMaxRanks];
#pragma unroll
  for (int g = 0; g < kMaxRanks; ++g) {  // all peer loads in flight before the first merge step
    s[g] = -INFINITY;
    id[g] = -1;
    if (g < p.world && lane < K) {
      s[g] = p.sc[g][base + lane];
      id[g] = p.ids[g][base + lane];
    }
  }
#pragma unroll
  for (int g = 0; g < kMaxRanks; ++g) {
    if (g >= p.world) break;
    for (int k = 0; k < K; ++k) {
      const float cs = __shfl_sync(0xffffffffu, s[g], k);
      const int cid = __shfl_sync(0xffffffffu, id[g], k);
      if (cid < 0) break;  // lists are padded at the tail
      const float ws = __shfl_sync(0xffffffffu, ls, K - 1);
      const int wi = __shfl_sync(0xffffffffu, li, K - 1);
      if (!score_better(cs, cid, ws, wi)) break;  // sorted input: the rest of this list loses too
      score_list_insert(ls, li, cs, cid, lane, kmask, K);
    }
  }
  if (lane < K) {
    const bool empty = li == 0x7fffffff;
    const int oi = empty ? -1 : li;
    const float os = empty ? -INFINITY : ls;
#pragma unroll 1
    for (int r = 0; r < p.world; ++r) {
      p.out_ids[r][base + lane] = oi;
      p.out_sc[r][base + lane] = os;
    }
  }
  __threadfence_system();  // see shard_exchange_kernel
}

static int check_desc(const macr_shard_desc *d, int B, const char *who) {
  MACR_CHECK_ARG(d, "%s: null descriptor", who);
  MACR_CHECK_ARG(d->world >= 1 && d->world <= kMaxRanks && d->rank >= 0 && d->rank < d->world,
                 "%s: rank %d / world %d outside [0,%d]", who, d->rank, d->world, kMaxRanks);
  MACR_CHECK_ARG(d->u_lo <= d->u_hi && d->i_lo <= d->i_hi && d->u_lo >= 0 && d->i_lo >= 0,
                 "%s: bad id ranges", who);
  MACR_CHECK_ARG(d->max_batch > 0 && B > 0 && B <= d->max_batch, "%s: batch %d outside (0,%d]", who, B,
                 d->max_batch);
  return MACR_OK;
}

}  // namespace macr

using namespace macr;

extern "C" int macr_shard_pack(const float *U_local, const float *I_local, const macr_shard_desc *desc,
                               const int32_t *ids3, int B, int parity, int32_t *local_ids3, float *ex,
                               macr_stream_t stream) {
  int rc = check_desc(desc, B, "macr_shard_pack");
  if (rc) return rc;
  MACR_CHECK_ARG(U_local && I_local && ids3 && local_ids3 && ex, "macr_shard_pack: null pointer");
  MACR_CHECK_ARG(parity == 0 || parity == 1, "macr_shard_pack: parity must be 0 or 1");
  PeerTabs none{};
  shard_exchange_kernel<false><<<(unsigned)((3LL * B * 16 + 255) / 256), 256, 0, as_stream(stream)>>>(
      U_local, I_local, *desc, ids3, B, parity, local_ids3, ex, none, nullptr);
  MACR_LAUNCH_CHECK();
  return MACR_OK;
}

extern "C" int macr_shard_unpack(float *U_local, float *I_local, const macr_shard_desc *desc,
                                 const float *ex, int B, int parity, macr_stream_t stream) {
  int rc = check_desc(desc, B, "macr_shard_unpack");
  if (rc) return rc;
  MACR_CHECK_ARG(U_local && I_local && ex, "macr_shard_unpack: null pointer");
  MACR_CHECK_ARG(parity == 0 || parity == 1, "macr_shard_unpack: parity must be 0 or 1");
  const long long mb = desc->max_batch;
  float *ug = U_local + ((desc->u_hi - desc->u_lo) + parity * mb) * kD;
  float *ig = I_local + ((desc->i_hi - desc->i_lo) + parity * 2 * mb) * kD;
  cudaStream_t s = as_stream(stream);
  MACR_CUDA(cudaMemcpyAsync(ug, ex, sizeof(float) * kD * (size_t)B, cudaMemcpyDeviceToDevice, s));
  MACR_CUDA(cudaMemcpyAsync(ig, ex + (size_t)B * kD, sizeof(float) * kD * 2 * (size_t)B,
                            cudaMemcpyDeviceToDevice, s));
  return MACR_OK;
}

extern "C" int macr_shard_push(const float *U_local, const float *I_local, const macr_shard_desc *desc,
                               const int32_t *ids3, int B, int parity, int32_t *local_ids3,
                               float *const *peer_U_ghost_host, float *const *peer_I_ghost_host,
                               macr_stream_t stream) {
  int rc = check_desc(desc, B, "macr_shard_push");
  if (rc) return rc;
  MACR_CHECK_ARG(U_local && I_local && ids3 && local_ids3 && peer_U_ghost_host && peer_I_ghost_host,
                 "macr_shard_push: null pointer");
  MACR_CHECK_ARG(parity == 0 || parity == 1, "macr_shard_push: parity must be 0 or 1");
  PeerTabs peers{};
  for (int r = 0; r < desc->world; ++r) {
    peers.u[r] = peer_U_ghost_host[r];
    peers.i[r] = peer_I_ghost_host[r];
    MACR_CHECK_ARG(r == desc->rank || (peers.u[r] && peers.i[r]), "macr_shard_push: null peer %d", r);
  }
  shard_exchange_kernel<true><<<(unsigned)((3LL * B * 16 + 255) / 256), 256, 0, as_stream(stream)>>>(
      U_local, I_local, *desc, ids3, B, parity, local_ids3, nullptr, peers, nullptr);
  MACR_LAUNCH_CHECK();
  return MACR_OK;
}

extern "C" int macr_shard_barrier(uint64_t *const *peer_flags_host, int rank, int world, uint64_t epoch,
                                  int *err_flag, macr_stream_t stream) {
  MACR_CHECK_ARG(peer_flags_host && err_flag, "macr_shard_barrier: null pointer");
  MACR_CHECK_ARG(world >= 1 && world <= kMaxRanks && rank >= 0 && rank < world,
                 "macr_shard_barrier: rank %d / world %d", rank, world);
  if (world == 1) return MACR_OK;
  PeerFlags f{};
  for (int r = 0; r < world; ++r) {
    MACR_CHECK_ARG(peer_flags_host[r], "macr_shard_barrier: null flags of rank %d", r);
    f.p[r] = reinterpret_cast<unsigned long long *>(peer_flags_host[r]);
  }
  // ~10 s at 2 GHz: ranks are host-synchronised before the first step, so a longer wait means a
  // dead peer; the step then reports it through macr_shard_check instead of hanging the device
  shard_barrier_kernel<<<1, 32, 0, as_stream(stream)>>>(f, rank, world, (unsigned long long)epoch,
                                                        20000000000LL, err_flag);
  MACR_LAUNCH_CHECK();
  return MACR_OK;
}

// ---- peer-mappable device memory (CUDA IPC) ---------------------------------------------------
extern "C" int macr_ipc_alloc(size_t bytes, void **dev_ptr, unsigned char handle_out[MACR_IPC_HANDLE_BYTES]) {
  MACR_CHECK_ARG(dev_ptr && handle_out && bytes > 0, "macr_ipc_alloc: bad argument");
  static_assert(sizeof(cudaIpcMemHandle_t) == MACR_IPC_HANDLE_BYTES, "IPC handle size");
  void *p = nullptr;
  MACR_CUDA(cudaMalloc(&p, bytes));
  cudaIpcMemHandle_t h;
  cudaError_t e = cudaIpcGetMemHandle(&h, p);
  if (e != cudaSuccess) {
    cudaFree(p);
    return fail(MACR_ERR_CUDA, "cudaIpcGetMemHandle: %s", cudaGetErrorString(e));
  }
  memcpy(handle_out, &h, sizeof(h));
  *dev_ptr = p;
  return MACR_OK;
}

extern "C" int macr_ipc_open(const unsigned char handle[MACR_IPC_HANDLE_BYTES], void **peer_ptr) {
  MACR_CHECK_ARG(handle && peer_ptr, "macr_ipc_open: null argument");
  cudaIpcMemHandle_t h;
  memcpy(&h, handle, sizeof(h));
  MACR_CUDA(cudaIpcOpenMemHandle(peer_ptr, h, cudaIpcMemLazyEnablePeerAccess));
  return MACR_OK;
}

extern "C" int macr_ipc_close(void *peer_ptr) {
  if (peer_ptr) MACR_CUDA(cudaIpcCloseMemHandle(peer_ptr));
  return MACR_OK;
}

extern "C" int macr_ipc_free(void *dev_ptr) {
  if (dev_ptr) MACR_CUDA(cudaFree(dev_ptr));
  return MACR_OK;
}

extern "C" int macr_topk_merge_peers(const int32_t *const *cand_ids_host, const float *const *cand_scores_host,
                                     int32_t *const *out_ids_host, float *const *out_scores_host, int world,
                                     int K, int row0, int rows, macr_stream_t stream) {
  MACR_CHECK_ARG(cand_ids_host && cand_scores_host && out_ids_host && out_scores_host,
                 "macr_topk_merge_peers: null pointer table");
  MACR_CHECK_ARG(world >= 1 && world <= kMaxRanks, "macr_topk_merge_peers: world %d outside [1,%d]", world,
                 kMaxRanks);
  MACR_CHECK_ARG(K >= 1 && K <= 32, "macr_topk_merge_peers: K must be in [1,32] (got %d)", K);
  MACR_CHECK_ARG(row0 >= 0 && rows >= 0, "macr_topk_merge_peers: negative row range");
  if (rows == 0) return MACR_OK;
  MergePeers p{};
  p.world = world;
  for (int r = 0; r < world; ++r) {
    MACR_CHECK_ARG(cand_ids_host[r] && cand_scores_host[r] && out_ids_host[r] && out_scores_host[r],
                   "macr_topk_merge_peers: null pointer (rank %d)", r);
    p.ids[r] = cand_ids_host[r];
    p.sc[r] = cand_scores_host[r];
    p.out_ids[r] = out_ids_host[r];
    p.out_sc[r] = out_scores_host[r];
  }
  topk_merge_peers_kernel<<<(rows + 7) / 8, 256, 0, as_stream(stream)>>>(p, row0, rows, K);
  MACR_LAUNCH_CHECK();
  return MACR_OK;
}
