// shard.cuh -- internal declarations of the peer-memory exchange kernels (shard.cu)
#pragma once
#include "common.cuh"

namespace macr {

constexpr int kMaxRanks = MACR_SHARD_MAX_RANKS;

// all-gather by peer stores: up to two contiguous row ranges of a local [rows][64] buffer are
// copied to the same rows of every peer's buffer (dst[k][r] = address in THIS process of range k's
// first row in rank r's buffer; dst[k][rank] is ignored)
struct PeerPush {
  const float *src[2];
  long long rows[2];
  float *dst[2][kMaxRanks];
  int rank, world;
};
int launch_peer_push(const PeerPush &p, cudaStream_t s);

// row-partitioned MF step, exchange inside the step graph: ids of the step being executed are
// read from st->gids_base, renumbered into st->ids_base (same step slot) and every owned row is
// stored into the ghost slot of every peer (parity = st->t & 1); see shard.cu
struct PeerGhosts {
  float *u[kMaxRanks];  // row n_local of rank r's local user table, mapped into this process
  float *i[kMaxRanks];
};
struct StepState;
int launch_shard_push_st(const float *U_local, const float *I_local, const macr_shard_desc &desc,
                         const StepState *st, int B, const PeerGhosts &peers, cudaStream_t s);

// the batch's rows only: for each of the step's 3B ids (users | pos | neg, read from the step state)
// whose node row this rank owns, the row of `buf` ([N][64], node rows = users then items) is stored
// to the same row of every peer's buffer -- what a training step reads of the layer mean
struct PeerBufs {
  float *p[kMaxRanks];
};
int launch_peer_push_batch_rows(const float *buf, const PeerBufs &peers, const macr_shard_desc &desc,
                                long long n_users, const StepState *st, int B, cudaStream_t s);

// flag barrier over peer memory whose epoch lives on the device (so it can sit in a CUDA graph):
// *epoch_ctr is incremented by one per call; flags[r] = rank r's uint64[kMaxRanks] arrival flags
struct PeerFlagsDev {
  unsigned long long *p[kMaxRanks];
};
int launch_peer_barrier_dev(const PeerFlagsDev &f, unsigned long long *epoch_ctr, int *err, int rank,
                            int world, cudaStream_t s);

}  // namespace macr
