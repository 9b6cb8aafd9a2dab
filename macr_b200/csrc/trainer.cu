// trainer.cu -- step handles: the whole `sess.run([opt_two_bce_both, loss..., mf..., reg...])`
// of macr_mf/train.py:492-496 and macr_lightgcn/LightGCN.py:598-607 behind one C call.
//
// A handle owns scratch and one captured CUDA graph per batch size.  Nothing in the graph depends
// on per-step host values: the Adam step state (beta powers), the current batch pointer and the
// loss destination live in a device-resident StepState that the last kernel of every step
// advances, so an epoch of pre-staged batches replays the same graph back to back.
//
// MF step DAG (two captured streams):
//     gather_dots -> grid(BxB) -> finalize ----------------+
//        \-> batch_plan(users|items) -> adam_sweep(untouched rows, HBM-bound)  --+-> row_grads
//                                                      -> adam_rows -> adam_vec(w,w_user) -> losses
// The sweep does not depend on the gradients, so it overlaps the MUFU-bound grid.
#include <stdlib.h>

#include <map>
#include <new>

#include "shard.cuh"
#include "spmm.cuh"

namespace macr {

__global__ void set_io_kernel(StepState *st, const int32_t *ids_base, float *loss_base,
                              const int32_t *gids_base) {
  if (threadIdx.x == 0 && blockIdx.x == 0) {
    st->ids_base = ids_base;
    st->gids_base = gids_base;
    st->loss_base = loss_base;
    st->step_idx = 0;
  }
}
__global__ void set_powers_kernel(StepState *st, float b1p, float b2p, long long t) {
  if (threadIdx.x == 0 && blockIdx.x == 0) {
    st->b1p = b1p;
    st->b2p = b2p;
    st->t = t;
  }
}

struct TrainerBase {
  float *U, *mU, *vU, *I, *mI, *vI, *w, *mw, *vw, *wu, *mwu, *vwu;
  int64_t nu, ni;
  int maxB;
  macr_hparams hp;
  cudaStream_t s;         // caller's stream: graphs are launched here
  cudaStream_t cs, side, side2;  // private streams the step DAG is captured on (the legacy default
                          // stream cannot be captured)
  cudaEvent_t ev_fork, ev_join, ev_join2;
  StepState *st;
  int32_t *ids_stage;
  float *loss_stage;
  float *scal;  // 11 x maxB: yp yn sp sn su regsq dyp dyn dsp dsn dsu
  void *gridws;
  PlanBufs planU, planI;
  int32_t *plan_mem;
  float *gU, *gI, *gw_part, *gwu_part, *unit_part;
  float *snap;          // [3][maxB][64] row snapshot of the step (users | pos | neg)
  unsigned *tail_ticket;
  float *pinned_losses;
  int64_t steps_done;
  const int32_t *cur_ids_base;
  float *cur_loss_base;
  bool single_mode_set;
  // epoch staging of *_run_host: device copies of the host batches / losses, grown on demand
  int32_t *epoch_ids = nullptr;
  float *epoch_losses = nullptr, *epoch_losses_pinned = nullptr;
  size_t epoch_ids_cap = 0, epoch_loss_cap = 0;
  std::map<int, cudaGraphExec_t> graphs;       // train step, keyed by B
  std::map<int, cudaGraphExec_t> graphs_eval;  // loss-only step (LightGCN test loss)

  float *sc(int k, int) const { return scal + (size_t)k * maxB; }

  int alloc_common(cudaStream_t stream) {
    s = stream;
    int rci = plan_init();
    if (rci) return rci;
    MACR_CUDA(cudaStreamCreateWithFlags(&cs, cudaStreamNonBlocking));
    MACR_CUDA(cudaStreamCreateWithFlags(&side, cudaStreamNonBlocking));
    MACR_CUDA(cudaStreamCreateWithFlags(&side2, cudaStreamNonBlocking));
    MACR_CUDA(cudaEventCreateWithFlags(&ev_join2, cudaEventDisableTiming));
    MACR_CUDA(cudaEventCreateWithFlags(&ev_fork, cudaEventDisableTiming));
    MACR_CUDA(cudaEventCreateWithFlags(&ev_join, cudaEventDisableTiming));
    MACR_CUDA(cudaMalloc(&st, sizeof(StepState)));
    MACR_CUDA(cudaMemset(st, 0, sizeof(StepState)));
    MACR_CUDA(cudaMalloc(&ids_stage, sizeof(int32_t) * 3 * (size_t)maxB));
    MACR_CUDA(cudaMalloc(&loss_stage, sizeof(float) * 4));
    MACR_CUDA(cudaMalloc(&scal, sizeof(float) * 11 * (size_t)maxB));
    MACR_CUDA(cudaMalloc(&gridws, grid_ws_layout(maxB, nullptr).bytes + 4096));
    MACR_CUDA(cudaMemset(gridws, 0xff, grid_ws_layout(maxB, nullptr).bytes + 4096));  // empty slots
    MACR_CUDA(cudaMalloc(&snap, sizeof(float) * kD * 3 * (size_t)maxB));
    MACR_CUDA(cudaMalloc(&tail_ticket, sizeof(unsigned) * 4));
    MACR_CUDA(cudaMemset(tail_ticket, 0, sizeof(unsigned) * 4));
    // plan buffers: users (B ids) + items (2B ids): uniq, seg_off(+1), seg_pos, n_uniq + scratch
    const size_t pm = (size_t)maxB * 3 + 8 + (size_t)maxB * 6 + 8;
    MACR_CUDA(cudaMalloc(&plan_mem, sizeof(int32_t) * pm + plan_ws_bytes(maxB) +
                                        plan_ws_bytes(2 * maxB)));
    MACR_CUDA(cudaMemset(plan_mem, 0, sizeof(int32_t) * pm + plan_ws_bytes(maxB) +
                                          plan_ws_bytes(2 * maxB)));
    {
      int32_t *p = plan_mem;
      int32_t *uU = p; p += maxB;
      int32_t *oU = p; p += maxB + 1;
      int32_t *sU = p; p += maxB;
      int32_t *nU = p; p += 1;
      int32_t *uI = p; p += 2 * maxB;
      int32_t *oI = p; p += 2 * maxB + 1;
      int32_t *sI = p; p += 2 * maxB;
      int32_t *nI = p; p += 1;
      p = plan_mem + pm;
      planU = plan_carve(uU, oU, sU, nU, p, maxB);
      planI = plan_carve(uI, oI, sI, nI, reinterpret_cast<char *>(p) + plan_ws_bytes(maxB),
                         2 * maxB);
    }
    MACR_CUDA(cudaMalloc(&unit_part, sizeof(float) * kD * 3 * (size_t)maxB));
    MACR_CUDA(cudaMalloc(&gU, sizeof(float) * kD * (size_t)maxB));
    MACR_CUDA(cudaMalloc(&gI, sizeof(float) * kD * 2 * (size_t)maxB));
    const int parts = row_grads_max_parts(maxB);
    MACR_CUDA(cudaMalloc(&gw_part, sizeof(float) * kD * (size_t)parts));
    MACR_CUDA(cudaMalloc(&gwu_part, sizeof(float) * kD * (size_t)parts));
    MACR_CUDA(cudaMallocHost(&pinned_losses, sizeof(float) * 4));
    steps_done = 0;
    cur_ids_base = nullptr;
    cur_loss_base = nullptr;
    single_mode_set = false;
    set_powers_kernel<<<1, 1, 0, s>>>(st, hp.beta1, hp.beta2, 0);
    MACR_LAUNCH_CHECK();
    return MACR_OK;
  }

  void free_common() {
    for (auto &kv : graphs) cudaGraphExecDestroy(kv.second);
    for (auto &kv : graphs_eval) cudaGraphExecDestroy(kv.second);
    cudaFree(st); cudaFree(ids_stage); cudaFree(loss_stage); cudaFree(scal); cudaFree(gridws);
    cudaFree(plan_mem); cudaFree(snap); cudaFree(tail_ticket); cudaFree(unit_part); cudaFree(gU); cudaFree(gI); cudaFree(gw_part); cudaFree(gwu_part);
    cudaFreeHost(pinned_losses);
    cudaFree(epoch_ids); cudaFree(epoch_losses); cudaFreeHost(epoch_losses_pinned);
    cudaFree(local_ids);
    cudaEventDestroy(ev_fork); cudaEventDestroy(ev_join);
    cudaEventDestroy(ev_join2);
    cudaStreamDestroy(side);
    cudaStreamDestroy(side2);
    cudaStreamDestroy(cs);
  }

  // row-partitioned MF (macr_mf_trainer_shard): the caller's ids are GLOBAL; the step graph's
  // exchange kernel renumbers each step's ids into this buffer, which the other kernels read
  bool renumber_ids = false;
  int32_t *local_ids = nullptr;
  size_t local_ids_cap = 0;

  int set_io(const int32_t *ids_base, float *loss_base, int n_steps, int B) {
    const int32_t *gids = nullptr;
    if (renumber_ids) {
      const size_t need = (size_t)n_steps * 3 * (size_t)B;
      if (need > local_ids_cap) {
        MACR_CUDA(cudaStreamSynchronize(s));
        cudaFree(local_ids);
        local_ids = nullptr;
        local_ids_cap = 0;
        const size_t cap = need > 2 * local_ids_cap ? need : 2 * local_ids_cap;  // geometric growth
        MACR_CUDA(cudaMalloc(&local_ids, sizeof(int32_t) * cap));
        local_ids_cap = cap;
      }
      gids = ids_base;
      ids_base = local_ids;
    }
    set_io_kernel<<<1, 1, 0, s>>>(st, ids_base, loss_base, gids);
    MACR_LAUNCH_CHECK();
    cur_ids_base = ids_base;
    cur_loss_base = loss_base;
    return MACR_OK;
  }

  int set_steps(int64_t t) {
    float b1p = hp.beta1, b2p = hp.beta2;  // fp32 repeated multiplication, like adam.py _finish
    for (int64_t k = 0; k < t; ++k) {
      b1p = b1p * hp.beta1;
      b2p = b2p * hp.beta2;
    }
    set_powers_kernel<<<1, 1, 0, s>>>(st, b1p, b2p, (long long)t);
    MACR_LAUNCH_CHECK();
    steps_done = t;
    return MACR_OK;
  }
};

// -------------------------------------------------------------------------------------------
// MF
// -------------------------------------------------------------------------------------------
}  // namespace macr

struct macr_mf_trainer : macr::TrainerBase {
  uint32_t *bmU, *bmI;
  int launches;
  int mode = MACR_TRAIN_RUBIBCEBOTH;
  // row-partitioned mode (macr_mf_trainer_shard): U / I are this rank's local tables (owned rows +
  // ghost rows, csrc/shard.cu); the exchange of the batch's rows runs inside the step graph
  bool sharded = false;
  macr_shard_desc sd{};
  macr::PeerGhosts ghosts{};
  macr::PeerFlagsDev peerF{};
  unsigned long long *flags = nullptr;  // [0..15] arrival flags, [16] barrier epoch, [17] error (int)
};

namespace macr {

static int mf_enqueue(macr_mf_trainer *h, int B) {
  const macr_hparams &hp = h->hp;
  cudaStream_t s = h->cs, side = h->side, side2 = h->side2;
  float *yp = h->sc(0, B), *yn = h->sc(1, B), *sp = h->sc(2, B), *sn = h->sc(3, B),
        *su = h->sc(4, B), *rq = h->sc(5, B), *dyp = h->sc(6, B), *dyn = h->sc(7, B),
        *dsp = h->sc(8, B), *dsn = h->sc(9, B), *dsu = h->sc(10, B);
  GridWs g = grid_ws_layout(B, h->gridws);
  g.item_gate_only = h->mode == MACR_TRAIN_RUBIBCE;
  int rc;
  if (h->sharded) {
    // renumber this step's ids and store every owned row into the peers' ghost slots; the plan and
    // the sweep need the renumbered ids only, so the flag barrier (peers' rows have landed) sits on
    // the main branch alone and the HBM-bound sweep never waits for a peer
    rc = launch_shard_push_st(h->U, h->I, h->sd, h->st, B, h->ghosts, s);
    if (rc) return rc;
  }
  // fork: the plan (2 CTAs) and the dense sweep need only the ids and the step state
  MACR_CUDA(cudaEventRecord(h->ev_fork, s));
  MACR_CUDA(cudaStreamWaitEvent(side, h->ev_fork, 0));
  MACR_CUDA(cudaStreamWaitEvent(side2, h->ev_fork, 0));
  rc = launch_batch_plan2(nullptr, h->st, 0, B, h->nu, h->planU, nullptr, nullptr, B, 2 * B,
                          h->ni, h->planI, nullptr, side);
  if (rc) return rc;
  MACR_CUDA(cudaEventRecord(h->ev_join, side));
  rc = launch_mark_touched(h->st, nullptr, B, h->bmU, h->bmI, side2);
  if (rc) return rc;
  // row-partitioned: the owned rows only -- peers store into the ghost rows while this runs
  const int64_t sweep_u = h->sharded ? h->sd.u_hi - h->sd.u_lo : h->nu;
  const int64_t sweep_i = h->sharded ? h->sd.i_hi - h->sd.i_lo : h->ni;
  rc = launch_adam_sweep2(h->U, h->mU, h->vU, sweep_u, h->bmU, h->I, h->mI, h->vI, sweep_i, h->bmI,
                          hp.lr, h->st, hp.beta1, hp.beta2, hp.eps, side2);
  if (rc) return rc;
  MACR_CUDA(cudaEventRecord(h->ev_join2, side2));
  // main: gather (+ row snapshot) -> B x B grid (+ band folds) -> row gradients + Adam + tail
  if (h->sharded) {
    rc = launch_peer_barrier_dev(h->peerF, h->flags + 16, reinterpret_cast<int *>(h->flags + 17),
                                 h->sd.rank, h->sd.world, s);
    if (rc) return rc;
  }
  rc = launch_gather_dots(h->U, h->I, h->U, h->I, h->w, h->wu, nullptr, nullptr, nullptr, h->st, B,
                          yp, yn, sp, sn, su, rq, h->snap, &g, s);
  if (rc) return rc;
  if (h->mode == MACR_TRAIN_NORMALBCE)
    rc = launch_plain_bce(yp, yn, B, hp, rq, h->st, nullptr, dyp, dyn, dsp, dsn, dsu, s);
  else
    rc = launch_grid_bce(yp, yn, B, hp, g, dyp, dyn, dsp, dsn, dsu, 1, rq, h->st, nullptr, s, true);
  if (rc) return rc;
  MACR_CUDA(cudaStreamWaitEvent(s, h->ev_join, 0));
  MACR_CUDA(cudaStreamWaitEvent(s, h->ev_join2, 0));
  const float lam = hp.decay / (float)hp.batch_size_flag;
  const AdamTabs tabs{h->U, h->mU, h->vU, h->I, h->mI, h->vI, h->bmU, h->bmI,
                      hp.beta1, hp.beta2, hp.eps, hp.lr, h->st};
  // vectors outside the mode's graph get no gradient: TF's minimize() leaves them and their slots
  const int frozen = h->mode == MACR_TRAIN_NORMALBCE ? 3 : h->mode == MACR_TRAIN_RUBIBCE ? 2 : 0;
  const TailArgs tail{1, h->w, h->mw, h->vw, h->wu, h->mwu, h->vwu, hp, h->st, h->tail_ticket,
                      frozen};
  rc = launch_row_grads(h->snap, h->w, h->wu, B, dyp, dyn, dsp, dsn, dsu, lam, h->planU, h->planI,
                        h->gU, h->gI, h->unit_part, h->gw_part, h->gwu_part, nullptr, &tabs, &tail, s, true);
  if (rc) return rc;
  h->launches = h->sharded ? 8 : 6;  // [push, barrier |] plan, mark, sweep | gather, grid, row-grads(+Adam+tail)
  return MACR_OK;
}

// `unroll` consecutive steps in one graph: nothing in a step depends on host values (the batch
// pointer, the Adam powers and the loss slot live in the device-resident StepState), so a run of
// steps is the same node sequence repeated -- and a kernel -> kernel edge inside a graph is
// cheaper than the boundary between two graph launches.
template <class H, class F>
static int get_graph(H *h, std::map<int, cudaGraphExec_t> &cache, int B, F enqueue,
                     cudaGraphExec_t *out, int unroll = 1) {
  const int key = B | (unroll << 16);
  auto it = cache.find(key);
  if (it != cache.end()) {
    *out = it->second;
    return MACR_OK;
  }
  cudaGraph_t graph = nullptr;
  MACR_CUDA(cudaStreamBeginCapture(h->cs, cudaStreamCaptureModeThreadLocal));
  int rc = MACR_OK;
  for (int k = 0; k < unroll && !rc; ++k) rc = enqueue(h, B);
  cudaError_t e = cudaStreamEndCapture(h->cs, &graph);
  if (rc) {
    if (graph) cudaGraphDestroy(graph);
    return rc;
  }
  if (e != cudaSuccess) return fail(MACR_ERR_CUDA, "graph capture: %s", cudaGetErrorString(e));
  cudaGraphExec_t exec = nullptr;
  e = cudaGraphInstantiate(&exec, graph, 0);
  cudaGraphDestroy(graph);
  if (e != cudaSuccess) return fail(MACR_ERR_CUDA, "graph instantiate: %s", cudaGetErrorString(e));
  cache[key] = exec;
  *out = exec;
  return MACR_OK;
}

// steps per graph of the epoch calls: 8 (measured at the gowalla shape: 57.9 us per step with one
// graph launch per step, 55.5 / 54.5 / 54.0 us with 2 / 4 / 8 steps per graph); MACR_GRAPH_UNROLL overrides
static int graph_unroll() {
  static int u = -1;
  if (u < 0) {
    const char *e = getenv("MACR_GRAPH_UNROLL");
    u = e ? atoi(e) : 8;
    if (u < 1 || u > 64) u = 8;
  }
  return u;
}

}  // namespace macr

using namespace macr;

static int check_tables(const void *U, const void *mU, const void *vU, const void *I,
                        const void *mI, const void *vI, const void *w, const void *mw,
                        const void *vw, const void *wu, const void *mwu, const void *vwu) {
  return (U && mU && vU && I && mI && vI && w && mw && vw && wu && mwu && vwu) ? 1 : 0;
}

extern "C" int macr_mf_trainer_create(macr_mf_trainer **out, float *U, float *mU, float *vU,
                                      int64_t n_users, float *I, float *mI, float *vI,
                                      int64_t n_items, float *w, float *mw, float *vw,
                                      float *w_user, float *mwu, float *vwu, int d, int max_batch,
                                      const macr_hparams *hp, macr_stream_t stream) {
  MACR_CHECK_ARG(out && hp, "macr_mf_trainer_create: null out/hparams");
  MACR_CHECK_ARG(d == kD, "macr_mf_trainer_create: d must be %d (got %d)", kD, d);
  MACR_CHECK_ARG(max_batch > 0 && max_batch <= 8192,
                 "macr_mf_trainer_create: max_batch must be in [1,8192] (got %d)", max_batch);
  MACR_CHECK_ARG(n_users > 0 && n_items > 0, "macr_mf_trainer_create: empty table");
  MACR_CHECK_ARG(check_tables(U, mU, vU, I, mI, vI, w, mw, vw, w_user, mwu, vwu),
                 "macr_mf_trainer_create: null table pointer");
  MACR_CHECK_ARG(hp->batch_size_flag > 0, "macr_mf_trainer_create: batch_size_flag must be > 0");
  macr_mf_trainer *h = new (std::nothrow) macr_mf_trainer();
  MACR_CHECK_ARG(h, "macr_mf_trainer_create: out of host memory");
  h->U = U; h->mU = mU; h->vU = vU; h->I = I; h->mI = mI; h->vI = vI;
  h->w = w; h->mw = mw; h->vw = vw; h->wu = w_user; h->mwu = mwu; h->vwu = vwu;
  h->nu = n_users; h->ni = n_items; h->maxB = max_batch; h->hp = *hp; h->launches = 0;
  int rc = h->alloc_common(as_stream(stream));
  if (rc) return rc;
  const size_t wu_words = (size_t)((n_users + 31) / 32), wi_words = (size_t)((n_items + 31) / 32);
  MACR_CUDA(cudaMalloc(&h->bmU, sizeof(uint32_t) * wu_words));
  MACR_CUDA(cudaMalloc(&h->bmI, sizeof(uint32_t) * wi_words));
  MACR_CUDA(cudaMemsetAsync(h->bmU, 0, sizeof(uint32_t) * wu_words, h->s));
  MACR_CUDA(cudaMemsetAsync(h->bmI, 0, sizeof(uint32_t) * wi_words, h->s));
  MACR_CUDA(cudaStreamSynchronize(h->s));
  *out = h;
  return MACR_OK;
}

template <class H, class F>
static int run_steps(H *h, std::map<int, cudaGraphExec_t> &cache, F enq, const int32_t *batches,
                     int n_steps, int B, float *losses) {
  MACR_CHECK_ARG(h, "trainer: null handle");
  MACR_CHECK_ARG(B > 0 && B <= h->maxB, "trainer: batch %d outside (0,%d]", B, h->maxB);
  MACR_CHECK_ARG(n_steps >= 0, "trainer: negative step count");
  if (n_steps == 0) return MACR_OK;
  cudaGraphExec_t exec, exec_u = nullptr;
  int rc = get_graph(h, cache, B, enq, &exec);
  if (rc) return rc;
  const int U = graph_unroll();
  if (U > 1 && n_steps >= U) {
    rc = get_graph(h, cache, B, enq, &exec_u, U);
    if (rc) return rc;
  }
  rc = h->set_io(batches, losses, n_steps, B);
  if (rc) return rc;
  int k = 0;
  if (exec_u)
    for (; k + U <= n_steps; k += U) MACR_CUDA(cudaGraphLaunch(exec_u, h->s));
  for (; k < n_steps; ++k) MACR_CUDA(cudaGraphLaunch(exec, h->s));
  return MACR_OK;
}

static int stage_ids_device(TrainerBase *h, const int32_t *u, const int32_t *p, const int32_t *n,
                            int B, cudaMemcpyKind kind) {
  const size_t nb = sizeof(int32_t) * (size_t)B;
  if (p == u + B && n == u + 2 * B) {
    MACR_CUDA(cudaMemcpyAsync(h->ids_stage, u, 3 * nb, kind, h->s));
  } else {
    MACR_CUDA(cudaMemcpyAsync(h->ids_stage, u, nb, kind, h->s));
    MACR_CUDA(cudaMemcpyAsync(h->ids_stage + B, p, nb, kind, h->s));
    MACR_CUDA(cudaMemcpyAsync(h->ids_stage + 2 * B, n, nb, kind, h->s));
  }
  return MACR_OK;
}

extern "C" int macr_mf_trainer_step(macr_mf_trainer *h, const int32_t *users, const int32_t *pos,
                                    const int32_t *neg, int B, float *losses_out) {
  MACR_CHECK_ARG(h && users && pos && neg, "macr_mf_trainer_step: null argument");
  MACR_CHECK_ARG(B > 0 && B <= h->maxB, "macr_mf_trainer_step: batch %d outside (0,%d]", B, h->maxB);
  int rc = stage_ids_device(h, users, pos, neg, B, cudaMemcpyDeviceToDevice);
  if (rc) return rc;
  rc = run_steps(h, h->graphs, mf_enqueue, h->ids_stage, 1, B, h->loss_stage);
  if (rc) return rc;
  h->steps_done += 1;
  if (losses_out)
    MACR_CUDA(cudaMemcpyAsync(losses_out, h->loss_stage, sizeof(float) * 4,
                              cudaMemcpyDeviceToDevice, h->s));
  return MACR_OK;
}

extern "C" int macr_mf_trainer_step_host(macr_mf_trainer *h, const int32_t *users_host,
                                         const int32_t *pos_host, const int32_t *neg_host, int B,
                                         float *losses_host) {
  MACR_CHECK_ARG(h && users_host && pos_host && neg_host, "macr_mf_trainer_step_host: null argument");
  MACR_CHECK_ARG(B > 0 && B <= h->maxB, "macr_mf_trainer_step_host: batch %d outside (0,%d]", B,
                 h->maxB);
  int rc = stage_ids_device(h, users_host, pos_host, neg_host, B, cudaMemcpyHostToDevice);
  if (rc) return rc;
  rc = run_steps(h, h->graphs, mf_enqueue, h->ids_stage, 1, B, h->loss_stage);
  if (rc) return rc;
  h->steps_done += 1;
  MACR_CUDA(cudaMemcpyAsync(h->pinned_losses, h->loss_stage, sizeof(float) * 4,
                            cudaMemcpyDeviceToHost, h->s));
  MACR_CUDA(cudaStreamSynchronize(h->s));
  if (losses_host) memcpy(losses_host, h->pinned_losses, sizeof(float) * 3);
  return MACR_OK;
}

// host batches [n_steps][3][B] -> device (one copy), n_steps graph replays, losses [n_steps][4]
// back to the host (one copy), one synchronisation
template <class H, class F>
static int run_steps_host(H *h, std::map<int, cudaGraphExec_t> &cache, F enq,
                          const int32_t *batches_host, int n_steps, int B, float *losses_host) {
  MACR_CHECK_ARG(B > 0 && B <= h->maxB, "trainer: batch %d outside (0,%d]", B, h->maxB);
  MACR_CHECK_ARG(n_steps >= 0, "trainer: negative step count");
  if (n_steps == 0) return MACR_OK;
  const size_t n_ids = (size_t)n_steps * 3 * (size_t)B, n_loss = (size_t)n_steps * 4;
  if (n_ids > h->epoch_ids_cap) {
    cudaFree(h->epoch_ids);
    h->epoch_ids = nullptr;
    h->epoch_ids_cap = 0;
    MACR_CUDA(cudaMalloc(&h->epoch_ids, sizeof(int32_t) * n_ids));
    h->epoch_ids_cap = n_ids;
  }
  if (n_loss > h->epoch_loss_cap) {
    cudaFree(h->epoch_losses);
    cudaFreeHost(h->epoch_losses_pinned);
    h->epoch_losses = h->epoch_losses_pinned = nullptr;
    h->epoch_loss_cap = 0;
    MACR_CUDA(cudaMalloc(&h->epoch_losses, sizeof(float) * n_loss));
    MACR_CUDA(cudaMallocHost(&h->epoch_losses_pinned, sizeof(float) * n_loss));
    h->epoch_loss_cap = n_loss;
  }
  MACR_CUDA(cudaMemcpyAsync(h->epoch_ids, batches_host, sizeof(int32_t) * n_ids,
                            cudaMemcpyHostToDevice, h->s));
  int rc = run_steps(h, cache, enq, h->epoch_ids, n_steps, B, h->epoch_losses);
  if (rc) return rc;
  MACR_CUDA(cudaMemcpyAsync(h->epoch_losses_pinned, h->epoch_losses, sizeof(float) * n_loss,
                            cudaMemcpyDeviceToHost, h->s));
  MACR_CUDA(cudaStreamSynchronize(h->s));
  if (losses_host) memcpy(losses_host, h->epoch_losses_pinned, sizeof(float) * n_loss);
  return MACR_OK;
}

extern "C" int macr_mf_trainer_run_host(macr_mf_trainer *h, const int32_t *batches_host,
                                        int n_steps, int B, float *losses_host) {
  MACR_CHECK_ARG(h && batches_host, "macr_mf_trainer_run_host: null argument");
  int rc = run_steps_host(h, h->graphs, mf_enqueue, batches_host, n_steps, B, losses_host);
  if (rc) return rc;
  h->steps_done += n_steps;
  return MACR_OK;
}

extern "C" int macr_mf_trainer_run(macr_mf_trainer *h, const int32_t *batches, int n_steps, int B,
                                   float *losses) {
  MACR_CHECK_ARG(h && batches && losses, "macr_mf_trainer_run: null argument");
  int rc = run_steps(h, h->graphs, mf_enqueue, batches, n_steps, B, losses);
  if (rc) return rc;
  h->steps_done += n_steps;
  return MACR_OK;
}

extern "C" int macr_mf_trainer_set_mode(macr_mf_trainer *h, int mode) {
  MACR_CHECK_ARG(h, "macr_mf_trainer_set_mode: null handle");
  MACR_CHECK_ARG(mode == MACR_TRAIN_RUBIBCEBOTH || mode == MACR_TRAIN_NORMALBCE ||
                     mode == MACR_TRAIN_RUBIBCE,
                 "macr_mf_trainer_set_mode: unknown mode %d", mode);
  if (mode != h->mode) {  // the captured step graphs belong to the old mode
    cudaStreamSynchronize(h->s);
    for (auto &kv : h->graphs) cudaGraphExecDestroy(kv.second);
    h->graphs.clear();
    h->mode = mode;
  }
  return MACR_OK;
}

extern "C" int macr_mf_trainer_launches_per_step(const macr_mf_trainer *h) {
  return h ? h->launches : 0;
}
extern "C" int64_t macr_mf_trainer_steps_done(const macr_mf_trainer *h) {
  return h ? h->steps_done : -1;
}
extern "C" int macr_mf_trainer_set_steps_done(macr_mf_trainer *h, int64_t t) {
  MACR_CHECK_ARG(h && t >= 0, "macr_mf_trainer_set_steps_done: bad argument");
  return h->set_steps(t);
}
// ---- row-partitioned mode -------------------------------------------------------------------
extern "C" int macr_mf_trainer_ipc_export(macr_mf_trainer *h,
                                          unsigned char handle[MACR_IPC_HANDLE_BYTES]) {
  MACR_CHECK_ARG(h && handle, "macr_mf_trainer_ipc_export: null argument");
  if (!h->flags) {
    MACR_CUDA(cudaMalloc(&h->flags, sizeof(unsigned long long) * 32));
    MACR_CUDA(cudaMemset(h->flags, 0, sizeof(unsigned long long) * 32));
  }
  cudaIpcMemHandle_t ih;
  MACR_CUDA(cudaIpcGetMemHandle(&ih, h->flags));
  memcpy(handle, &ih, sizeof(ih));
  return MACR_OK;
}

extern "C" int macr_mf_trainer_shard(macr_mf_trainer *h, const macr_shard_desc *desc,
                                     float *const *peer_U_ghost, float *const *peer_I_ghost,
                                     uint64_t *const *peer_flags) {
  MACR_CHECK_ARG(h && desc, "macr_mf_trainer_shard: null argument");
  MACR_CHECK_ARG(desc->world >= 1 && desc->world <= kMaxRanks && desc->rank >= 0 &&
                     desc->rank < desc->world,
                 "macr_mf_trainer_shard: rank %d / world %d", desc->rank, desc->world);
  MACR_CHECK_ARG(desc->max_batch == h->maxB, "macr_mf_trainer_shard: max_batch %d != trainer's %d",
                 desc->max_batch, h->maxB);
  MACR_CHECK_ARG(h->nu == (desc->u_hi - desc->u_lo) + 2LL * h->maxB &&
                     h->ni == (desc->i_hi - desc->i_lo) + 4LL * h->maxB,
                 "macr_mf_trainer_shard: the local tables must hold the owned rows + 2 x max_batch "
                 "(users) / 4 x max_batch (items) ghost rows");
  MACR_CHECK_ARG(h->steps_done == 0 && h->graphs.empty(), "macr_mf_trainer_shard: call it before the first step");
  MACR_CHECK_ARG(h->flags, "macr_mf_trainer_shard: call macr_mf_trainer_ipc_export first");
  for (int r = 0; r < desc->world; ++r) {
    if (r == desc->rank) continue;
    MACR_CHECK_ARG(peer_U_ghost && peer_I_ghost && peer_flags && peer_U_ghost[r] && peer_I_ghost[r] &&
                       peer_flags[r],
                   "macr_mf_trainer_shard: null peer pointer (rank %d)", r);
    h->ghosts.u[r] = peer_U_ghost[r];
    h->ghosts.i[r] = peer_I_ghost[r];
    h->peerF.p[r] = reinterpret_cast<unsigned long long *>(peer_flags[r]);
  }
  h->peerF.p[desc->rank] = h->flags;
  h->sd = *desc;
  h->sharded = true;
  h->renumber_ids = true;
  return MACR_OK;
}

extern "C" int macr_mf_trainer_peer_error(macr_mf_trainer *h, int *err_out) {
  MACR_CHECK_ARG(h && err_out, "macr_mf_trainer_peer_error: null argument");
  *err_out = 0;
  if (!h->flags) return MACR_OK;
  MACR_CUDA(cudaStreamSynchronize(h->s));
  MACR_CUDA(cudaMemcpy(err_out, h->flags + 17, sizeof(int), cudaMemcpyDeviceToHost));
  return MACR_OK;
}

extern "C" int macr_mf_trainer_destroy(macr_mf_trainer *h) {
  if (!h) return MACR_OK;
  cudaStreamSynchronize(h->s);
  cudaStreamSynchronize(h->side);
  h->free_common();
  cudaFree(h->flags);
  cudaFree(h->bmU);
  cudaFree(h->bmI);
  delete h;
  return MACR_OK;
}

// -------------------------------------------------------------------------------------------
// LightGCN
// -------------------------------------------------------------------------------------------
struct macr_lgcn_trainer : macr::TrainerBase {
  const int32_t *rowptr, *col;
  const float *val;
  int L;
  float *Emean, *tmp, *g3;  // [N][64], 2x[N][64], [N][64]
  uint32_t *g3_nz;          // bitmap: rows of g3 that hold a gradient this step
  uint32_t *need_bm = nullptr;  // bitmap: node rows the batch reads (last forward layer of a training step)
  macr::SpmmPlan plan;      // static segment decomposition of the adjacency
  bool emb_dirty;
  int launches;
  int mode = MACR_TRAIN_RUBIBCEBOTH;  // MACR_TRAIN_NORMALBCE: `--loss bce`
  // row-partitioned mode (macr_lgcn_trainer_shard): this rank owns the rows
  // [u_lo,u_hi) of the user table and [i_lo,i_hi) of the item table; every [N][64] buffer stays
  // full-size, its owned rows are computed here and stored into the peers' buffers (NVLink peer
  // stores + flag barrier) wherever the next kernel reads rows of other ranks
  bool sharded = false;
  macr_shard_desc sd{};
  float *peerU[macr::kMaxRanks], *peerI[macr::kMaxRanks], *peerE[macr::kMaxRanks], *peerT[macr::kMaxRanks];
  macr::PeerFlagsDev peerF{};
  unsigned long long *flags = nullptr;  // [0..15] arrival flags, [16] barrier epoch, [17] error (int)
};

namespace macr {

// all-gather of this rank's owned rows of one [N][64] buffer into every peer's copy
enum { XCH_TABLES = 0, XCH_EMEAN = 1, XCH_TMP0 = 2, XCH_TMP1 = 3 };
static int lgcn_exchange(macr_lgcn_trainer *h, int which, cudaStream_t s) {
  if (!h->sharded || h->sd.world <= 1) return MACR_OK;
  const macr_shard_desc &d = h->sd;
  const int64_t N = h->nu + h->ni;
  PeerPush p{};
  p.rank = d.rank;
  p.world = d.world;
  p.rows[0] = d.u_hi - d.u_lo;
  p.rows[1] = d.i_hi - d.i_lo;
  const long long uo = d.u_lo * kD, io_tab = d.i_lo * kD, io_buf = (h->nu + d.i_lo) * kD;
  if (which == XCH_TABLES) {
    p.src[0] = h->U + uo;
    p.src[1] = h->I + io_tab;
  } else {
    const float *base = which == XCH_EMEAN ? h->Emean : h->tmp + (which - XCH_TMP0) * N * kD;
    p.src[0] = base + uo;
    p.src[1] = base + io_buf;
  }
  for (int r = 0; r < d.world; ++r) {
    if (r == d.rank) continue;
    if (which == XCH_TABLES) {
      p.dst[0][r] = h->peerU[r] + uo;
      p.dst[1][r] = h->peerI[r] + io_tab;
    } else {
      float *base = which == XCH_EMEAN ? h->peerE[r] : h->peerT[r] + (which - XCH_TMP0) * N * kD;
      p.dst[0][r] = base + uo;
      p.dst[1][r] = base + io_buf;
    }
  }
  int rc = launch_peer_push(p, s);
  if (rc) return rc;
  return launch_peer_barrier_dev(h->peerF, h->flags + 16, reinterpret_cast<int *>(h->flags + 17), d.rank,
                                 d.world, s);
}

// E_mean of the current tables (LightGCN.py:288-309).  Row-partitioned: E0's rows of other ranks are
// fetched first (their owners have just updated them), every layer is computed for the owned rows
// and all-gathered, and so is the layer mean.  -> number of kernels launched
// batch_rows (nullable): the training step reads E_mean at the batch's rows only, so the LAST layer
// is computed for those rows alone (E_mean's other rows are left stale: emb_dirty stays set).
static int lgcn_forward(macr_lgcn_trainer *h, cudaStream_t s, int *launches,
                        const uint32_t *batch_rows = nullptr, int B = 0) {
  const int per_spmm = h->plan.n_multi ? 2 : 1;
  if (!h->sharded && (batch_rows == nullptr || h->L == 0)) {
    if (launches) *launches += h->L * per_spmm;
    return launch_lgcn_propagate(h->rowptr, h->col, h->val, h->U, h->nu, h->I, h->ni, h->L, h->Emean,
                                 h->tmp, s, &h->plan);
  }
  const int64_t N = h->nu + h->ni;
  int rc = h->sharded ? lgcn_exchange(h, XCH_TABLES, s) : MACR_OK;
  if (rc) return rc;
  RowSrc e0{h->U, h->I, h->nu};
  if (h->L == 0) {
    if (launches) *launches += 2;
    return launch_lgcn_propagate(h->rowptr, h->col, h->val, h->U, h->nu, h->I, h->ni, 0, h->Emean, h->tmp, s,
                                 &h->plan);
  }
  float *buf[2] = {h->tmp, h->tmp + N * kD};
  RowSrc x = e0;
  for (int k = 0; k < h->L; ++k) {
    const bool last = k == h->L - 1;
    float *y = last ? nullptr : buf[k & 1];
    RowSrc accin = (k == 0) ? e0 : RowSrc{h->Emean, h->Emean, N};
    rc = launch_spmm_planned(&h->plan, h->rowptr, h->col, h->val, N, x, nullptr, y, accin, h->Emean,
                             last ? (float)(h->L + 1) : 0.f, nullptr, s, last ? batch_rows : nullptr);
    if (rc) return rc;
    if (!last) {
      rc = h->sharded ? lgcn_exchange(h, XCH_TMP0 + (k & 1), s) : MACR_OK;
      if (rc) return rc;
      x = RowSrc{y, y, N};
    }
  }
  if (launches) *launches += h->L * per_spmm + (h->sharded ? 2 * (h->L + 1) : 0);
  if (!h->sharded) return MACR_OK;
  if (batch_rows == nullptr) return lgcn_exchange(h, XCH_EMEAN, s);
  // training step: only the batch's rows of the layer mean were computed and only they are read
  // on the other ranks -- 3B rows (<= 6 MB) instead of the owned range of an [N][64] buffer
  PeerBufs pe{};
  for (int r = 0; r < h->sd.world; ++r) pe.p[r] = r == h->sd.rank ? nullptr : h->peerE[r];
  rc = launch_peer_push_batch_rows(h->Emean, pe, h->sd, h->nu, h->st, B, s);
  if (rc) return rc;
  return launch_peer_barrier_dev(h->peerF, h->flags + 16, reinterpret_cast<int *>(h->flags + 17),
                                 h->sd.rank, h->sd.world, s);
}

static int lgcn_enqueue_impl(macr_lgcn_trainer *h, int B, int train) {
  const macr_hparams &hp = h->hp;
  cudaStream_t s = h->cs, side = h->side;
  const int64_t N = h->nu + h->ni;
  float *yp = h->sc(0, B), *yn = h->sc(1, B), *sp = h->sc(2, B), *sn = h->sc(3, B),
        *su = h->sc(4, B), *rq = h->sc(5, B), *dyp = h->sc(6, B), *dyn = h->sc(7, B),
        *dsp = h->sc(8, B), *dsn = h->sc(9, B), *dsu = h->sc(10, B);
  GridWs g = grid_ws_layout(B, h->gridws);
  g.item_gate_only = h->mode == MACR_TRAIN_RUBIBCE;  // `--loss bce1`, LightGCN.py:431-461
  const float *Ue = h->Emean, *Ie = h->Emean + h->nu * kD;
  int rc, launches = 0;
  if (train) {
    MACR_CUDA(cudaEventRecord(h->ev_fork, s));
    MACR_CUDA(cudaStreamWaitEvent(side, h->ev_fork, 0));
    rc = launch_batch_plan2(nullptr, h->st, 0, B, h->nu, h->planU, nullptr, nullptr, B, 2 * B,
                            h->ni, h->planI, nullptr, side);
    if (rc) return rc;
    MACR_CUDA(cudaMemsetAsync(h->g3, 0, sizeof(float) * N * kD, side));
    MACR_CUDA(cudaMemsetAsync(h->g3_nz, 0, sizeof(uint32_t) * ((N + 31) / 32), side));
    // second side branch: the bitmap of the batch's node rows (read by the last forward layer)
    cudaStream_t side2 = h->side2;
    MACR_CUDA(cudaStreamWaitEvent(side2, h->ev_fork, 0));
    MACR_CUDA(cudaMemsetAsync(h->need_bm, 0, sizeof(uint32_t) * ((N + 31) / 32), side2));
    rc = launch_mark_batch_rows(h->st, B, h->nu, h->need_bm, side2);
    if (rc) return rc;
    MACR_CUDA(cudaEventRecord(h->ev_join2, side2));
    MACR_CUDA(cudaEventRecord(h->ev_join, side));
    launches += 2;
    MACR_CUDA(cudaStreamWaitEvent(s, h->ev_join2, 0));  // a few microseconds of work, long done
  }
  rc = lgcn_forward(h, s, &launches, train ? h->need_bm : nullptr, B);
  if (rc) return rc;
  rc = launch_gather_dots(Ue, Ie, h->U, h->I, h->w, h->wu, nullptr, nullptr, nullptr, h->st, B, yp,
                          yn, sp, sn, su, rq, h->snap, &g, s);
  if (rc) return rc;
  if (h->mode == MACR_TRAIN_NORMALBCE)  // LightGCN.py:415-429: element-wise BCE, loss = mf + emb
    rc = launch_plain_bce(yp, yn, B, hp, rq, h->st, nullptr, dyp, dyn, dsp, dsn, dsu, s);
  else
    rc = launch_grid_bce(yp, yn, B, hp, g, dyp, dyn, dsp, dsn, dsu, train, rq, h->st, nullptr, s);
  if (rc) return rc;
  launches += 2;
  if (train) {
    MACR_CUDA(cudaStreamWaitEvent(s, h->ev_join, 0));
    int n_part = 0;
    rc = launch_row_grads(h->snap, h->w, h->wu, B, dyp, dyn, dsp, dsn, dsu, 0.f, h->planU,
                          h->planI, h->gU, h->gI, h->unit_part, h->gw_part, h->gwu_part, &n_part,
                          nullptr, nullptr, s);
    if (rc) return rc;
    rc = launch_scatter_rows(h->planU, h->gU, h->planI, h->gI, B, h->nu, (float)(h->L + 1), h->g3,
                             h->g3_nz, s);
    if (rc) return rc;
    launches += 2;
    // backward through the layer stack: acc_L = g3; acc_{k-1} = g3 + A^T acc_k  (A symmetric)
    float *buf[2] = {h->tmp, h->tmp + N * kD};
    const float *acc = h->g3;
    for (int k = 0; k < h->L; ++k) {
      RowSrc x{acc, acc, N};
      // acc_L = g3 is row-sparse (<= 3B rows): nonzeros pointing at all-zero rows are skipped
      rc = launch_spmm_planned(&h->plan, h->rowptr, h->col, h->val, N, x, h->g3, buf[k & 1], x,
                               nullptr, 0.f, k == 0 ? h->g3_nz : nullptr, s);
      if (rc) return rc;
      acc = buf[k & 1];
      launches += h->plan.n_multi ? 2 : 1;
      if (h->sharded && k + 1 < h->L) {  // the next layer reads acc rows of every rank
        rc = lgcn_exchange(h, XCH_TMP0 + (k & 1), s);
        if (rc) return rc;
        launches += 2;
      }
    }
    float *grad = const_cast<float *>(acc);
    const float lam = hp.decay / (float)hp.batch_size_flag;
    rc = launch_l2_rows(h->planU, h->planI, B, h->U, h->I, h->nu, lam, grad, s);
    if (rc) return rc;
    // dense Adam: every row of both tables, or -- row-partitioned -- the rows this rank owns
    const int64_t u0 = h->sharded ? h->sd.u_lo : 0, u1 = h->sharded ? h->sd.u_hi : h->nu;
    const int64_t i0 = h->sharded ? h->sd.i_lo : 0, i1 = h->sharded ? h->sd.i_hi : h->ni;
    rc = launch_adam_dense(h->U + u0 * kD, h->mU + u0 * kD, h->vU + u0 * kD, grad + u0 * kD,
                           (u1 - u0) * kD, hp.lr, h->st, hp.beta1, hp.beta2, hp.eps, s);
    if (rc) return rc;
    rc = launch_adam_dense(h->I + i0 * kD, h->mI + i0 * kD, h->vI + i0 * kD,
                           grad + (h->nu + i0) * kD, (i1 - i0) * kD, hp.lr, h->st, hp.beta1,
                           hp.beta2, hp.eps, s);
    if (rc) return rc;
    launches += 3;
    const int frozen = h->mode == MACR_TRAIN_NORMALBCE ? 3 : h->mode == MACR_TRAIN_RUBIBCE ? 2 : 0;
    rc = launch_step_tail(h->w, h->mw, h->vw, h->wu, h->mwu, h->vwu, h->gw_part, h->gwu_part,
                          n_part, hp, h->st, 1 | (frozen << 1), s);
    if (rc) return rc;
  } else {
    rc = launch_step_tail(h->w, h->mw, h->vw, h->wu, h->mwu, h->vwu, h->gw_part, h->gwu_part, 0,
                          hp, h->st, 0, s);
    if (rc) return rc;
  }
  launches += 1;
  if (train) h->launches = launches;
  return MACR_OK;
}
static int lgcn_enqueue_train(macr_lgcn_trainer *h, int B) { return lgcn_enqueue_impl(h, B, 1); }
static int lgcn_enqueue_eval(macr_lgcn_trainer *h, int B) { return lgcn_enqueue_impl(h, B, 0); }

}  // namespace macr

extern "C" int macr_lgcn_trainer_create(macr_lgcn_trainer **out, const int32_t *rowptr,
                                        const int32_t *col, const float *val, float *U, float *mU,
                                        float *vU, int64_t n_users, float *I, float *mI, float *vI,
                                        int64_t n_items, float *w, float *mw, float *vw,
                                        float *w_user, float *mwu, float *vwu, int d, int n_layers,
                                        int max_batch, const macr_hparams *hp,
                                        macr_stream_t stream) {
  MACR_CHECK_ARG(out && hp, "macr_lgcn_trainer_create: null out/hparams");
  MACR_CHECK_ARG(d == kD, "macr_lgcn_trainer_create: d must be %d (got %d)", kD, d);
  MACR_CHECK_ARG(max_batch > 0 && max_batch <= 8192,
                 "macr_lgcn_trainer_create: max_batch must be in [1,8192] (got %d)", max_batch);
  MACR_CHECK_ARG(n_layers >= 0 && n_layers <= 16, "macr_lgcn_trainer_create: bad n_layers %d",
                 n_layers);
  MACR_CHECK_ARG(rowptr && col && val, "macr_lgcn_trainer_create: null adjacency");
  MACR_CHECK_ARG(n_users > 0 && n_items > 0, "macr_lgcn_trainer_create: empty table");
  MACR_CHECK_ARG(check_tables(U, mU, vU, I, mI, vI, w, mw, vw, w_user, mwu, vwu),
                 "macr_lgcn_trainer_create: null table pointer");
  MACR_CHECK_ARG(hp->batch_size_flag > 0, "macr_lgcn_trainer_create: batch_size_flag must be > 0");
  macr_lgcn_trainer *h = new (std::nothrow) macr_lgcn_trainer();
  MACR_CHECK_ARG(h, "macr_lgcn_trainer_create: out of host memory");
  h->U = U; h->mU = mU; h->vU = vU; h->I = I; h->mI = mI; h->vI = vI;
  h->w = w; h->mw = mw; h->vw = vw; h->wu = w_user; h->mwu = mwu; h->vwu = vwu;
  h->nu = n_users; h->ni = n_items; h->maxB = max_batch; h->hp = *hp;
  h->rowptr = rowptr; h->col = col; h->val = val; h->L = n_layers; h->launches = 0;
  int rc = h->alloc_common(as_stream(stream));
  if (rc) return rc;
  const size_t ne = (size_t)(n_users + n_items) * kD;
  MACR_CUDA(cudaMalloc(&h->Emean, sizeof(float) * ne));
  MACR_CUDA(cudaMalloc(&h->tmp, sizeof(float) * ne * 2));
  MACR_CUDA(cudaMalloc(&h->g3, sizeof(float) * ne));
  MACR_CUDA(cudaMalloc(&h->g3_nz, sizeof(uint32_t) * ((size_t)(n_users + n_items + 31) / 32 + 1)));
  MACR_CUDA(cudaMalloc(&h->need_bm, sizeof(uint32_t) * ((size_t)(n_users + n_items + 31) / 32 + 1)));
  rc = build_spmm_plan(rowptr, n_users + n_items, &h->plan);
  if (rc) return rc;
  MACR_CUDA(cudaMalloc(&h->flags, sizeof(unsigned long long) * 32));
  MACR_CUDA(cudaMemset(h->flags, 0, sizeof(unsigned long long) * 32));
  h->emb_dirty = true;
  MACR_CUDA(cudaStreamSynchronize(h->s));
  *out = h;
  return MACR_OK;
}

extern "C" int macr_lgcn_trainer_step(macr_lgcn_trainer *h, const int32_t *users,
                                      const int32_t *pos, const int32_t *neg, int B, int train,
                                      float *losses_out) {
  MACR_CHECK_ARG(h && users && pos && neg, "macr_lgcn_trainer_step: null argument");
  MACR_CHECK_ARG(B > 0 && B <= h->maxB, "macr_lgcn_trainer_step: batch %d outside (0,%d]", B,
                 h->maxB);
  int rc = stage_ids_device(h, users, pos, neg, B, cudaMemcpyDeviceToDevice);
  if (rc) return rc;
  rc = train ? run_steps(h, h->graphs, lgcn_enqueue_train, h->ids_stage, 1, B, h->loss_stage)
             : run_steps(h, h->graphs_eval, lgcn_enqueue_eval, h->ids_stage, 1, B, h->loss_stage);
  if (rc) return rc;
  if (train) {
    h->steps_done += 1;
    h->emb_dirty = true;
  } else {
    h->emb_dirty = false;  // Emean now holds the propagation of the current parameters
  }
  if (losses_out)
    MACR_CUDA(cudaMemcpyAsync(losses_out, h->loss_stage, sizeof(float) * 4,
                              cudaMemcpyDeviceToDevice, h->s));
  return MACR_OK;
}

extern "C" int macr_lgcn_trainer_step_host(macr_lgcn_trainer *h, const int32_t *users_host,
                                           const int32_t *pos_host, const int32_t *neg_host, int B,
                                           int train, float *losses_host) {
  MACR_CHECK_ARG(h && users_host && pos_host && neg_host, "macr_lgcn_trainer_step_host: null argument");
  MACR_CHECK_ARG(B > 0 && B <= h->maxB, "macr_lgcn_trainer_step_host: batch %d outside (0,%d]", B,
                 h->maxB);
  int rc = stage_ids_device(h, users_host, pos_host, neg_host, B, cudaMemcpyHostToDevice);
  if (rc) return rc;
  rc = train ? run_steps(h, h->graphs, lgcn_enqueue_train, h->ids_stage, 1, B, h->loss_stage)
             : run_steps(h, h->graphs_eval, lgcn_enqueue_eval, h->ids_stage, 1, B, h->loss_stage);
  if (rc) return rc;
  if (train) {
    h->steps_done += 1;
    h->emb_dirty = true;
  } else {
    h->emb_dirty = false;
  }
  MACR_CUDA(cudaMemcpyAsync(h->pinned_losses, h->loss_stage, sizeof(float) * 4,
                            cudaMemcpyDeviceToHost, h->s));
  MACR_CUDA(cudaStreamSynchronize(h->s));
  if (losses_host) memcpy(losses_host, h->pinned_losses, sizeof(float) * 3);
  return MACR_OK;
}

extern "C" int macr_lgcn_trainer_run(macr_lgcn_trainer *h, const int32_t *batches, int n_steps,
                                     int B, int train, float *losses) {
  MACR_CHECK_ARG(h && batches && losses, "macr_lgcn_trainer_run: null argument");
  int rc = train ? run_steps(h, h->graphs, lgcn_enqueue_train, batches, n_steps, B, losses)
                 : run_steps(h, h->graphs_eval, lgcn_enqueue_eval, batches, n_steps, B, losses);
  if (rc) return rc;
  if (train && n_steps > 0) {
    h->steps_done += n_steps;
    h->emb_dirty = true;
  }
  return MACR_OK;
}

extern "C" int macr_lgcn_trainer_run_host(macr_lgcn_trainer *h, const int32_t *batches_host,
                                          int n_steps, int B, int train, float *losses_host) {
  MACR_CHECK_ARG(h && batches_host, "macr_lgcn_trainer_run_host: null argument");
  int rc = train ? run_steps_host(h, h->graphs, lgcn_enqueue_train, batches_host, n_steps, B,
                                  losses_host)
                 : run_steps_host(h, h->graphs_eval, lgcn_enqueue_eval, batches_host, n_steps, B,
                                  losses_host);
  if (rc) return rc;
  if (train && n_steps > 0) {
    h->steps_done += n_steps;
    h->emb_dirty = true;
  }
  return MACR_OK;
}

extern "C" int macr_lgcn_trainer_set_mode(macr_lgcn_trainer *h, int mode) {
  MACR_CHECK_ARG(h, "macr_lgcn_trainer_set_mode: null handle");
  MACR_CHECK_ARG(mode == MACR_TRAIN_RUBIBCEBOTH || mode == MACR_TRAIN_NORMALBCE ||
                     mode == MACR_TRAIN_RUBIBCE,
                 "macr_lgcn_trainer_set_mode: unknown mode %d", mode);
  if (mode != h->mode) {
    cudaStreamSynchronize(h->s);
    for (auto &kv : h->graphs) cudaGraphExecDestroy(kv.second);
    for (auto &kv : h->graphs_eval) cudaGraphExecDestroy(kv.second);
    h->graphs.clear();
    h->graphs_eval.clear();
    h->mode = mode;
  }
  return MACR_OK;
}

extern "C" int macr_lgcn_trainer_embeddings(macr_lgcn_trainer *h, const float **Emean) {
  MACR_CHECK_ARG(h && Emean, "macr_lgcn_trainer_embeddings: null argument");
  if (h->emb_dirty) {
    int rc = lgcn_forward(h, h->s, nullptr);
    if (rc) return rc;
    h->emb_dirty = false;
  }
  *Emean = h->Emean;
  return MACR_OK;
}

extern "C" int macr_lgcn_trainer_launches_per_step(const macr_lgcn_trainer *h) {
  return h ? h->launches : 0;
}
extern "C" int64_t macr_lgcn_trainer_steps_done(const macr_lgcn_trainer *h) {
  return h ? h->steps_done : -1;
}
extern "C" int macr_lgcn_trainer_set_steps_done(macr_lgcn_trainer *h, int64_t t) {
  MACR_CHECK_ARG(h && t >= 0, "macr_lgcn_trainer_set_steps_done: bad argument");
  h->emb_dirty = true;  // checkpoint resume: the caller has just overwritten the tables
  return h->set_steps(t);
}
// ---- row-partitioned mode -------------------------------------------------------------------
extern "C" int macr_lgcn_trainer_ipc_export(macr_lgcn_trainer *h,
                                            unsigned char handles[3][MACR_IPC_HANDLE_BYTES]) {
  MACR_CHECK_ARG(h && handles, "macr_lgcn_trainer_ipc_export: null argument");
  void *bufs[3] = {h->Emean, h->tmp, h->flags};
  for (int k = 0; k < 3; ++k) {
    cudaIpcMemHandle_t ih;
    MACR_CUDA(cudaIpcGetMemHandle(&ih, bufs[k]));
    memcpy(handles[k], &ih, sizeof(ih));
  }
  return MACR_OK;
}

extern "C" int macr_lgcn_trainer_shard(macr_lgcn_trainer *h, const macr_shard_desc *desc,
                                       float *const *peer_U, float *const *peer_I,
                                       float *const *peer_Emean, float *const *peer_tmp,
                                       uint64_t *const *peer_flags) {
  MACR_CHECK_ARG(h && desc, "macr_lgcn_trainer_shard: null argument");
  MACR_CHECK_ARG(desc->world >= 1 && desc->world <= kMaxRanks && desc->rank >= 0 &&
                     desc->rank < desc->world,
                 "macr_lgcn_trainer_shard: rank %d / world %d", desc->rank, desc->world);
  MACR_CHECK_ARG(0 <= desc->u_lo && desc->u_lo <= desc->u_hi && desc->u_hi <= h->nu &&
                     0 <= desc->i_lo && desc->i_lo <= desc->i_hi && desc->i_hi <= h->ni,
                 "macr_lgcn_trainer_shard: owned ranges outside the tables");
  MACR_CHECK_ARG(h->steps_done == 0 && h->graphs.empty() && h->graphs_eval.empty(),
                 "macr_lgcn_trainer_shard: call it before the first step");
  for (int r = 0; r < desc->world; ++r) {
    if (r == desc->rank) continue;
    MACR_CHECK_ARG(peer_U && peer_I && peer_Emean && peer_tmp && peer_flags && peer_U[r] && peer_I[r] &&
                       peer_Emean[r] && peer_tmp[r] && peer_flags[r],
                   "macr_lgcn_trainer_shard: null peer pointer (rank %d)", r);
    h->peerU[r] = peer_U[r];
    h->peerI[r] = peer_I[r];
    h->peerE[r] = peer_Emean[r];
    h->peerT[r] = peer_tmp[r];
    h->peerF.p[r] = reinterpret_cast<unsigned long long *>(peer_flags[r]);
  }
  h->peerF.p[desc->rank] = h->flags;
  h->sd = *desc;
  // the segment plan of the rows this rank owns (a row's segmentation depends on the row alone,
  // so owned rows come out bit-identical to the single-GPU propagation)
  free_spmm_plan(&h->plan);
  const int64_t ranges[4] = {desc->u_lo, desc->u_hi, h->nu + desc->i_lo, h->nu + desc->i_hi};
  int rc = build_spmm_plan(h->rowptr, h->nu + h->ni, &h->plan, ranges);
  if (rc) return rc;
  h->sharded = true;
  h->emb_dirty = true;
  return MACR_OK;
}

extern "C" int macr_lgcn_trainer_peer_error(macr_lgcn_trainer *h, int *err_out) {
  MACR_CHECK_ARG(h && err_out, "macr_lgcn_trainer_peer_error: null argument");
  MACR_CUDA(cudaStreamSynchronize(h->s));
  MACR_CUDA(cudaMemcpy(err_out, h->flags + 17, sizeof(int), cudaMemcpyDeviceToHost));
  return MACR_OK;
}

extern "C" int macr_lgcn_trainer_destroy(macr_lgcn_trainer *h) {
  if (!h) return MACR_OK;
  cudaStreamSynchronize(h->s);
  cudaStreamSynchronize(h->side);
  h->free_common();
  cudaFree(h->Emean);
  cudaFree(h->tmp);
  cudaFree(h->g3);
  cudaFree(h->g3_nz);
  cudaFree(h->need_bm);
  cudaFree(h->flags);
  free_spmm_plan(&h->plan);
  delete h;
  return MACR_OK;
}
