/*
 * macr_b200.h -- C ABI of libmacr_b200.so, the B200 (sm_100a) implementation of
 * MACR's training / scoring hot path.
 *
 * Conventions (all entry points):
 *   - every array pointer is a DEVICE pointer owned by the caller unless the
 *     parameter name ends in _host; nothing is allocated behind the caller's
 *     back except inside the opaque trainer handles (created/destroyed explicitly);
 *   - `stream` is a cudaStream_t passed as void* (NULL = legacy default stream);
 *   - return value: 0 = ok, <0 = error; macr_last_error() returns a thread-local
 *     message for the last failing call on this host thread;
 *   - callable from any host thread; no Python / GIL interaction;
 *   - all floating point is fp32, all ids are int32 (reference placeholders are
 *     tf.int32 / tf.float32: macr_mf/model.py:27-29, macr_lightgcn/LightGCN.py:61-63);
 *   - embedding width d must be 64 (README commands all use --embed_size 64).
 *
 * Each declaration cites the reference interface (file:line under /root/reference)
 * that it replaces.  The reference has no plugin API: the path sits behind a
 * TF-1.14 session (`sess.run(fetches, feed_dict)`) and one Cython/C++ evaluator
 * FFI; INTEGRATION.md shows the binding a maintainer would add for each.
 */
#ifndef MACR_B200_H
#define MACR_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define MACR_OK 0
#define MACR_ERR_INVALID (-1) /* bad argument (shape, null pointer, unsupported d/K) */
#define MACR_ERR_CUDA (-2)    /* a CUDA runtime call or kernel launch failed        */
#define MACR_ERR_WORKSPACE (-3) /* caller-provided workspace too small              */

#define MACR_EMBED_DIM 64
#define MACR_MAX_TOPK 32 /* one rank per lane of a warp */

typedef void *macr_stream_t; /* cudaStream_t */

const char *macr_last_error(void);
int macr_abi_version(void);
/* number of SMs of the current device (grid sizing by callers / bench) */
int macr_device_sm_count(int *out_sms);

/* ------------------------------------------------------------------------- *
 * K1+K2  gather + five dot products (+ L2 sum of squares)
 * replaces: tf.nn.embedding_lookup x3 macr_mf/model.py:35-37, the reductions
 *   :186-187 and the three [B,64]x[64,1] matmuls :194-196; tf.nn.l2_loss x3 :219;
 *   LightGCN.py:145-150 (propagated rows for scores, raw rows for the L2 term),
 *   :496-497,:504-506,:525-526.
 *  Ue/Ie : tables the scores are taken from   (MF: the variables themselves;
 *          LightGCN: the propagated ua/ia embeddings)
 *  Ur/Ir : tables the L2 term is taken from   (MF: same pointers as Ue/Ie)
 *  out   : yp[b]=u.p  yn[b]=u.n  sp[b]=p.w  sn[b]=n.w  su[b]=u.w_user
 *          regsq[b]=|u|^2+|p|^2+|n|^2 (of the Ur/Ir rows)
 * ------------------------------------------------------------------------- */
int macr_gather_dots(const float *Ue, const float *Ie, const float *Ur, const float *Ir,
                     const float *w, const float *w_user,
                     const int32_t *users, const int32_t *pos, const int32_t *neg,
                     int B, int d,
                     float *yp, float *yn, float *sp, float *sn, float *su, float *regsq,
                     macr_stream_t stream);

/* ------------------------------------------------------------------------- *
 * K3  B x B sigmoid-gated BCE grid, forward + backward in one pass
 * replaces: the [B]*[B,1] broadcast macr_mf/model.py:204-205, the three means
 *   :211,:213,:215, their sum :217 and TF autodiff of all of it (model.py:74);
 *   identically LightGCN.py:513-514,:517-523.
 *   P[i,j]=yp[j]*sig(sp[i])*sig(su[i])   N[i,j]=yn[j]*sig(sn[i])*sig(su[i])
 *  losses3 = {L_ori, L_item, L_user} (means, before alpha/beta weighting)
 *  d_yp,d_yn,d_sp,d_sn,d_su = gradient of (L_ori + alpha*L_item + beta*L_user)
 *  ws: scratch, at least macr_grid_bce_workspace_bytes(B) bytes.
 * ------------------------------------------------------------------------- */
size_t macr_grid_bce_workspace_bytes(int B);
int macr_grid_bce_fwd_bwd(const float *yp, const float *yn, const float *sp, const float *sn,
                          const float *su, int B, float alpha, float beta,
                          float *losses3,
                          float *d_yp, float *d_yn, float *d_sp, float *d_sn, float *d_su,
                          void *ws, size_t ws_bytes, macr_stream_t stream);

/* ------------------------------------------------------------------------- *
 * K5a  batch plan: group the batch positions that hit the same table row
 * replaces: TF-1.14 optimizer.py _deduplicate_indexed_slices (array_ops.unique +
 *   unsorted_segment_sum) that AdamOptimizer runs on the IndexedSlices gradient of
 *   embedding_lookup (third-party, not vendored).
 *  ids   : n_ids row ids (users: B ids; items: pos followed by neg, 2B ids)
 *  out   : uniq_rows[n_uniq] ascending (array_ops.unique lists them in first-occurrence
 *          order; the order of the unique rows enters no result), seg_off[n_uniq+1],
 *          seg_pos[n_ids] (positions into `ids`, ascending inside a segment = the order
 *          unsorted_segment_sum adds in), *n_uniq (device int)
 *          touched_bitmap (nullable): one bit per table row, set for every row in `ids`
 *          (caller zero-initialises once; macr_adam_rows clears the bits)
 *  ws    : scratch, at least macr_batch_plan_workspace_bytes(n_ids) bytes.
 * ------------------------------------------------------------------------- */
size_t macr_batch_plan_workspace_bytes(int n_ids);
int macr_batch_plan(const int32_t *ids, int n_ids, int64_t table_rows,
                    int32_t *uniq_rows, int32_t *seg_off, int32_t *seg_pos, int32_t *n_uniq,
                    uint32_t *touched_bitmap, void *ws, size_t ws_bytes, macr_stream_t stream);

/* ------------------------------------------------------------------------- *
 * K5b  TF-1.14 Adam, dense semantics (adam.py _apply_sparse_shared):
 *   m <- b1*m (all rows); m[idx] += (1-b1)*g;  v <- b2*v (all rows);
 *   v[idx] += (1-b2)*g*g;  var <- var - lr_t*m/(sqrt(v)+eps) (all rows)
 * replaces: tf.train.AdamOptimizer(lr).minimize(...)  macr_mf/model.py:74,
 *   macr_lightgcn/LightGCN.py:201.
 *  macr_adam_sweep_untouched : rows whose bit in touched_bitmap is 0
 *     (bitmap may be NULL = all rows): pure decay + move, 24 B / element.
 *  macr_adam_rows            : rows listed in uniq_rows with their summed
 *     gradient rows grad_rows[n_uniq][d]; clears their bitmap bits.
 *  macr_adam_dense           : every row has a gradient (LightGCN tables:
 *     the dense dE0 is converted to IndexedSlices over range(rows), so the
 *     same formula applies) ; grad is [rows][d].
 *  macr_adam_vec             : training_ops ApplyAdam for w / w_user
 *     (m += (g-m)(1-b1); v += (g*g-v)(1-b2); var -= m*lr_t/(sqrt(v)+eps)).
 * ------------------------------------------------------------------------- */
int macr_adam_sweep_untouched(float *var, float *m, float *v, int64_t rows, int d,
                              const uint32_t *touched_bitmap,
                              float lr_t, float beta1, float beta2, float eps,
                              macr_stream_t stream);
int macr_adam_rows(float *var, float *m, float *v, int64_t rows, int d,
                   const int32_t *uniq_rows, const float *grad_rows, int n_uniq,
                   uint32_t *touched_bitmap,
                   float lr_t, float beta1, float beta2, float eps, macr_stream_t stream);
int macr_adam_dense(float *var, float *m, float *v, const float *grad, int64_t rows, int d,
                    float lr_t, float beta1, float beta2, float eps, macr_stream_t stream);
int macr_adam_vec(float *var, float *m, float *v, const float *grad, int n,
                  float lr_t, float beta1, float beta2, float eps, macr_stream_t stream);

/* ------------------------------------------------------------------------- *
 * Whole MF training step behind one handle (what `sess.run([opt_two_bce_both,
 * loss_two_bce_both, mf_loss_two_bce_both, reg_loss_two_bce_both], feed)` does,
 * macr_mf/train.py:492-496).  The handle owns scratch + a captured CUDA graph;
 * tables and Adam slots stay caller-owned device memory and are updated in place.
 * ------------------------------------------------------------------------- */
typedef struct macr_mf_trainer macr_mf_trainer;

typedef struct macr_hparams {
  float lr;        /* --lr            macr_mf/parse.py:47                          */
  float beta1;     /* 0.9   TF AdamOptimizer default                              */
  float beta2;     /* 0.999                                                       */
  float eps;       /* 1e-8                                                        */
  float alpha;     /* --alpha         parse.py:16                                 */
  float beta;      /* --beta          parse.py:18                                 */
  float decay;     /* --regs          parse.py:40 (LightGCN: regs[0])             */
  int32_t batch_size_flag; /* args.batch_size: divisor of the L2 term, model.py:220 */
} macr_hparams;

int macr_mf_trainer_create(macr_mf_trainer **out,
                           float *U, float *mU, float *vU, int64_t n_users,
                           float *I, float *mI, float *vI, int64_t n_items,
                           float *w, float *mw, float *vw,
                           float *w_user, float *mwu, float *vwu,
                           int d, int max_batch, const macr_hparams *hp, macr_stream_t stream);
/* ids: device int32 [B] each.  losses_out: device float[4] =
 *   {loss, mf_loss, reg_loss, L_ori} written on the stream (no host sync).      */
int macr_mf_trainer_step(macr_mf_trainer *h, const int32_t *users, const int32_t *pos,
                         const int32_t *neg, int B, float *losses_out);
/* same, ids in host memory (pinned recommended): H2D copy, step, D2H of 3 floats,
 * stream-synchronised before return -- the session-style call of train.py:492.   */
int macr_mf_trainer_step_host(macr_mf_trainer *h, const int32_t *users_host,
                              const int32_t *pos_host, const int32_t *neg_host, int B,
                              float *losses_host /*[3] loss, mf, reg*/);
/* epoch mode (the loop of train.py:470-499 with the batches pre-staged in HBM):
 * batches: device int32 [n_steps][3][B] (users | pos | neg per step), losses: device
 * float [n_steps][4].  Replays the captured step graph n_steps times; asynchronous.  */
int macr_mf_trainer_run(macr_mf_trainer *h, const int32_t *batches, int n_steps, int B,
                        float *losses);
/* the same epoch with the batches in HOST memory (pinned recommended): one H2D copy of
 * [n_steps][3][B] ids, n_steps graph replays, one D2H copy of the [n_steps][4] losses
 * {loss, mf_loss, reg_loss, L_ori}, stream-synchronised before return.  The sampler of
 * train.py:471 consumes only its own RNG streams, so staging an epoch of sampled batches
 * and running them in one call gives the results of the per-step loop.              */
int macr_mf_trainer_run_host(macr_mf_trainer *h, const int32_t *batches_host, int n_steps, int B,
                             float *losses_host);
/* Which `--train` graph the trainer steps (default MACR_TRAIN_RUBIBCEBOTH, the MACR hot path).
 * MACR_TRAIN_NORMALBCE is the README's baseline command (README.md:30, model.py:277-287,:100):
 * element-wise BCE on (y_pos, y_neg) with "+1e-9", same L2 term, TF Adam on the two embedding
 * tables only (w, w_user are not part of that graph).
 * MACR_TRAIN_RUBIBCE is `--train rubibce` (model.py:158-183, :83-85; LightGCN `--loss bce1`,
 * LightGCN.py:431-461, :190-194): the item gate only -- grid P[i,j] = y_pos[j]*sig(s_pos[i]),
 * mf = L_ori + alpha*L_item, TF Adam on the two tables and w; w_user and its slots are never
 * written.  losses[] keep their layout (loss, mf, reg|emb, L_ori).
 * Call before the first step of a run. */
#define MACR_TRAIN_RUBIBCEBOTH 0
#define MACR_TRAIN_NORMALBCE 1
#define MACR_TRAIN_RUBIBCE 2
int macr_mf_trainer_set_mode(macr_mf_trainer *h, int mode);
/* number of this library's kernels launched by one step (for bench gpu_launches) */
int macr_mf_trainer_launches_per_step(const macr_mf_trainer *h);
int64_t macr_mf_trainer_steps_done(const macr_mf_trainer *h);
/* overwrite the Adam step counter (checkpoint resume) */
int macr_mf_trainer_set_steps_done(macr_mf_trainer *h, int64_t t);
int macr_mf_trainer_destroy(macr_mf_trainer *h);

/* ------------------------------------------------------------------------- *
 * K6  CSR SpMM  Y = A X  (A = D^-1/2 A D^-1/2, float32 CSR, sorted columns)
 * replaces: 100 row-folds of tf.sparse_tensor_dense_matmul per layer
 *   macr_lightgcn/LightGCN.py:257-269,:297-305.
 *  mean_accum (nullable): mean_accum = (accum_init ? 0 : mean_accum) ... see
 *  macr_lgcn_propagate for the fused layer mean.
 * ------------------------------------------------------------------------- */
int macr_spmm_csr(const int32_t *rowptr, const int32_t *col, const float *val, int64_t n_rows,
                  const float *X, int d, float *Y, macr_stream_t stream);
/* E_mean = (E0 + A E0 + ... + A^L E0)/(L+1); E0 = concat(U, I) read in place.
 * replaces LightGCN._create_lightgcn_embed LightGCN.py:288-309.
 * tmp: 2*N*d floats scratch.  Emean: N*d floats out (rows [0,n_users) users).   */
int macr_lgcn_propagate(const int32_t *rowptr, const int32_t *col, const float *val,
                        const float *U, int64_t n_users, const float *I, int64_t n_items,
                        int d, int n_layers, float *Emean, float *tmp, macr_stream_t stream);
/* The same two operations on a static work decomposition of the adjacency: every row is cut
 * once into segments of <= 64 nonzeros (one half-warp each; rows of several segments are summed
 * in segment order), so popular items (27 504 nonzeros in ml_10m) do not serialise the launch.
 * The adjacency of a run never changes (LightGCN.py:257-269 builds it once): create the plan
 * once from the device rowptr (one D2H copy, host-side construction), reuse it for every call.
 * Results differ from the unplanned entry points only in the rounding of long rows. */
typedef struct macr_spmm_plan macr_spmm_plan;
int macr_spmm_plan_create(const int32_t *rowptr, int64_t n_rows, macr_spmm_plan **out);
int macr_spmm_plan_destroy(macr_spmm_plan *plan);
int macr_spmm_csr_planned(const macr_spmm_plan *plan, const int32_t *rowptr, const int32_t *col,
                          const float *val, int64_t n_rows, const float *X, int d, float *Y,
                          macr_stream_t stream);
int macr_lgcn_propagate_planned(const macr_spmm_plan *plan, const int32_t *rowptr,
                                const int32_t *col, const float *val, const float *U,
                                int64_t n_users, const float *I, int64_t n_items, int d,
                                int n_layers, float *Emean, float *tmp, macr_stream_t stream);

typedef struct macr_lgcn_trainer macr_lgcn_trainer;
int macr_lgcn_trainer_create(macr_lgcn_trainer **out,
                             const int32_t *rowptr, const int32_t *col, const float *val,
                             float *U, float *mU, float *vU, int64_t n_users,
                             float *I, float *mI, float *vI, int64_t n_items,
                             float *w, float *mw, float *vw,
                             float *w_user, float *mwu, float *vwu,
                             int d, int n_layers, int max_batch, const macr_hparams *hp,
                             macr_stream_t stream);
/* losses_out: device float[4] = {loss, mf_loss, emb_loss, L_ori}
 * (LightGCN.py:598-607 fetch list; reg_loss is the constant 0 of :530).
 * train=0 computes the losses only (train_thread_test, LightGCN.py:616-647).    */
int macr_lgcn_trainer_step(macr_lgcn_trainer *h, const int32_t *users, const int32_t *pos,
                           const int32_t *neg, int B, int train, float *losses_out);
int macr_lgcn_trainer_step_host(macr_lgcn_trainer *h, const int32_t *users_host,
                                const int32_t *pos_host, const int32_t *neg_host, int B,
                                int train, float *losses_host /*[3]*/);
int macr_lgcn_trainer_run(macr_lgcn_trainer *h, const int32_t *batches, int n_steps, int B,
                          int train, float *losses);
/* host-memory variant of macr_lgcn_trainer_run, see macr_mf_trainer_run_host */
int macr_lgcn_trainer_run_host(macr_lgcn_trainer *h, const int32_t *batches_host, int n_steps,
                               int B, int train, float *losses_host);
/* MACR_TRAIN_RUBIBCEBOTH = `--loss bceboth` (default), MACR_TRAIN_RUBIBCE = `--loss bce1`,
 * MACR_TRAIN_NORMALBCE = `--loss bce`
 * (README.md:59, LightGCN.py:415-429,:186); see macr_mf_trainer_set_mode */
int macr_lgcn_trainer_set_mode(macr_lgcn_trainer *h, int mode);
/* propagated tables of the current parameters (device, owned by the handle):
 * users at Emean, items at Emean + n_users*d                                     */
int macr_lgcn_trainer_embeddings(macr_lgcn_trainer *h, const float **Emean);
int macr_lgcn_trainer_launches_per_step(const macr_lgcn_trainer *h);
int64_t macr_lgcn_trainer_steps_done(const macr_lgcn_trainer *h);
int macr_lgcn_trainer_set_steps_done(macr_lgcn_trainer *h, int64_t t);
int macr_lgcn_trainer_destroy(macr_lgcn_trainer *h);

/* ------------------------------------------------------------------------- *
 * K7+K8  full-catalogue counterfactual score + train-item mask + top-K
 * replaces: sess.run(model.rubi_ratings_both, {users: batch, pos_items: range(I)})
 *   macr_mf/train.py:249-251 / utility/batch_test.py:85-88, i.e. model.py:45,:199
 *   S[t,i] = ((u_t . i_i) - c) * sig(i_i . w) * sig(u_t . w_user), then the mask
 *   (train.py:133 set difference / batch_test.py:124-129 -inf) and the per-row
 *   top-K (train.py:95 heapq.nlargest, tools.h:13-22 partial_sort_copy).
 *
 *  macr_score_gates: sig_i[i] = sigmoid(I_i . w), sig_u[t] = sigmoid(U_q[t] . w_user)
 *    dot = fp32 FMA chain k=0..63, sigmoid evaluated in fp64 and rounded to fp32
 *    (so the CPU oracle reproduces the gates bit for bit).
 *  macr_score_topk:
 *    Uq [T][d] query-user rows (already gathered), It [n_items][d] this rank's item
 *    rows whose global ids start at item_id_offset; y = fp32 FMA chain k=0..d-1;
 *    score = ((y - c) * sig_i) * sig_u.  mask_rowptr/mask_col: CSR over the T query
 *    users of GLOBAL item ids to exclude, columns ascending inside a row (nullable).
 *    Order: score descending, ties -> lower item id first.  Rows with fewer than K
 *    unmasked items are padded with id -1 / score -inf.
 *    out_ids [T][K] global ids, out_scores [T][K].
 * ------------------------------------------------------------------------- */
int macr_score_gates(const float *rows, int64_t n, int d, const float *wvec, float *sig_out,
                     macr_stream_t stream);
int macr_gather_rows(const float *table, const int32_t *ids, int n, int d, float *out,
                     macr_stream_t stream);
size_t macr_score_topk_workspace_bytes(int T, int64_t n_items, int K);
int macr_score_topk(const float *Uq, int T, const float *It, int64_t n_items, int d,
                    const float *sig_i, const float *sig_u, float c,
                    const int32_t *mask_rowptr, const int32_t *mask_col,
                    int K, int32_t item_id_offset,
                    int32_t *out_ids, float *out_scores,
                    void *ws, size_t ws_bytes, macr_stream_t stream);
/* Same contract and bit-identical results as macr_score_topk, computed on the tensor cores:
 * two TMA-fed tcgen05 (kind::f16, bf16 operands, fp32 accumulators in TMEM) passes over the
 * T x n_items tile grid -- per-batch maxima -> a proven per-row threshold -> candidate filter
 * (train items of the row are masked inside the pass: mask columns MUST ascend inside a row) --
 * then an exact fp32 re-rank of the few dozen candidates per row.  Rows whose candidate list
 * overflows, rows without K unmasked groups and rows whose train list exceeds 1/16 of the
 * catalogue are done by the exact kernel (macr_b200/csrc/score_tc.cu, score.cu).  Needs
 * 2048 <= n_items <~ 1.5 M per call (smaller catalogues: macr_score_topk; larger: shard).
 * ws must be 1024-byte aligned.  stats (nullable, device int64[2]) is incremented by
 * {rows done by the exact kernel, candidates re-ranked}. */
size_t macr_score_topk_tc_workspace_bytes(int T, int64_t n_items, int K);
int macr_score_topk_tc(const float *Uq, int T, const float *It, int64_t n_items, int d,
                       const float *sig_i, const float *sig_u, float c,
                       const int32_t *mask_rowptr, const int32_t *mask_col,
                       int K, int32_t item_id_offset,
                       int32_t *out_ids, float *out_scores,
                       void *ws, size_t ws_bytes, int64_t *stats, macr_stream_t stream);
/* The item-side operands of that path (bf16 rows scaled by sig_i, the -c*sig_i pieces, the largest
 * item norm) depend on (It, sig_i, c) alone.  An evaluation scores many query blocks against ONE
 * model (train.py:174-180 walks the test users in batches, batch_test.py:38-43 likewise), so the
 * preparation can be done once per evaluation instead of once per call:
 *   macr_score_tc_prepare_items   fills items_ws (macr_score_tc_items_bytes(n_items) bytes,
 *                                 1024-byte aligned)
 *   macr_score_topk_tc_prepared   macr_score_topk_tc without the item preparation; It, sig_i, c,
 *                                 n_items must be the ones items_ws was prepared from (It and
 *                                 sig_i are still read: the re-rank recomputes exact fp32 scores);
 *                                 ws: macr_score_topk_tc_prepared_workspace_bytes(T, n_items, K).
 * Same ids and scores, bit for bit, as macr_score_topk_tc. */
size_t macr_score_tc_items_bytes(int64_t n_items);
int macr_score_tc_prepare_items(const float *It, int64_t n_items, int d, const float *sig_i,
                                float c, void *items_ws, size_t items_bytes, macr_stream_t stream);
size_t macr_score_topk_tc_prepared_workspace_bytes(int T, int64_t n_items, int K);
int macr_score_topk_tc_prepared(const float *Uq, int T, const float *It, int64_t n_items, int d,
                                const float *sig_i, const float *sig_u, float c,
                                const int32_t *mask_rowptr, const int32_t *mask_col,
                                int K, int32_t item_id_offset,
                                int32_t *out_ids, float *out_scores, const void *items_ws,
                                void *ws, size_t ws_bytes, int64_t *stats, macr_stream_t stream);
/* dense score matrix, the literal rubi_ratings_both fetch ([T][n_items] fp32, no mask) */
int macr_score_matrix(const float *Uq, int T, const float *It, int64_t n_items, int d,
                      const float *sig_i, const float *sig_u, float c,
                      float *out, macr_stream_t stream);
/* merge G candidate lists per row (multi-GPU shards after the all-gather, or the
 * item-chunk partials inside one GPU): in [G][T][K] -> out [T][K], same order rule */
int macr_topk_merge(const int32_t *ids, const float *scores, int T, int K, int G,
                    int32_t *out_ids, float *out_scores, macr_stream_t stream);

/* ------------------------------------------------------------------------- *
 * top-K of a caller-supplied score matrix: drop-in for
 *   void c_top_k_array_index(float*,int columns,int rows,int top_k,int threads,int* rankings)
 *   macr_lightgcn/evaluator/cpp/include/tools.h:24-33 (device pointers here).
 * ------------------------------------------------------------------------- */
int macr_topk_rows(const float *scores, int columns_num, int rows_num, int top_k,
                   int32_t *rankings, macr_stream_t stream);

/* ------------------------------------------------------------------------- *
 * K9  fold-out metrics from top-K ids; output layout identical to
 *   void evaluate_foldout(int users_num,int* rankings,int rank_len,int** ground_truths,
 *                         int* ground_truths_num,int thread_num,float* results)
 *   macr_lightgcn/evaluator/cpp/include/evaluate_foldout.h:115-195:
 *   out[t][0:K]=precision@1..K, [K:2K]=recall, [2K:3K]=AP, [3K:4K]=NDCG, [4K:5K]=MRR.
 *   ground truth as CSR (truth_rowptr[T+1], truth_col) instead of int**.
 *   inv_log2[k] = 1.0/log2(k+2) as double, k<K, computed by the HOST libm so the
 *   float accumulation matches the reference bit for bit.
 * ------------------------------------------------------------------------- */
int macr_foldout_metrics(const int32_t *topk_ids, int T, int K,
                         const int32_t *truth_rowptr, const int32_t *truth_col,
                         const double *inv_log2, float *out, macr_stream_t stream);

/* ------------------------------------------------------------------------- *
 * Batch samplers (host code, no GPU work): bit-exact twins of
 *   Data.sample()  macr_mf/load_data.py:543-566            (CPython `random`)
 *   Data.sample()  macr_lightgcn/utility/load_data.py:174-212
 *                  (`random.sample` for users, legacy `np.random.randint(size=1)` for items)
 * py_state / np_state: uint32[625] = the 624 Mersenne-Twister words + the position, i.e.
 * random.getstate()[1] and np.random.get_state()[1:3]; advanced in place -- put them back with
 * random.setstate / np.random.set_state and the interpreter-side generators continue exactly
 * where the reference's would.  users_pop: the population list (`self.users` /
 * `self.exist_users`, list order); n_users: the `batch_size <= n_users` test of the reference.
 * rowptr/order: per-user positive lists in the reference's list order (int64 CSR over user id);
 * sorted / ban_sorted: the ids the negative draw rejects, ascending per user.
 * Out: users, pos, neg int32[B].
 * ------------------------------------------------------------------------- */
int macr_sample_mf(uint32_t *py_state, const int32_t *users_pop, int n_pop, int n_users,
                   int n_items, const int64_t *rowptr, const int32_t *order,
                   const int32_t *sorted, int B, int32_t *users, int32_t *pos, int32_t *neg);
int macr_sample_lgcn(uint32_t *py_state, uint32_t *np_state, const int32_t *users_pop, int n_pop,
                     int n_users, int n_items, const int64_t *pos_rowptr,
                     const int32_t *pos_order, const int64_t *ban_rowptr,
                     const int32_t *ban_sorted, int B, int32_t *users, int32_t *pos,
                     int32_t *neg);

/* Epoch forms: n_batches consecutive calls of the functions above in one go, out = int32
 * [n_batches][3][B] (users, pos, neg per batch -- the layout macr_mf_trainer_run_host stages).
 * Same streams and triples word for word (the epoch loops of macr_mf/train.py:470-499 and
 * macr_lightgcn/LightGCN.py:765-773 call sample() once per step), but drawn speculatively in
 * chunks so the sequential part never waits for the lists, by branch-free word-driven loops
 * and, for sparse lists, a second thread that verifies behind the draws:
 * 71 M triples/s against ~14 M for the functions above.
 * tags: hashed (user, rejected id) pair set over rowptr / sorted (ban_rowptr / ban_sorted):
 * uint16[8 << log2_buckets], 16-byte aligned, filled by macr_pairset_build (size it for ~2-3
 * pairs per bucket; a full bucket only costs exact look-ups, never a wrong answer). */
int macr_pairset_build(const int64_t *rowptr, const int32_t *ids, int n_rows, uint16_t *tags,
                       int log2_buckets);
int macr_sample_mf_epoch(uint32_t *py_state, const int32_t *users_pop, int n_pop, int n_users,
                         int n_items, const int64_t *rowptr, const int32_t *order,
                         const int32_t *sorted, const uint16_t *tags, int log2_buckets, int B,
                         int n_batches, int32_t *out);
int macr_sample_lgcn_epoch(uint32_t *py_state, uint32_t *np_state, const int32_t *users_pop,
                           int n_pop, int n_users, int n_items, const int64_t *pos_rowptr,
                           const int32_t *pos_order, const int64_t *ban_rowptr,
                           const int32_t *ban_sorted, const uint16_t *ban_tags, int log2_buckets,
                           int B, int n_batches, int32_t *out);

/* ------------------------------------------------------------------------- *
 * Row-partitioned training over the GPUs of one box (SURVEY.md 8e rows "dense Adam
 * sweep" and "gather + grid + row grads"): every rank owns a contiguous id range
 * of both tables (with their Adam slots); the rows of the batch are exchanged once
 * per step and the unchanged single-GPU step runs on local ids.
 * replaces (distributed form of): tf.nn.embedding_lookup x3  macr_mf/model.py:35-37
 *   on a variable that lives on one device in the reference.
 * Local table layout: [n_local owned rows | 2 parities x max_batch (users) or
 *   2*max_batch (items) ghost rows]; ghost slot of batch position b: users b, pos b,
 *   neg B+b.  local_ids3 = users|pos|neg renumbered (owned: id-lo, foreign: ghost).
 *  macr_shard_pack   : owned rows -> ex[3B][64], zeros elsewhere (sum over ranks by
 *                      the caller, e.g. ncclAllReduce; exact: one owner per row)
 *  macr_shard_unpack : ex -> ghost slots (two device copies)
 *  macr_shard_push   : the fused form: owned rows are stored straight into the ghost
 *                      slots of every peer over NVLink; peer_*_ghost_host[r] = address
 *                      in THIS process (macr_ipc_open) of row n_local of rank r's table
 *  macr_shard_barrier: flag barrier over peer memory after the push; peer_flags_host[r]
 *                      = rank r's uint64[world] flag array (zero-initialised, IPC memory);
 *                      epoch must grow by one per call; *err_flag (device int) becomes
 *                      1+r if peer r did not arrive within ~10 s (no device hang).
 * macr_ipc_* : cudaMalloc'ed memory with its cudaIpcMemHandle_t (64 opaque bytes) so the
 *   other ranks of the box can map it.
 * ------------------------------------------------------------------------- */
#define MACR_SHARD_MAX_RANKS 16
#define MACR_IPC_HANDLE_BYTES 64
typedef struct macr_shard_desc {
  int32_t rank, world;
  int64_t u_lo, u_hi, i_lo, i_hi; /* owned global id ranges [lo, hi) */
  int32_t max_batch;
} macr_shard_desc;
int macr_shard_pack(const float *U_local, const float *I_local, const macr_shard_desc *desc,
                    const int32_t *ids3, int B, int parity, int32_t *local_ids3, float *ex,
                    macr_stream_t stream);
int macr_shard_unpack(float *U_local, float *I_local, const macr_shard_desc *desc,
                      const float *ex, int B, int parity, macr_stream_t stream);
int macr_shard_push(const float *U_local, const float *I_local, const macr_shard_desc *desc,
                    const int32_t *ids3, int B, int parity, int32_t *local_ids3,
                    float *const *peer_U_ghost_host, float *const *peer_I_ghost_host,
                    macr_stream_t stream);
int macr_shard_barrier(uint64_t *const *peer_flags_host, int rank, int world, uint64_t epoch,
                       int *err_flag, macr_stream_t stream);
/* The same exchange INSIDE the MF step graph (one rank per GPU): create the trainer on the local
 * tables (owned rows + ghost rows, U / I in macr_ipc_alloc memory), export the handle of its flag
 * array, then hand it every peer's ghost bases and flags before the first step.  From then on the
 * ids given to macr_mf_trainer_step / _run / _run_host are GLOBAL ids, identical on every rank; the
 * captured step renumbers them, stores the owned rows into the peers' ghost slots, lets the dense
 * sweep start at once and waits for the peers' rows on the gather branch only.
 * macr_mf_trainer_peer_error: synchronises; *err_out = 1 + r if peer r missed a barrier. */
int macr_mf_trainer_ipc_export(macr_mf_trainer *h, unsigned char handle[MACR_IPC_HANDLE_BYTES]);
int macr_mf_trainer_shard(macr_mf_trainer *h, const macr_shard_desc *desc,
                          float *const *peer_U_ghost, float *const *peer_I_ghost,
                          uint64_t *const *peer_flags);
int macr_mf_trainer_peer_error(macr_mf_trainer *h, int *err_out);

/* Row-partitioned LightGCN (SURVEY.md 8e row "LightGCN SpMM": 1-D row partition of A_hat, an
 * all-gather of E_k per layer): rank r owns user rows [u_lo,u_hi) and item rows [i_lo,i_hi).
 * replaces (distributed form of): _create_lightgcn_embed's row folds
 *   macr_lightgcn/LightGCN.py:257-269,297-305, and the dense ApplyAdam of :201 on the owned rows.
 * Create the trainer with the FULL-size tables (U, I in macr_ipc_alloc memory; only the owned rows
 * and their Adam slots are kept current here, rows of other ranks are refreshed by their owners)
 * and a CSR that holds the nonzeros of the owned rows only (rowptr still has N+1 entries).
 * macr_lgcn_trainer_ipc_export hands out the IPC handles of {E_mean, layer buffers, flags};
 * macr_lgcn_trainer_shard (before the first step) takes every peer's mapping of {U, I, E_mean,
 * layer buffers, flags}.  From then on each layer is computed for the owned rows and stored into
 * the peers' buffers over NVLink (peer stores + flag barrier, inside the step's CUDA graph); owned
 * rows, losses, w and w_user are bit-identical to the single-GPU trainer.
 * macr_lgcn_trainer_peer_error: synchronises; *err_out = 1 + r if peer r missed a barrier. */
int macr_lgcn_trainer_ipc_export(macr_lgcn_trainer *h, unsigned char handles[3][MACR_IPC_HANDLE_BYTES]);
int macr_lgcn_trainer_shard(macr_lgcn_trainer *h, const macr_shard_desc *desc,
                            float *const *peer_U, float *const *peer_I, float *const *peer_Emean,
                            float *const *peer_tmp, uint64_t *const *peer_flags);
int macr_lgcn_trainer_peer_error(macr_lgcn_trainer *h, int *err_out);
/* Item-partitioned scoring (SURVEY.md 8e row "full-catalogue scoring"): exchange + merge of the
 * shards' [T][K] candidate lists in ONE kernel over NVLink peer memory.  This rank merges the
 * query rows [row0, row0 + rows): cand_*_host[r] = rank r's candidate buffers ([T][K], ids global,
 * lists sorted and padded with -1 / -inf) as mapped in THIS process (macr_ipc_open; own buffers for
 * r == rank), out_*_host[r] = rank r's result buffers: the merged rows are stored into all of them.
 * Same order rule as macr_topk_merge (score desc, lower id first; shards in rank order), so the
 * result equals the unsharded call bit for bit.  Bracket it with two macr_shard_barrier calls
 * (candidates of every rank complete / merged rows landed everywhere). */
int macr_topk_merge_peers(const int32_t *const *cand_ids_host, const float *const *cand_scores_host,
                          int32_t *const *out_ids_host, float *const *out_scores_host, int world,
                          int K, int row0, int rows, macr_stream_t stream);
int macr_ipc_alloc(size_t bytes, void **dev_ptr, unsigned char handle_out[MACR_IPC_HANDLE_BYTES]);
int macr_ipc_open(const unsigned char handle[MACR_IPC_HANDLE_BYTES], void **peer_ptr);
int macr_ipc_close(void *peer_ptr);
int macr_ipc_free(void *dev_ptr);

#ifdef __cplusplus
}
#endif
#endif /* MACR_B200_H */
