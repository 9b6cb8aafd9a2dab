/*
 * macr_oracle.h -- CPU restatement of MACR's hot path.  TEST INFRASTRUCTURE ONLY.
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference
 * legs may load this library.  The product (macr_b200/) never links or calls it.
 *
 * PARITY STATUS: "parity unpinned" at the TensorFlow boundary.  The reference's
 * numeric core is TF-1.14 graph code (macr_mf/model.py, macr_lightgcn/LightGCN.py);
 * TF is not installable here and the reference ships no tests or golden vectors
 * (SURVEY.md section 4, section 8c).  What IS pinned:
 *   - the closed-form gradients below against torch autograd of the literal
 *     restatement in oracle/literal_torch.py (tests/test_oracle_cpu.py);
 *   - top-K + fold-out metrics against the reference's own C++ evaluator compiled
 *     from /root/reference into oracle/_ref/ (tests/test_oracle_vs_ref.py);
 *   - samplers / adjacency against the reference's own Python imported from
 *     /root/reference (tests/golden/, tests/golden/make_golden.py).
 *
 * Arithmetic convention: every tensor the reference materialises is fp32; reductions
 * whose order TF does not specify are accumulated in double and rounded once.
 */
#ifndef MACR_ORACLE_H
#define MACR_ORACLE_H
#include <stdint.h>
#ifdef __cplusplus
extern "C" {
#endif

typedef struct oracle_hparams {
  float lr, beta1, beta2, eps, alpha, beta, decay;
  int32_t batch_size_flag;
} oracle_hparams;

void oracle_set_threads(int n);
int oracle_get_threads(void);

/* macr_mf/model.py:35-37,186-187,194-196,219 */
void oracle_gather_dots(const float *Ue, const float *Ie, const float *Ur, const float *Ir,
                        const float *w, const float *wu, const int32_t *u, const int32_t *p,
                        const int32_t *n, int B, int d, float *yp, float *yn, float *sp,
                        float *sn, float *su, float *regsq);

/* macr_mf/model.py:204-217 and the gradient TF autodiff would produce */
void oracle_grid_bce(const float *yp, const float *yn, const float *sp, const float *sn,
                     const float *su, int B, float alpha, float beta, float *losses3,
                     float *d_yp, float *d_yn, float *d_sp, float *d_sn, float *d_su);

/* TF-1.14 adam.py _apply_sparse_shared after optimizer.py _deduplicate_indexed_slices */
void oracle_adam_sparse(float *var, float *m, float *v, int64_t rows, int d,
                        const int32_t *idx, const float *grad_rows, int n_idx, float lr_t,
                        float beta1, float beta2, float eps);
/* TF-1.14 training_ops.cc ApplyAdam (dense variables w, w_user) */
void oracle_adam_dense_vec(float *var, float *m, float *v, const float *g, int n, float lr_t,
                           float beta1, float beta2, float eps);
float oracle_adam_lr_t(float lr, float beta1_power, float beta2_power);

/* one full `--train rubibceboth` step, macr_mf/train.py:492-496.
 * pw[2] = {beta1_power, beta2_power} in/out.  losses[4] = {loss, mf, reg, L_ori}. */
void oracle_mf_step(float *U, float *mU, float *vU, int64_t n_users, float *I, float *mI,
                    float *vI, int64_t n_items, float *w, float *mw, float *vw, float *wu,
                    float *mwu, float *vwu, int d, const int32_t *u, const int32_t *p,
                    const int32_t *n, int B, const oracle_hparams *hp, float *pw,
                    float *losses);
/* `--train normalbce` (model.py:277-287, :100): element-wise BCE, Adam on the two tables only */
void oracle_plain_bce(const float *yp, const float *yn, int B, float *mf_loss, float *dyp,
                      float *dyn);
void oracle_mf_step_normal(float *U, float *mU, float *vU, int64_t n_users, float *I, float *mI,
                           float *vI, int64_t n_items, int d, const int32_t *u, const int32_t *p,
                           const int32_t *n, int B, const oracle_hparams *hp, float *pw,
                           float *losses);

/* `--train rubibce` (model.py:158-183, :83-85): item gate only -- grid P[i,j] = yp_j*sig(sp_i),
 * mf = L_ori + alpha*L_item, Adam on the two tables and w; w_user is never written */
void oracle_grid_bce_item(const float *yp, const float *yn, const float *sp, const float *sn,
                          int B, float alpha, float *losses3, float *d_yp, float *d_yn,
                          float *d_sp, float *d_sn);
void oracle_mf_step_item(float *U, float *mU, float *vU, int64_t n_users, float *I, float *mI,
                         float *vI, int64_t n_items, float *w, float *mw, float *vw, float *wu,
                         int d, const int32_t *u, const int32_t *p, const int32_t *n, int B,
                         const oracle_hparams *hp, float *pw, float *losses);

/* macr_lightgcn/LightGCN.py:297-305 (one layer, all folds) */
void oracle_spmm_csr(const int32_t *rowptr, const int32_t *col, const float *val,
                     int64_t n_rows, const float *X, int d, float *Y);
/* Y = A^T X for a general CSR A (gradient of the above) */
void oracle_spmm_csr_t(const int32_t *rowptr, const int32_t *col, const float *val,
                       int64_t n_rows, const float *X, int d, float *Y);
/* LightGCN.py:288-309 */
void oracle_lgcn_propagate(const int32_t *rowptr, const int32_t *col, const float *val,
                           const float *U, int64_t n_users, const float *I, int64_t n_items,
                           int d, int n_layers, float *Emean);
/* one `--loss bceboth` step, LightGCN.py:598-607; train=0 -> losses only (:616-647).
 * losses[4] = {loss, mf, emb, L_ori} */
void oracle_lgcn_step(const int32_t *rowptr, const int32_t *col, const float *val, float *U,
                      float *mU, float *vU, int64_t n_users, float *I, float *mI, float *vI,
                      int64_t n_items, float *w, float *mw, float *vw, float *wu, float *mwu,
                      float *vwu, int d, int n_layers, const int32_t *u, const int32_t *p,
                      const int32_t *n, int B, int train, const oracle_hparams *hp, float *pw,
                      float *losses);
/* `--loss bce` (LightGCN.py:415-429,:186) */
void oracle_lgcn_step_normal(const int32_t *rowptr, const int32_t *col, const float *val, float *U,
                             float *mU, float *vU, int64_t n_users, float *I, float *mI,
                             float *vI, int64_t n_items, float *w, float *wu, int d, int L,
                             const int32_t *u, const int32_t *p, const int32_t *n, int B, int train,
                             const oracle_hparams *hp, float *pw, float *losses);

/* `--loss bce1` (LightGCN.py:431-461, :190-194) */
void oracle_lgcn_step_item(const int32_t *rowptr, const int32_t *col, const float *val, float *U,
                           float *mU, float *vU, int64_t n_users, float *I, float *mI, float *vI,
                           int64_t n_items, float *w, float *mw, float *vw, float *wu, int d,
                           int L, const int32_t *u, const int32_t *p, const int32_t *n, int B,
                           int train, const oracle_hparams *hp, float *pw, float *losses);

/* model.py:199 scoring */
void oracle_score_gates(const float *rows, int64_t n, int d, const float *wvec, float *sig);
void oracle_score_matrix(const float *Uq, int T, const float *It, int64_t n_items, int d,
                         const float *sig_i, const float *sig_u, float c, float *out);
/* mask + top-K (score desc, lower id first); ids are offset by item_id_offset */
void oracle_score_topk(const float *Uq, int T, const float *It, int64_t n_items, int d,
                       const float *sig_i, const float *sig_u, float c,
                       const int32_t *mask_rowptr, const int32_t *mask_col, int K,
                       int32_t item_id_offset, int32_t *out_ids, float *out_scores);
void oracle_topk_rows(const float *scores, int columns_num, int rows_num, int top_k,
                      int32_t *rankings);
void oracle_topk_merge(const int32_t *ids, const float *scores, int T, int K, int G,
                       int32_t *out_ids, float *out_scores);
/* evaluate_foldout.h:16-195 */
void oracle_foldout_metrics(const int32_t *topk_ids, int T, int K, const int32_t *truth_rowptr,
                            const int32_t *truth_col, float *out);
void oracle_inv_log2_table(int K, double *out);

#ifdef __cplusplus
}
#endif
#endif
