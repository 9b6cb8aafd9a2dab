/*
 * macr_oracle.c -- CPU restatement of MACR's hot path.  TEST INFRASTRUCTURE ONLY
 * (see macr_oracle.h for who may load it and for the parity-pinning status).
 *
 * Written from the cited reference lines (paths relative to /root/reference):
 *   macr_mf/model.py:35-45,59-60,72-74,185-222,313-314
 *   macr_lightgcn/LightGCN.py:145-150,166,197-201,288-309,495-532,554-555
 *   macr_lightgcn/evaluator/cpp/include/tools.h:13-33, evaluate_foldout.h:16-195
 * plus TensorFlow 1.14's published Adam (python/training/adam.py _apply_sparse_shared,
 * python/training/optimizer.py _deduplicate_indexed_slices, core/kernels/training_ops.cc
 * ApplyAdam) -- third-party, pinned only in prose at README.md:10, not vendored.
 *
 * Build: gcc -O2 -ffp-contract=off -fopenmp -shared -fPIC (oracle/Makefile).
 * -ffp-contract=off keeps every fp32 op individually rounded like TF's Eigen kernels.
 */
#include "macr_oracle.h"

#include <math.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

static int g_threads = 0;

void oracle_set_threads(int n) {
  g_threads = n;
#ifdef _OPENMP
  if (n > 0) omp_set_num_threads(n);
#endif
}
int oracle_get_threads(void) {
#ifdef _OPENMP
  return g_threads > 0 ? g_threads : omp_get_max_threads();
#else
  return 1;
#endif
}

static const float kEps = 1e-10f; /* model.py:211 "+1e-10" on an fp32 tensor */

static inline float sigmoidf_(float x) { return 1.0f / (1.0f + expf(-x)); }

static float dot_d(const float *a, const float *b, int d) {
  double s = 0.0;
  for (int k = 0; k < d; ++k) s += (double)a[k] * (double)b[k];
  return (float)s;
}

/* ---- model.py:35-37 gathers, :186-187 row dots, :194-196 branch matmuls, :219 l2 ---- */
void oracle_gather_dots(const float *Ue, const float *Ie, const float *Ur, const float *Ir,
                        const float *w, const float *wu, const int32_t *u, const int32_t *p,
                        const int32_t *n, int B, int d, float *yp, float *yn, float *sp,
                        float *sn, float *su, float *regsq) {
#pragma omp parallel for schedule(static)
  for (int b = 0; b < B; ++b) {
    const float *ue = Ue + (int64_t)u[b] * d, *pe = Ie + (int64_t)p[b] * d,
                *ne = Ie + (int64_t)n[b] * d;
    yp[b] = dot_d(ue, pe, d);
    yn[b] = dot_d(ue, ne, d);
    sp[b] = dot_d(pe, w, d);
    sn[b] = dot_d(ne, w, d);
    su[b] = dot_d(ue, wu, d);
    const float *ur = Ur + (int64_t)u[b] * d, *pr = Ir + (int64_t)p[b] * d,
                *nr = Ir + (int64_t)n[b] * d;
    double s = 0.0;
    for (int k = 0; k < d; ++k)
      s += (double)ur[k] * ur[k] + (double)pr[k] * pr[k] + (double)nr[k] * nr[k];
    regsq[b] = (float)s;
  }
}

/* ---- model.py:204-217: the [B]*[B,1] broadcast grid, its three means, and the gradient
 *      of  L_ori + alpha*L_item + beta*L_user  w.r.t. the five [B] vectors ---- */
/* item_only != 0: `--train rubibce` / `--loss bce1` (model.py:158-183; LightGCN.py:431-461): the
 * grid is pos_scores*sigmoid(pos_item_scores) only -- no user branch in the product, no L_user,
 * nothing flows to user_scores / w_user */
static void grid_bce_impl(const float *yp, const float *yn, const float *sp, const float *sn,
                          const float *su, int B, float alpha, float beta, float *losses3,
                          float *d_yp, float *d_yn, float *d_sp, float *d_sn, float *d_su,
                          int item_only) {
  float *a = (float *)malloc(sizeof(float) * B * 3);
  float *an = a + B, *g = a + 2 * B;
  for (int i = 0; i < B; ++i) {
    a[i] = sigmoidf_(sp[i]);  /* tf.nn.sigmoid(self.pos_item_scores) */
    an[i] = sigmoidf_(sn[i]); /* tf.nn.sigmoid(self.neg_item_scores) */
    g[i] = item_only ? 1.0f : sigmoidf_(su[i]); /* tf.nn.sigmoid(self.user_scores) */
  }
  double *colP = (double *)calloc((size_t)B * 2, sizeof(double));
  double *colN = colP + B;
  double *rowP = (double *)calloc((size_t)B * 2, sizeof(double));
  double *rowN = rowP + B;
  double loss_sum = 0.0;
  const int want_grad = d_yp != NULL;

#pragma omp parallel
  {
    double *cP = (double *)calloc((size_t)B * 2, sizeof(double));
    double *cN = cP + B;
    double lsum = 0.0;
#pragma omp for schedule(static)
    for (int i = 0; i < B; ++i) {
      const float ai = a[i], ani = an[i], gi = g[i];
      double rp = 0.0, rn = 0.0;
      for (int j = 0; j < B; ++j) {
        /* model.py:204  pos_scores*sigmoid(pos_item_scores)*sigmoid(user_scores),
         * evaluated left to right: ([B]*[B,1])*[B,1] -> element [i,j] */
        const float P = item_only ? yp[j] * ai : (yp[j] * ai) * gi; /* model.py:172 */
        const float N = item_only ? yn[j] * ani : (yn[j] * ani) * gi;
        const float s = sigmoidf_(P);
        const float t = sigmoidf_(N);
        const float sp_e = s + kEps;          /* sigmoid(pos)+1e-10   */
        const float q = (1.0f - t) + kEps;    /* 1-sigmoid(neg)+1e-10 */
        lsum += (double)(-logf(sp_e)) + (double)(-logf(q)); /* model.py:211 */
        if (want_grad) {
          const float dP = -(s * (1.0f - s)) / sp_e;
          const float dN = (t * (1.0f - t)) / q;
          cP[j] += (double)dP * ((double)ai * gi);
          cN[j] += (double)dN * ((double)ani * gi);
          rp += (double)dP * yp[j];
          rn += (double)dN * yn[j];
        }
      }
      rowP[i] = rp;
      rowN[i] = rn;
    }
#pragma omp critical
    {
      loss_sum += lsum;
      for (int j = 0; j < 2 * B; ++j) colP[j] += cP[j];
    }
    free(cP);
  }

  const double invBB = 1.0 / ((double)B * (double)B), invB = 1.0 / (double)B;
  double l_item = 0.0, l_user = 0.0;
  for (int i = 0; i < B; ++i) {
    const float ea = a[i] + kEps, ean = (1.0f - an[i]) + kEps;
    const float eg = g[i] + kEps, eg1 = (1.0f - g[i]) + kEps;
    l_item += (double)(-logf(ea)) + (double)(-logf(ean)); /* model.py:213 */
    if (!item_only) l_user += (double)(-logf(eg)) + (double)(-logf(eg1)); /* model.py:215 */
    if (want_grad) {
      d_yp[i] = (float)(colP[i] * invBB);
      d_yn[i] = (float)(colN[i] * invBB);
      double da = rowP[i] * invBB * g[i] + (double)alpha * invB * (-1.0 / ea);
      double dan = rowN[i] * invBB * g[i] + (double)alpha * invB * (1.0 / ean);
      double dg = (rowP[i] * a[i] + rowN[i] * an[i]) * invBB +
                  (double)beta * invB * (-1.0 / eg + 1.0 / eg1);
      d_sp[i] = (float)(da * ((double)a[i] * (1.0f - a[i])));
      d_sn[i] = (float)(dan * ((double)an[i] * (1.0f - an[i])));
      d_su[i] = item_only ? 0.0f : (float)(dg * ((double)g[i] * (1.0f - g[i])));
    }
  }
  losses3[0] = (float)(loss_sum * invBB);
  losses3[1] = (float)(l_item * invB);
  losses3[2] = (float)(l_user * invB);
  free(a);
  free(colP);
  free(rowP);
}

void oracle_grid_bce(const float *yp, const float *yn, const float *sp, const float *sn,
                     const float *su, int B, float alpha, float beta, float *losses3,
                     float *d_yp, float *d_yn, float *d_sp, float *d_sn, float *d_su) {
  grid_bce_impl(yp, yn, sp, sn, su, B, alpha, beta, losses3, d_yp, d_yn, d_sp, d_sn, d_su, 0);
}

void oracle_grid_bce_item(const float *yp, const float *yn, const float *sp, const float *sn,
                          int B, float alpha, float *losses3, float *d_yp, float *d_yn,
                          float *d_sp, float *d_sn) {
  float *dsu = (float *)malloc(sizeof(float) * (size_t)B);
  grid_bce_impl(yp, yn, sp, sn, sp /*unused*/, B, alpha, 0.0f, losses3, d_yp, d_yn, d_sp, d_sn, dsu,
                1);
  free(dsu);
}

/* ---- TF-1.14 adam.py: lr_t = lr * sqrt(1 - beta2_power) / (1 - beta1_power), fp32 ---- */
float oracle_adam_lr_t(float lr, float b1p, float b2p) {
  return (lr * sqrtf(1.0f - b2p)) / (1.0f - b1p);
}

/* ---- optimizer.py _deduplicate_indexed_slices + adam.py _apply_sparse_shared ---- */
void oracle_adam_sparse(float *var, float *m, float *v, int64_t rows, int d, const int32_t *idx,
                        const float *grad_rows, int n_idx, float lr_t, float beta1, float beta2,
                        float eps) {
  /* array_ops.unique keeps first-occurrence order; unsorted_segment_sum adds in order */
  int32_t *slot = (int32_t *)malloc(sizeof(int32_t) * (size_t)rows);
  memset(slot, 0xff, sizeof(int32_t) * (size_t)rows);
  int32_t *uniq = (int32_t *)malloc(sizeof(int32_t) * (size_t)(n_idx > 0 ? n_idx : 1));
  float *summed = (float *)calloc((size_t)(n_idx > 0 ? n_idx : 1) * d, sizeof(float));
  int nu = 0;
  for (int q = 0; q < n_idx; ++q) {
    int32_t r = idx[q];
    if (slot[r] < 0) {
      slot[r] = nu;
      uniq[nu++] = r;
    }
    float *dst = summed + (int64_t)slot[r] * d;
    const float *src = grad_rows + (int64_t)q * d;
    for (int k = 0; k < d; ++k) dst[k] = dst[k] + src[k];
  }
  const float omb1 = 1.0f - beta1, omb2 = 1.0f - beta2;
  const int64_t n = rows * d;
#pragma omp parallel for schedule(static)
  for (int64_t e = 0; e < n; ++e) { /* m_t = assign(m, m*beta1); v_t = assign(v, v*beta2) */
    m[e] = m[e] * beta1;
    v[e] = v[e] * beta2;
  }
  for (int s = 0; s < nu; ++s) { /* scatter_add of grad*(1-b1) and (grad*grad)*(1-b2) */
    float *mr = m + (int64_t)uniq[s] * d, *vr = v + (int64_t)uniq[s] * d;
    const float *gr = summed + (int64_t)s * d;
    for (int k = 0; k < d; ++k) {
      mr[k] = mr[k] + gr[k] * omb1;
      vr[k] = vr[k] + (gr[k] * gr[k]) * omb2;
    }
  }
#pragma omp parallel for schedule(static)
  for (int64_t e = 0; e < n; ++e) /* var -= lr * m_t / (sqrt(v_t) + eps) */
    var[e] = var[e] - (lr_t * m[e]) / (sqrtf(v[e]) + eps);
  free(slot);
  free(uniq);
  free(summed);
}

/* ---- training_ops.cc ApplyAdam<CPU> ---- */
void oracle_adam_dense_vec(float *var, float *m, float *v, const float *g, int n, float lr_t,
                           float beta1, float beta2, float eps) {
  const float omb1 = 1.0f - beta1, omb2 = 1.0f - beta2;
  for (int e = 0; e < n; ++e) {
    m[e] = m[e] + (g[e] - m[e]) * omb1;
    v[e] = v[e] + (g[e] * g[e] - v[e]) * omb2;
    var[e] = var[e] - (m[e] * lr_t) / (sqrtf(v[e]) + eps);
  }
}

/* row gradients of the batch (what TF hands to Adam as IndexedSlices) */
static void row_grads(const float *Ue, const float *Ie, const float *Ur, const float *Ir,
                      const float *w, const float *wu, const int32_t *u, const int32_t *p,
                      const int32_t *n, int B, int d, const float *d_yp, const float *d_yn,
                      const float *d_sp, const float *d_sn, const float *d_su, float lam,
                      int add_l2, float *gU, float *gPN, float *gw, float *gwu) {
  double *aw = (double *)calloc((size_t)d * 2, sizeof(double));
  double *awu = aw + d;
  for (int b = 0; b < B; ++b) {
    const float *ue = Ue + (int64_t)u[b] * d, *pe = Ie + (int64_t)p[b] * d,
                *ne = Ie + (int64_t)n[b] * d;
    const float *ur = Ur + (int64_t)u[b] * d, *pr = Ir + (int64_t)p[b] * d,
                *nr = Ir + (int64_t)n[b] * d;
    float *gu = gU + (int64_t)b * d, *gp = gPN + (int64_t)b * d,
          *gn = gPN + (int64_t)(B + b) * d;
    for (int k = 0; k < d; ++k) {
      float x = d_yp[b] * pe[k] + d_yn[b] * ne[k] + d_su[b] * wu[k];
      float y = d_yp[b] * ue[k] + d_sp[b] * w[k];
      float z = d_yn[b] * ue[k] + d_sn[b] * w[k];
      if (add_l2) { /* tf.nn.l2_loss grad = x; scaled by decay/batch_size (model.py:219-221) */
        x += lam * ur[k];
        y += lam * pr[k];
        z += lam * nr[k];
      }
      gu[k] = x;
      gp[k] = y;
      gn[k] = z;
      aw[k] += (double)d_sp[b] * pe[k] + (double)d_sn[b] * ne[k];
      awu[k] += (double)d_su[b] * ue[k];
    }
  }
  for (int k = 0; k < d; ++k) {
    gw[k] = (float)aw[k];
    gwu[k] = (float)awu[k];
  }
  free(aw);
}

static void batch_losses(const float *losses3, const float *regsq, int B,
                         const oracle_hparams *hp, float *losses) {
  double rs = 0.0;
  for (int b = 0; b < B; ++b) rs += regsq[b];
  /* regularizer = (l2(u)+l2(p)+l2(n))/batch_size; l2_loss = sum(x^2)/2  (model.py:219-221) */
  const float reg = hp->decay * ((float)(rs * 0.5) / (float)hp->batch_size_flag);
  const float mf = losses3[0] + hp->alpha * losses3[1] + hp->beta * losses3[2]; /* :217 */
  losses[0] = mf + reg; /* model.py:73 */
  losses[1] = mf;
  losses[2] = reg;
  losses[3] = losses3[0];
}

static void mf_step_impl(float *U, float *mU, float *vU, int64_t n_users, float *I, float *mI,
                         float *vI, int64_t n_items, float *w, float *mw, float *vw, float *wu,
                         float *mwu, float *vwu, int d, const int32_t *u, const int32_t *p,
                         const int32_t *n, int B, const oracle_hparams *hp, float *pw,
                         float *losses, int item_only) {
  float *sc = (float *)malloc(sizeof(float) * (size_t)B * 11);
  float *yp = sc, *yn = sc + B, *sp = sc + 2 * B, *sn = sc + 3 * B, *su = sc + 4 * B,
        *rq = sc + 5 * B, *dyp = sc + 6 * B, *dyn = sc + 7 * B, *dsp = sc + 8 * B,
        *dsn = sc + 9 * B, *dsu = sc + 10 * B;
  float l3[3];
  oracle_gather_dots(U, I, U, I, w, wu, u, p, n, B, d, yp, yn, sp, sn, su, rq);
  grid_bce_impl(yp, yn, sp, sn, su, B, hp->alpha, hp->beta, l3, dyp, dyn, dsp, dsn, dsu,
                item_only);
  batch_losses(l3, rq, B, hp, losses); /* item_only: l3[2] = 0 -> mf = L_ori + alpha*L_item */

  float *gU = (float *)malloc(sizeof(float) * (size_t)B * d * 3);
  float *gPN = gU + (size_t)B * d;
  float gw[256], gwu[256];
  const float lam = hp->decay / (float)hp->batch_size_flag;
  row_grads(U, I, U, I, w, wu, u, p, n, B, d, dyp, dyn, dsp, dsn, dsu, lam, 1, gU, gPN, gw,
            gwu);
  int32_t *pn = (int32_t *)malloc(sizeof(int32_t) * (size_t)B * 2);
  memcpy(pn, p, sizeof(int32_t) * B);
  memcpy(pn + B, n, sizeof(int32_t) * B); /* IndexedSlices of the two lookups, concatenated */

  const float lr_t = oracle_adam_lr_t(hp->lr, pw[0], pw[1]);
  oracle_adam_sparse(U, mU, vU, n_users, d, u, gU, B, lr_t, hp->beta1, hp->beta2, hp->eps);
  oracle_adam_sparse(I, mI, vI, n_items, d, pn, gPN, 2 * B, lr_t, hp->beta1, hp->beta2,
                     hp->eps);
  oracle_adam_dense_vec(w, mw, vw, gw, d, lr_t, hp->beta1, hp->beta2, hp->eps);
  if (!item_only) /* no gradient reaches user_branch: minimize() leaves it and its slots alone */
    oracle_adam_dense_vec(wu, mwu, vwu, gwu, d, lr_t, hp->beta1, hp->beta2, hp->eps);
  pw[0] = pw[0] * hp->beta1; /* adam.py _finish */
  pw[1] = pw[1] * hp->beta2;
  free(sc);
  free(gU);
  free(pn);
}

void oracle_mf_step(float *U, float *mU, float *vU, int64_t n_users, float *I, float *mI,
                    float *vI, int64_t n_items, float *w, float *mw, float *vw, float *wu,
                    float *mwu, float *vwu, int d, const int32_t *u, const int32_t *p,
                    const int32_t *n, int B, const oracle_hparams *hp, float *pw,
                    float *losses) {
  mf_step_impl(U, mU, vU, n_users, I, mI, vI, n_items, w, mw, vw, wu, mwu, vwu, d, u, p, n, B, hp,
               pw, losses, 0);
}

/* `--train rubibce` (model.py:83-85 opt_two_bce, :158-183): wu is read by nothing that matters
 * (the user logits are computed and dropped) and is never written */
void oracle_mf_step_item(float *U, float *mU, float *vU, int64_t n_users, float *I, float *mI,
                         float *vI, int64_t n_items, float *w, float *mw, float *vw, float *wu,
                         int d, const int32_t *u, const int32_t *p, const int32_t *n, int B,
                         const oracle_hparams *hp, float *pw, float *losses) {
  mf_step_impl(U, mU, vU, n_users, I, mI, vI, n_items, w, mw, vw, wu, NULL, NULL, d, u, p, n, B, hp,
               pw, losses, 1);
}

/* ---- `--train normalbce` (the README's baseline command, README.md:30): model.py:277-287 ----
 * mf_loss = mean_b( -log(sig(yp_b) + 1e-9) - log(1 - sig(yn_b) + 1e-9) ), reg as above;
 * model.py:100 `self.opt = AdamOptimizer(lr).minimize(self.loss)`: only the two embedding tables
 * receive gradients (w, w_user are not in this graph and stay untouched). */
void oracle_plain_bce(const float *yp, const float *yn, int B, float *mf_loss, float *dyp,
                      float *dyn) {
  const float eps9 = 1e-9f;
  double acc = 0.0;
  for (int b = 0; b < B; ++b) {
    const float s = sigmoidf_(yp[b]), t = sigmoidf_(yn[b]);
    const float se = s + eps9, q = (1.0f - t) + eps9;
    acc += (double)(-logf(se)) + (double)(-logf(q));
    if (dyp) {
      dyp[b] = -(s * (1.0f - s)) / se / (float)B;
      dyn[b] = (t * (1.0f - t)) / q / (float)B;
    }
  }
  *mf_loss = (float)(acc / (double)B);
}

void oracle_mf_step_normal(float *U, float *mU, float *vU, int64_t n_users, float *I, float *mI,
                           float *vI, int64_t n_items, int d, const int32_t *u, const int32_t *p,
                           const int32_t *n, int B, const oracle_hparams *hp, float *pw,
                           float *losses) {
  float *sc = (float *)malloc(sizeof(float) * (size_t)B * 11);
  float *yp = sc, *yn = sc + B, *sp = sc + 2 * B, *sn = sc + 3 * B, *su = sc + 4 * B,
        *rq = sc + 5 * B, *dyp = sc + 6 * B, *dyn = sc + 7 * B, *dz = sc + 8 * B;
  float *zero_w = (float *)calloc((size_t)d, sizeof(float));
  oracle_gather_dots(U, I, U, I, zero_w, zero_w, u, p, n, B, d, yp, yn, sp, sn, su, rq);
  float mf;
  oracle_plain_bce(yp, yn, B, &mf, dyp, dyn);
  double rs = 0.0;
  for (int b = 0; b < B; ++b) rs += rq[b];
  const float reg = hp->decay * ((float)(rs * 0.5) / (float)hp->batch_size_flag);
  losses[0] = mf + reg;
  losses[1] = mf;
  losses[2] = reg;
  losses[3] = mf;
  memset(dz, 0, sizeof(float) * (size_t)B);
  float *gU = (float *)malloc(sizeof(float) * (size_t)B * d * 3);
  float *gPN = gU + (size_t)B * d;
  float gw[256], gwu[256];
  const float lam = hp->decay / (float)hp->batch_size_flag;
  row_grads(U, I, U, I, zero_w, zero_w, u, p, n, B, d, dyp, dyn, dz, dz, dz, lam, 1, gU, gPN, gw,
            gwu);
  int32_t *pn = (int32_t *)malloc(sizeof(int32_t) * (size_t)B * 2);
  memcpy(pn, p, sizeof(int32_t) * B);
  memcpy(pn + B, n, sizeof(int32_t) * B);
  const float lr_t = oracle_adam_lr_t(hp->lr, pw[0], pw[1]);
  oracle_adam_sparse(U, mU, vU, n_users, d, u, gU, B, lr_t, hp->beta1, hp->beta2, hp->eps);
  oracle_adam_sparse(I, mI, vI, n_items, d, pn, gPN, 2 * B, lr_t, hp->beta1, hp->beta2,
                     hp->eps);
  pw[0] = pw[0] * hp->beta1;
  pw[1] = pw[1] * hp->beta2;
  free(sc);
  free(zero_w);
  free(gU);
  free(pn);
}

/* ---- LightGCN.py:297-305: side = A_hat @ ego (all 100 row folds concatenated) ---- */
void oracle_spmm_csr(const int32_t *rowptr, const int32_t *col, const float *val,
                     int64_t n_rows, const float *X, int d, float *Y) {
#pragma omp parallel for schedule(dynamic, 256)
  for (int64_t r = 0; r < n_rows; ++r) {
    double acc[256];
    for (int k = 0; k < d; ++k) acc[k] = 0.0;
    for (int32_t e = rowptr[r]; e < rowptr[r + 1]; ++e) {
      const float *x = X + (int64_t)col[e] * d;
      const double a = val[e];
      for (int k = 0; k < d; ++k) acc[k] += a * x[k];
    }
    for (int k = 0; k < d; ++k) Y[r * d + k] = (float)acc[k];
  }
}

void oracle_spmm_csr_t(const int32_t *rowptr, const int32_t *col, const float *val,
                       int64_t n_rows, const float *X, int d, float *Y) {
  double *acc = (double *)calloc((size_t)n_rows * d, sizeof(double));
  for (int64_t r = 0; r < n_rows; ++r)
    for (int32_t e = rowptr[r]; e < rowptr[r + 1]; ++e) {
      double *y = acc + (int64_t)col[e] * d;
      const float *x = X + r * d;
      for (int k = 0; k < d; ++k) y[k] += (double)val[e] * x[k];
    }
  for (int64_t e = 0; e < n_rows * d; ++e) Y[e] = (float)acc[e];
  free(acc);
}

/* ---- LightGCN.py:288-309 ---- */
static void lgcn_layers(const int32_t *rowptr, const int32_t *col, const float *val,
                        const float *U, int64_t n_users, const float *I, int64_t n_items, int d,
                        int L, float *Emean, float *E0) {
  const int64_t N = n_users + n_items, ne = N * d;
  memcpy(E0, U, sizeof(float) * (size_t)n_users * d); /* tf.concat([user, item], 0) */
  memcpy(E0 + n_users * d, I, sizeof(float) * (size_t)n_items * d);
  float *cur = (float *)malloc(sizeof(float) * (size_t)ne * 2), *nxt = cur + ne;
  memcpy(cur, E0, sizeof(float) * (size_t)ne);
  memcpy(Emean, E0, sizeof(float) * (size_t)ne);
  for (int k = 0; k < L; ++k) {
    oracle_spmm_csr(rowptr, col, val, N, cur, d, nxt);
    for (int64_t e = 0; e < ne; ++e) Emean[e] = Emean[e] + nxt[e]; /* stack + reduce_mean */
    float *t = cur;
    cur = nxt;
    nxt = t;
  }
  const float cnt = (float)(L + 1);
  for (int64_t e = 0; e < ne; ++e) Emean[e] = Emean[e] / cnt;
  free(cur < nxt ? cur : nxt);
}

void oracle_lgcn_propagate(const int32_t *rowptr, const int32_t *col, const float *val,
                           const float *U, int64_t n_users, const float *I, int64_t n_items,
                           int d, int n_layers, float *Emean) {
  float *E0 = (float *)malloc(sizeof(float) * (size_t)(n_users + n_items) * d);
  lgcn_layers(rowptr, col, val, U, n_users, I, n_items, d, n_layers, Emean, E0);
  free(E0);
}

/* normal != 0: `--loss bce` (LightGCN.py:415-429,:186): element-wise BCE with "+1e-9" on the
 * propagated rows, same L2 on the raw rows, w / w_user not part of the graph */
static void lgcn_step_impl(const int32_t *rowptr, const int32_t *col, const float *val, float *U,
                           float *mU, float *vU, int64_t n_users, float *I, float *mI, float *vI,
                           int64_t n_items, float *w, float *mw, float *vw, float *wu, float *mwu,
                           float *vwu, int d, int L, const int32_t *u, const int32_t *p,
                           const int32_t *n, int B, int train, const oracle_hparams *hp, float *pw,
                           float *losses, int mode) {
  const int normal = mode == 1, item_only = mode == 2; /* 2: `--loss bce1`, LightGCN.py:431-461 */
  const int64_t N = n_users + n_items, ne = N * d;
  float *Em = (float *)malloc(sizeof(float) * (size_t)ne * 2), *E0 = Em + ne;
  lgcn_layers(rowptr, col, val, U, n_users, I, n_items, d, L, Em, E0);
  const float *Ue = Em, *Ie = Em + n_users * d;

  float *sc = (float *)malloc(sizeof(float) * (size_t)B * 11);
  float *yp = sc, *yn = sc + B, *sp = sc + 2 * B, *sn = sc + 3 * B, *su = sc + 4 * B,
        *rq = sc + 5 * B, *dyp = sc + 6 * B, *dyn = sc + 7 * B, *dsp = sc + 8 * B,
        *dsn = sc + 9 * B, *dsu = sc + 10 * B;
  float l3[3];
  /* scores from propagated rows (LightGCN.py:145-147), L2 from raw rows (:148-150,525-526) */
  oracle_gather_dots(Ue, Ie, U, I, w, wu, u, p, n, B, d, yp, yn, sp, sn, su, rq);
  if (normal) {
    float mf;
    oracle_plain_bce(yp, yn, B, &mf, dyp, dyn);
    memset(dsp, 0, sizeof(float) * (size_t)B * 3); /* dsp, dsn, dsu are contiguous */
    double rs = 0.0;
    for (int b = 0; b < B; ++b) rs += rq[b];
    const float emb = hp->decay * ((float)(rs * 0.5) / (float)hp->batch_size_flag);
    losses[0] = mf + emb; /* LightGCN.py:185 loss_bce = mf_loss_bce + emb_loss_bce */
    losses[1] = mf;
    losses[2] = emb;
    losses[3] = mf;
  } else {
    grid_bce_impl(yp, yn, sp, sn, su, B, hp->alpha, hp->beta, l3, train ? dyp : NULL, dyn, dsp,
                  dsn, dsu, item_only);
    batch_losses(l3, rq, B, hp, losses); /* loss = mf_loss + emb_loss (LightGCN.py:200) */
  }
  if (!train) {
    free(Em);
    free(sc);
    return;
  }

  float *gU = (float *)malloc(sizeof(float) * (size_t)B * d * 3);
  float *gPN = gU + (size_t)B * d;
  float gw[256], gwu[256];
  row_grads(Ue, Ie, U, I, w, wu, u, p, n, B, d, dyp, dyn, dsp, dsn, dsu, 0.f, 0, gU, gPN, gw,
            gwu);
  /* d(Emean): scatter-add of the three lookups' slices */
  float *dEm = (float *)calloc((size_t)ne * 3, sizeof(float));
  float *acc = dEm + ne, *tmp = dEm + 2 * ne;
  for (int b = 0; b < B; ++b) {
    float *du = dEm + (int64_t)u[b] * d, *dp = dEm + (n_users + p[b]) * d,
          *dn = dEm + (n_users + n[b]) * d;
    for (int k = 0; k < d; ++k) {
      du[k] += gU[(int64_t)b * d + k];
      dp[k] += gPN[(int64_t)b * d + k];
      dn[k] += gPN[(int64_t)(B + b) * d + k];
    }
  }
  /* reduce_mean backward: every stacked layer receives dEm/(L+1); layer k+1 = A layer k */
  const float cnt = (float)(L + 1);
  for (int64_t e = 0; e < ne; ++e) dEm[e] = dEm[e] / cnt;
  memcpy(acc, dEm, sizeof(float) * (size_t)ne); /* grad of the last layer */
  for (int k = 0; k < L; ++k) {
    oracle_spmm_csr_t(rowptr, col, val, N, acc, d, tmp);
    for (int64_t e = 0; e < ne; ++e) acc[e] = dEm[e] + tmp[e];
  }
  /* + L2 slices on the raw rows; all rows are "indices" of the aggregated IndexedSlices */
  const float lam = hp->decay / (float)hp->batch_size_flag;
  for (int b = 0; b < B; ++b) {
    float *du = acc + (int64_t)u[b] * d;
    for (int k = 0; k < d; ++k) du[k] = du[k] + lam * U[(int64_t)u[b] * d + k];
  }
  for (int b = 0; b < B; ++b) {
    float *dp = acc + (n_users + p[b]) * d;
    for (int k = 0; k < d; ++k) dp[k] = dp[k] + lam * I[(int64_t)p[b] * d + k];
  }
  for (int b = 0; b < B; ++b) {
    float *dn = acc + (n_users + n[b]) * d;
    for (int k = 0; k < d; ++k) dn[k] = dn[k] + lam * I[(int64_t)n[b] * d + k];
  }
  const float lr_t = oracle_adam_lr_t(hp->lr, pw[0], pw[1]);
  const float omb1 = 1.0f - hp->beta1, omb2 = 1.0f - hp->beta2;
  for (int t = 0; t < 2; ++t) {
    float *var = t ? I : U, *m = t ? mI : mU, *v = t ? vI : vU;
    const float *g = t ? acc + n_users * d : acc;
    const int64_t cntE = (t ? n_items : n_users) * d;
    for (int64_t e = 0; e < cntE; ++e) {
      m[e] = m[e] * hp->beta1 + g[e] * omb1;
      v[e] = v[e] * hp->beta2 + (g[e] * g[e]) * omb2;
      var[e] = var[e] - (lr_t * m[e]) / (sqrtf(v[e]) + hp->eps);
    }
  }
  if (!normal) {
    oracle_adam_dense_vec(w, mw, vw, gw, d, lr_t, hp->beta1, hp->beta2, hp->eps);
    if (!item_only)
      oracle_adam_dense_vec(wu, mwu, vwu, gwu, d, lr_t, hp->beta1, hp->beta2, hp->eps);
  }
  pw[0] = pw[0] * hp->beta1;
  pw[1] = pw[1] * hp->beta2;
  free(Em);
  free(sc);
  free(gU);
  free(dEm);
}

void oracle_lgcn_step(const int32_t *rowptr, const int32_t *col, const float *val, float *U,
                      float *mU, float *vU, int64_t n_users, float *I, float *mI, float *vI,
                      int64_t n_items, float *w, float *mw, float *vw, float *wu, float *mwu,
                      float *vwu, int d, int L, const int32_t *u, const int32_t *p,
                      const int32_t *n, int B, int train, const oracle_hparams *hp, float *pw,
                      float *losses) {
  lgcn_step_impl(rowptr, col, val, U, mU, vU, n_users, I, mI, vI, n_items, w, mw, vw, wu, mwu, vwu,
                 d, L, u, p, n, B, train, hp, pw, losses, 0);
}

void oracle_lgcn_step_normal(const int32_t *rowptr, const int32_t *col, const float *val, float *U,
                             float *mU, float *vU, int64_t n_users, float *I, float *mI,
                             float *vI, int64_t n_items, float *w, float *wu, int d, int L,
                             const int32_t *u, const int32_t *p, const int32_t *n, int B, int train,
                             const oracle_hparams *hp, float *pw, float *losses) {
  lgcn_step_impl(rowptr, col, val, U, mU, vU, n_users, I, mI, vI, n_items, w, NULL, NULL, wu, NULL,
                 NULL, d, L, u, p, n, B, train, hp, pw, losses, 1);
}

void oracle_lgcn_step_item(const int32_t *rowptr, const int32_t *col, const float *val, float *U,
                           float *mU, float *vU, int64_t n_users, float *I, float *mI, float *vI,
                           int64_t n_items, float *w, float *mw, float *vw, float *wu, int d,
                           int L, const int32_t *u, const int32_t *p, const int32_t *n, int B,
                           int train, const oracle_hparams *hp, float *pw, float *losses) {
  lgcn_step_impl(rowptr, col, val, U, mU, vU, n_users, I, mI, vI, n_items, w, mw, vw, wu, NULL,
                 NULL, d, L, u, p, n, B, train, hp, pw, losses, 2);
}

/* ---- scoring: model.py:45 batch_ratings, :199 rubi_ratings_both ---- */
static inline float dot_fma_chain(const float *a, const float *b, int d) {
  float acc = 0.0f;
  for (int k = 0; k < d; ++k) acc = fmaf(a[k], b[k], acc);
  return acc;
}

void oracle_score_gates(const float *rows, int64_t n, int d, const float *wvec, float *sig) {
#pragma omp parallel for schedule(static)
  for (int64_t r = 0; r < n; ++r) {
    const float x = dot_fma_chain(rows + r * d, wvec, d);
    sig[r] = (float)(1.0 / (1.0 + exp(-(double)x)));
  }
}

static inline float score_one(const float *uq, const float *it, int d, float sig_i, float sig_u,
                              float c) {
  const float y = dot_fma_chain(uq, it, d);
  return ((y - c) * sig_i) * sig_u; /* (batch_ratings-rubi_c)*sig(item)^T*sig(user) */
}

void oracle_score_matrix(const float *Uq, int T, const float *It, int64_t n_items, int d,
                         const float *sig_i, const float *sig_u, float c, float *out) {
#pragma omp parallel for schedule(static)
  for (int t = 0; t < T; ++t)
    for (int64_t i = 0; i < n_items; ++i)
      out[(int64_t)t * n_items + i] =
          score_one(Uq + (int64_t)t * d, It + i * d, d, sig_i[i], sig_u[t], c);
}

/* strict total order: higher score first, then lower id */
static inline int better(float sa, int32_t ia, float sb, int32_t ib) {
  return sa > sb || (sa == sb && ia < ib);
}

static void topk_insert(int32_t *ids, float *sc, int K, int *cnt, float s, int32_t id) {
  if (*cnt == K && !better(s, id, sc[K - 1], ids[K - 1])) return;
  int pos = *cnt < K ? (*cnt)++ : K - 1;
  while (pos > 0 && better(s, id, sc[pos - 1], ids[pos - 1])) {
    sc[pos] = sc[pos - 1];
    ids[pos] = ids[pos - 1];
    --pos;
  }
  sc[pos] = s;
  ids[pos] = id;
}

void oracle_score_topk(const float *Uq, int T, const float *It, int64_t n_items, int d,
                       const float *sig_i, const float *sig_u, float c,
                       const int32_t *mask_rowptr, const int32_t *mask_col, int K,
                       int32_t off, int32_t *out_ids, float *out_scores) {
#pragma omp parallel for schedule(dynamic, 16)
  for (int t = 0; t < T; ++t) {
    int32_t *ids = out_ids + (int64_t)t * K;
    float *sc = out_scores + (int64_t)t * K;
    int cnt = 0;
    int32_t mp = mask_rowptr ? mask_rowptr[t] : 0, me = mask_rowptr ? mask_rowptr[t + 1] : 0;
    while (mp < me && mask_col[mp] < off) ++mp;
    for (int64_t i = 0; i < n_items; ++i) {
      const int32_t gid = (int32_t)(off + i);
      while (mp < me && mask_col[mp] < gid) ++mp;
      if (mp < me && mask_col[mp] == gid) continue; /* train item: train.py:133 / batch_test.py:129 */
      const float s = score_one(Uq + (int64_t)t * d, It + i * d, d, sig_i[i], sig_u[t], c);
      topk_insert(ids, sc, K, &cnt, s, gid);
    }
    for (int k = cnt; k < K; ++k) {
      ids[k] = -1;
      sc[k] = -INFINITY;
    }
  }
}

void oracle_topk_rows(const float *scores, int columns_num, int rows_num, int top_k,
                      int32_t *rankings) {
#pragma omp parallel for schedule(dynamic, 16)
  for (int r = 0; r < rows_num; ++r) {
    float sc[256];
    int cnt = 0;
    int32_t *ids = rankings + (int64_t)r * top_k;
    const float *row = scores + (int64_t)r * columns_num;
    for (int i = 0; i < columns_num; ++i) topk_insert(ids, sc, top_k, &cnt, row[i], i);
    for (int k = cnt; k < top_k; ++k) ids[k] = -1;
  }
}

void oracle_topk_merge(const int32_t *ids, const float *scores, int T, int K, int G,
                       int32_t *out_ids, float *out_scores) {
  for (int t = 0; t < T; ++t) {
    int cnt = 0;
    int32_t *oi = out_ids + (int64_t)t * K;
    float *os = out_scores + (int64_t)t * K;
    for (int g = 0; g < G; ++g)
      for (int k = 0; k < K; ++k) {
        const int64_t e = ((int64_t)g * T + t) * K + k;
        if (ids[e] < 0) continue;
        topk_insert(oi, os, K, &cnt, scores[e], ids[e]);
      }
    for (int k = cnt; k < K; ++k) {
      oi[k] = -1;
      os[k] = -INFINITY;
    }
  }
}

/* ---- evaluate_foldout.h:16-113: precision / recall / ap / ndcg / mrr curves @1..K ---- */
static int in_truth(const int32_t *truth, int n, int32_t x) {
  for (int i = 0; i < n; ++i)
    if (truth[i] == x) return 1;
  return 0;
}

void oracle_inv_log2_table(int K, double *out) {
  for (int i = 0; i < K; ++i) out[i] = 1.0 / log2((double)(i + 2));
}

void oracle_foldout_metrics(const int32_t *topk_ids, int T, int K, const int32_t *truth_rowptr,
                            const int32_t *truth_col, float *out) {
#pragma omp parallel for schedule(static)
  for (int t = 0; t < T; ++t) {
    const int32_t *rank = topk_ids + (int64_t)t * K;
    const int32_t *truth = truth_col + truth_rowptr[t];
    const int tl = truth_rowptr[t + 1] - truth_rowptr[t];
    float *o = out + (int64_t)t * 5 * K;
    int hits = 0;
    float sum_pre = 0.f, DCG = 0.f, iDCG = 0.f, rr = 0.f;
    int found = 0;
    for (int i = 0; i < K; ++i) {
      const int hit = in_truth(truth, tl, rank[i]);
      if (hit) {
        hits += 1;
        const float pre = (float)(1.0 * hits / (i + 1)); /* ap(): float pre = 1.0*hits/(i+1) */
        sum_pre += pre;
        DCG = (float)((double)DCG + 1.0 / log2((double)(i + 2)));
        if (!found) {
          found = 1;
          rr = (float)(1.0 / (i + 1));
        }
      }
      if (i < tl) iDCG = (float)((double)iDCG + 1.0 / log2((double)(i + 2)));
      o[0 * K + i] = (float)(1.0 * hits / (i + 1)); /* precision() */
      o[1 * K + i] = (float)(1.0 * hits / tl);      /* recall()    */
      o[2 * K + i] = sum_pre / (float)tl;           /* ap()        */
      o[3 * K + i] = DCG / iDCG;                    /* ndcg()      */
      o[4 * K + i] = rr;                            /* mrr()       */
    }
  }
}
