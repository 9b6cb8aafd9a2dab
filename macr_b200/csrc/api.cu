// api.cu -- error plumbing and the kernel-level C-ABI entry points of the training path
// (declared in include/macr_b200.h).  Step handles live in trainer.cu.
#include <stdarg.h>

#include "train_kernels.cuh"

namespace macr {

char *err_buf() {
  static thread_local char buf[512] = {0};
  return buf;
}

int fail(int code, const char *fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(err_buf(), 512, fmt, ap);
  va_end(ap);
  return code;
}

bool pdl_enabled() {
  static int on = -1;
  if (on < 0) {
    const char *e = getenv("MACR_PDL");
    on = (e && e[0] == '1') ? 1 : 0;
  }
  return on == 1;
}

int sm_count() {
  static int sms = 0;
  if (sms == 0) {
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess ||
        cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || sms <= 0)
      sms = 148;  // B200
  }
  return sms;
}

}  // namespace macr

using namespace macr;

extern "C" const char *macr_last_error(void) { return err_buf(); }
extern "C" int macr_abi_version(void) { return 1; }
extern "C" int macr_device_sm_count(int *out) {
  MACR_CHECK_ARG(out != nullptr, "macr_device_sm_count: null out");
  int dev = 0;
  MACR_CUDA(cudaGetDevice(&dev));
  MACR_CUDA(cudaDeviceGetAttribute(out, cudaDevAttrMultiProcessorCount, dev));
  return MACR_OK;
}

extern "C" int macr_gather_dots(const float *Ue, const float *Ie, const float *Ur, const float *Ir,
                                const float *w, const float *w_user, const int32_t *users,
                                const int32_t *pos, const int32_t *neg, int B, int d, float *yp,
                                float *yn, float *sp, float *sn, float *su, float *regsq,
                                macr_stream_t stream) {
  MACR_CHECK_ARG(d == kD, "macr_gather_dots: d must be %d (got %d)", kD, d);
  MACR_CHECK_ARG(B >= 0, "macr_gather_dots: negative batch");
  if (B == 0) return MACR_OK;
  MACR_CHECK_ARG(Ue && Ie && Ur && Ir && w && w_user && users && pos && neg && yp && yn && sp &&
                     sn && su && regsq,
                 "macr_gather_dots: null pointer");
  return launch_gather_dots(Ue, Ie, Ur, Ir, w, w_user, users, pos, neg, nullptr, B, yp, yn, sp, sn,
                            su, regsq, nullptr, nullptr, as_stream(stream));
}

extern "C" size_t macr_grid_bce_workspace_bytes(int B) {
  if (B <= 0) return 16;
  return grid_ws_layout(B, nullptr).bytes;
}

extern "C" int macr_grid_bce_fwd_bwd(const float *yp, const float *yn, const float *sp,
                                     const float *sn, const float *su, int B, float alpha,
                                     float beta, float *losses3, float *d_yp, float *d_yn,
                                     float *d_sp, float *d_sn, float *d_su, void *ws,
                                     size_t ws_bytes, macr_stream_t stream) {
  MACR_CHECK_ARG(B > 0, "macr_grid_bce_fwd_bwd: batch must be positive (got %d)", B);
  MACR_CHECK_ARG(yp && yn && sp && sn && su && losses3 && ws, "macr_grid_bce_fwd_bwd: null pointer");
  const int want_grad = d_yp != nullptr;
  if (want_grad)
    MACR_CHECK_ARG(d_yn && d_sp && d_sn && d_su, "macr_grid_bce_fwd_bwd: partial gradient outputs");
  GridWs g = grid_ws_layout(B, ws);
  if (ws_bytes < g.bytes)
    return fail(MACR_ERR_WORKSPACE, "macr_grid_bce_fwd_bwd: workspace %zu < %zu bytes", ws_bytes,
                g.bytes);
  cudaStream_t s = as_stream(stream);
  // partial-sum slots start empty (the folders re-arm them, but `ws` is caller scratch)
  MACR_CUDA(cudaMemsetAsync(ws, 0xff, g.part_bytes, s));
  int rc = launch_gates(sp, sn, su, B, g, s);
  if (rc) return rc;
  macr_hparams hp{};
  hp.alpha = alpha;
  hp.beta = beta;
  hp.batch_size_flag = 1;
  return launch_grid_bce(yp, yn, B, hp, g, d_yp, d_yn, d_sp, d_sn, d_su, want_grad, nullptr, nullptr,
                         losses3, s);
}

extern "C" size_t macr_batch_plan_workspace_bytes(int n_ids) { return plan_ws_bytes(n_ids); }

extern "C" int macr_batch_plan(const int32_t *ids, int n_ids, int64_t table_rows,
                               int32_t *uniq_rows, int32_t *seg_off, int32_t *seg_pos,
                               int32_t *n_uniq, uint32_t *touched_bitmap, void *ws,
                               size_t ws_bytes, macr_stream_t stream) {
  MACR_CHECK_ARG(n_ids > 0 && n_ids <= 16384, "macr_batch_plan: n_ids must be in [1,16384] (got %d)",
                 n_ids);
  MACR_CHECK_ARG(ids && uniq_rows && seg_off && seg_pos && n_uniq && ws,
                 "macr_batch_plan: null pointer");
  if (ws_bytes < plan_ws_bytes(n_ids))
    return fail(MACR_ERR_WORKSPACE, "macr_batch_plan: workspace %zu < %zu bytes", ws_bytes,
                plan_ws_bytes(n_ids));
  cudaStream_t s = as_stream(stream);
  PlanBufs out = plan_carve(uniq_rows, seg_off, seg_pos, n_uniq, ws, n_ids);
  PlanBufs none{};
  MACR_CHECK_ARG(table_rows > 0, "macr_batch_plan: table_rows must be positive");
  return launch_batch_plan2(ids, nullptr, 0, n_ids, table_rows, out, touched_bitmap, nullptr, 0, 0, 1,
                            none, nullptr, s);
}

extern "C" int macr_adam_sweep_untouched(float *var, float *m, float *v, int64_t rows, int d,
                                         const uint32_t *touched_bitmap, float lr_t, float beta1,
                                         float beta2, float eps, macr_stream_t stream) {
  MACR_CHECK_ARG(d == kD, "macr_adam_sweep_untouched: d must be %d (got %d)", kD, d);
  MACR_CHECK_ARG(rows >= 0 && var && m && v, "macr_adam_sweep_untouched: bad arguments");
  return launch_adam_sweep2(var, m, v, rows, touched_bitmap, nullptr, nullptr, nullptr, 0, nullptr,
                            lr_t, nullptr, beta1, beta2, eps, as_stream(stream));
}

extern "C" int macr_adam_rows(float *var, float *m, float *v, int64_t rows, int d,
                              const int32_t *uniq_rows, const float *grad_rows, int n_uniq,
                              uint32_t *touched_bitmap, float lr_t, float beta1, float beta2,
                              float eps, macr_stream_t stream) {
  MACR_CHECK_ARG(d == kD, "macr_adam_rows: d must be %d (got %d)", kD, d);
  MACR_CHECK_ARG(rows >= 0 && n_uniq >= 0, "macr_adam_rows: negative size");
  if (n_uniq == 0) return MACR_OK;
  MACR_CHECK_ARG(var && m && v && uniq_rows && grad_rows, "macr_adam_rows: null pointer");
  cudaStream_t s = as_stream(stream);
  int32_t *cnt = nullptr;
  MACR_CUDA(cudaMallocAsync(&cnt, sizeof(int32_t) * 2, s));
  MACR_CUDA(cudaMemcpyAsync(cnt, &n_uniq, sizeof(int32_t), cudaMemcpyHostToDevice, s));
  MACR_CUDA(cudaMemsetAsync(cnt + 1, 0, sizeof(int32_t), s));
  PlanBufs pu{};
  pu.uniq_rows = const_cast<int32_t *>(uniq_rows);
  pu.n_uniq = cnt;
  PlanBufs pi{};
  pi.n_uniq = cnt + 1;
  // n_uniq user-slot warps, no item-slot work
  int rc = launch_adam_rows2(var, m, v, pu, grad_rows, touched_bitmap, nullptr, nullptr, nullptr,
                             pi, nullptr, nullptr, n_uniq, lr_t, nullptr, beta1, beta2, eps, s);
  MACR_CUDA(cudaFreeAsync(cnt, s));
  return rc;
}

extern "C" int macr_adam_dense(float *var, float *m, float *v, const float *grad, int64_t rows,
                               int d, float lr_t, float beta1, float beta2, float eps,
                               macr_stream_t stream) {
  MACR_CHECK_ARG(d == kD, "macr_adam_dense: d must be %d (got %d)", kD, d);
  MACR_CHECK_ARG(rows >= 0 && var && m && v && grad, "macr_adam_dense: bad arguments");
  return launch_adam_dense(var, m, v, grad, rows * d, lr_t, nullptr, beta1, beta2, eps,
                           as_stream(stream));
}

extern "C" int macr_adam_vec(float *var, float *m, float *v, const float *grad, int n, float lr_t,
                             float beta1, float beta2, float eps, macr_stream_t stream) {
  MACR_CHECK_ARG(n == kD, "macr_adam_vec: n must be %d (got %d)", kD, n);
  MACR_CHECK_ARG(var && m && v && grad, "macr_adam_vec: null pointer");
  // one "partial" = the gradient itself; second vector aliased to the first with zero parts is
  // not expressible, so run the pair kernel on (var, var) halves: w <- grad, w_user <- unused
  cudaStream_t s = as_stream(stream);
  float *scratch = nullptr;
  MACR_CUDA(cudaMallocAsync(&scratch, sizeof(float) * kD * 4, s));
  MACR_CUDA(cudaMemsetAsync(scratch, 0, sizeof(float) * kD * 4, s));
  int rc = launch_adam_vec2(var, m, v, scratch, scratch + kD, scratch + 2 * kD, grad,
                            scratch + 3 * kD, 1, lr_t, nullptr, beta1, beta2, eps, s);
  MACR_CUDA(cudaFreeAsync(scratch, s));
  return rc;
}
