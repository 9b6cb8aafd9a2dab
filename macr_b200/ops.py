"""Tensor-level wrappers over the C ABI (include/macr_b200.h).

torch is used for device memory and streams only; every function below hands raw device
pointers to libmacr_b200.so and raises ``MacrError`` on any failure -- there is no eager /
CPU fallback path.
"""
import ctypes as C
import math

import numpy as np
import torch

from . import _lib
from ._lib import HParams, MacrError, ShardDesc, check, lib, ptr, stream_ptr

D = 64
MAX_TOPK = 32


def _cuda(t, dtype):
    if not (isinstance(t, torch.Tensor) and t.is_cuda and t.dtype == dtype and t.is_contiguous()):
        raise MacrError(f"expected a contiguous CUDA {dtype} tensor, got {type(t)} "
                        f"{getattr(t, 'dtype', None)} {getattr(t, 'device', None)}")
    return t


def _f(t):
    return ptr(_cuda(t, torch.float32))


def _i(t):
    return ptr(_cuda(t, torch.int32))


def sm_count():
    out = C.c_int(0)
    check(lib().macr_device_sm_count(C.byref(out)), "macr_device_sm_count")
    return out.value


def gather_dots(Ue, Ie, Ur, Ir, w, wu, users, pos, neg):
    """-> (yp, yn, sp, sn, su, regsq), each [B] fp32 (macr_mf/model.py:186-187,194-196,219)."""
    B = users.numel()
    out = torch.empty((6, max(B, 1)), dtype=torch.float32, device=Ue.device)
    check(lib().macr_gather_dots(_f(Ue), _f(Ie), _f(Ur), _f(Ir), _f(w), _f(wu), _i(users), _i(pos),
                                 _i(neg), B, Ue.shape[1], *[ptr(out[k]) for k in range(6)],
                                 stream_ptr()), "macr_gather_dots")
    return tuple(out[k, :B] for k in range(6))


def grid_bce(yp, yn, sp, sn, su, alpha, beta, want_grad=True, bufs=None):
    """-> (losses3 [L_ori, L_item, L_user], (d_yp, d_yn, d_sp, d_sn, d_su) or None).
    bufs: optional (ws uint8, losses f32[3], grads f32[5,B]) from an earlier call, reused as they are."""
    B = yp.numel()
    dev = yp.device
    nbytes = lib().macr_grid_bce_workspace_bytes(B)
    if bufs is not None:
        ws, losses, grads = bufs
    else:
        ws = torch.empty(nbytes, dtype=torch.uint8, device=dev)
        losses = torch.empty(3, dtype=torch.float32, device=dev)
        grads = torch.empty((5, B), dtype=torch.float32, device=dev) if want_grad else None
    gp = [ptr(grads[k]) for k in range(5)] if want_grad else [None] * 5
    check(lib().macr_grid_bce_fwd_bwd(_f(yp), _f(yn), _f(sp), _f(sn), _f(su), B, alpha, beta,
                                      ptr(losses), *gp, ptr(ws), nbytes, stream_ptr()),
          "macr_grid_bce_fwd_bwd")
    return losses, (tuple(grads[k] for k in range(5)) if want_grad else None)


def batch_plan(ids, table_rows, bitmap=None):
    """-> (uniq_rows[n_uniq], seg_off[n_uniq+1], seg_pos[n_ids]) as trimmed device tensors."""
    n = ids.numel()
    dev = ids.device
    uniq = torch.empty(n, dtype=torch.int32, device=dev)
    seg_off = torch.empty(n + 1, dtype=torch.int32, device=dev)
    seg_pos = torch.empty(n, dtype=torch.int32, device=dev)
    n_uniq = torch.zeros(1, dtype=torch.int32, device=dev)
    nbytes = lib().macr_batch_plan_workspace_bytes(n)
    ws = torch.empty(nbytes, dtype=torch.uint8, device=dev)
    check(lib().macr_batch_plan(_i(ids), n, table_rows, ptr(uniq), ptr(seg_off), ptr(seg_pos),
                                ptr(n_uniq), ptr(bitmap), ptr(ws), nbytes, stream_ptr()),
          "macr_batch_plan")
    k = int(n_uniq.item())
    return uniq[:k], seg_off[: k + 1], seg_pos


def adam_sweep_untouched(var, m, v, bitmap, lr_t, beta1=0.9, beta2=0.999, eps=1e-8):
    check(lib().macr_adam_sweep_untouched(_f(var), _f(m), _f(v), var.shape[0], var.shape[1],
                                          ptr(bitmap), lr_t, beta1, beta2, eps, stream_ptr()),
          "macr_adam_sweep_untouched")


def adam_rows(var, m, v, uniq_rows, grad_rows, bitmap, lr_t, beta1=0.9, beta2=0.999, eps=1e-8):
    check(lib().macr_adam_rows(_f(var), _f(m), _f(v), var.shape[0], var.shape[1], _i(uniq_rows),
                               _f(grad_rows), uniq_rows.numel(), ptr(bitmap), lr_t, beta1, beta2,
                               eps, stream_ptr()), "macr_adam_rows")


def adam_dense(var, m, v, grad, lr_t, beta1=0.9, beta2=0.999, eps=1e-8):
    check(lib().macr_adam_dense(_f(var), _f(m), _f(v), _f(grad), var.shape[0], var.shape[1], lr_t,
                                beta1, beta2, eps, stream_ptr()), "macr_adam_dense")


def adam_vec(var, m, v, grad, lr_t, beta1=0.9, beta2=0.999, eps=1e-8):
    check(lib().macr_adam_vec(_f(var), _f(m), _f(v), _f(grad), var.numel(), lr_t, beta1, beta2,
                              eps, stream_ptr()), "macr_adam_vec")


class SpmmPlan:
    """Static segment decomposition of one adjacency (macr_spmm_plan_*): build once per graph."""

    def __init__(self, rowptr):
        self._h = C.c_void_p()
        self.n_rows = rowptr.numel() - 1
        check(lib().macr_spmm_plan_create(_i(rowptr), self.n_rows, C.byref(self._h)),
              "macr_spmm_plan_create")

    def close(self):
        if self._h:
            lib().macr_spmm_plan_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


def spmm_csr(rowptr, col, val, X, plan=None):
    Y = torch.empty_like(X)
    if plan is not None:
        check(lib().macr_spmm_csr_planned(plan._h, _i(rowptr), _i(col), _f(val), X.shape[0], _f(X),
                                          X.shape[1], _f(Y), stream_ptr()), "macr_spmm_csr_planned")
        return Y
    check(lib().macr_spmm_csr(_i(rowptr), _i(col), _f(val), X.shape[0], _f(X), X.shape[1], _f(Y),
                              stream_ptr()), "macr_spmm_csr")
    return Y


def lgcn_propagate(rowptr, col, val, U, I, n_layers, plan=None):
    N = U.shape[0] + I.shape[0]
    E = torch.empty((N, U.shape[1]), dtype=torch.float32, device=U.device)
    tmp = torch.empty((2 * N, U.shape[1]), dtype=torch.float32, device=U.device)
    if plan is not None:
        check(lib().macr_lgcn_propagate_planned(plan._h, _i(rowptr), _i(col), _f(val), _f(U), U.shape[0],
                                                _f(I), I.shape[0], U.shape[1], n_layers, _f(E), _f(tmp),
                                                stream_ptr()), "macr_lgcn_propagate_planned")
        return E
    check(lib().macr_lgcn_propagate(_i(rowptr), _i(col), _f(val), _f(U), U.shape[0], _f(I),
                                    I.shape[0], U.shape[1], n_layers, _f(E), _f(tmp), stream_ptr()),
          "macr_lgcn_propagate")
    return E


def score_gates(rows, wvec):
    sig = torch.empty(rows.shape[0], dtype=torch.float32, device=rows.device)
    check(lib().macr_score_gates(_f(rows), rows.shape[0], rows.shape[1], _f(wvec), _f(sig),
                                 stream_ptr()), "macr_score_gates")
    return sig


def gather_rows(table, ids):
    out = torch.empty((ids.numel(), table.shape[1]), dtype=torch.float32, device=table.device)
    check(lib().macr_gather_rows(_f(table), _i(ids), ids.numel(), table.shape[1], _f(out),
                                 stream_ptr()), "macr_gather_rows")
    return out


TC_MIN_ITEMS = 4096  # smaller catalogues go to the exact fp32 kernel (the tcgen05 path needs >= 2048)


def _topk_out(out, T, K, dev):
    """(ids int32 [T,K], scores f32 [T,K]): fresh tensors, or the caller's (e.g. peer-mapped) ones."""
    if out is None:
        return torch.empty((T, K), dtype=torch.int32, device=dev), torch.empty((T, K), dtype=torch.float32, device=dev)
    ids, sc = out
    if tuple(ids.shape) != (T, K) or tuple(sc.shape) != (T, K):
        raise MacrError(f"out tensors must be [{T},{K}]")
    return _cuda(ids, torch.int32), _cuda(sc, torch.float32)


def _aligned_bytes(nbytes, dev):
    """-> (owner tensor, 1024-byte aligned address) of a device scratch buffer"""
    buf = torch.empty(nbytes + 1024, dtype=torch.uint8, device=dev)
    return buf, buf.data_ptr() + (-buf.data_ptr()) % 1024


class TcItems:
    """Item-side operands of the tensor-core scoring path (bf16 rows scaled by sig_i, the
    -c*sig_i pieces), prepared ONCE for (It, sig_i, c) and reused by every query block scored
    against them: `score_topk(..., prepared=TcItems(It, sig_i, c))`.  The caller vouches that
    It / sig_i do not change while the object is in use (one evaluation, one scorer)."""

    def __init__(self, It, sig_i, c):
        self.It, self.sig_i, self.c = _cuda(It, torch.float32), _cuda(sig_i, torch.float32), float(c)
        n_items = It.shape[0]
        self.nbytes = lib().macr_score_tc_items_bytes(n_items)
        self._buf, self.ptr = _aligned_bytes(self.nbytes, It.device)
        check(lib().macr_score_tc_prepare_items(_f(self.It), n_items, It.shape[1], _f(self.sig_i), self.c,
                                                C.c_void_p(self.ptr), self.nbytes, stream_ptr()),
              "macr_score_tc_prepare_items")

    def matches(self, It, sig_i, c):
        return It.data_ptr() == self.It.data_ptr() and It.shape == self.It.shape and \
            sig_i.data_ptr() == self.sig_i.data_ptr() and float(c) == self.c


def score_topk_tc(Uq, It, sig_i, sig_u, c, mask_rowptr, mask_col, K, item_id_offset=0, stats=None, out=None,
                  prepared=None):
    """Tensor-core (tcgen05 + TMA) score + mask + top-K; bit-identical to `score_topk_exact`.
    stats: optional int64[2] device tensor, += {rows re-done by the exact kernel, candidates}.
    prepared: a `TcItems` of the same (It, sig_i, c) -- skips the per-call item preparation."""
    T, n_items = Uq.shape[0], It.shape[0]
    dev = Uq.device
    ids, sc = _topk_out(out, T, K, dev)
    if prepared is not None:
        if not prepared.matches(It, sig_i, c):
            raise MacrError("score_topk_tc: `prepared` was built from other items / gates / c")
        nbytes = lib().macr_score_topk_tc_prepared_workspace_bytes(T, n_items, K)
        ws, at = _aligned_bytes(nbytes, dev)
        check(lib().macr_score_topk_tc_prepared(_f(Uq), T, _f(It), n_items, Uq.shape[1], _f(sig_i), _f(sig_u),
                                                c, ptr(mask_rowptr), ptr(mask_col), K, item_id_offset,
                                                ptr(ids), ptr(sc), C.c_void_p(prepared.ptr), C.c_void_p(at),
                                                nbytes, ptr(stats), stream_ptr()),
              "macr_score_topk_tc_prepared")
        return ids, sc
    nbytes = lib().macr_score_topk_tc_workspace_bytes(T, n_items, K)
    ws, at = _aligned_bytes(nbytes, dev)
    check(lib().macr_score_topk_tc(_f(Uq), T, _f(It), n_items, Uq.shape[1], _f(sig_i), _f(sig_u), c,
                                   ptr(mask_rowptr), ptr(mask_col), K, item_id_offset, ptr(ids),
                                   ptr(sc), C.c_void_p(at), nbytes, ptr(stats),
                                   stream_ptr()), "macr_score_topk_tc")
    return ids, sc


def uses_tc(n_items, K, T=1):
    """does `score_topk` take the tensor-core path for this shape?"""
    return n_items >= TC_MIN_ITEMS and K <= 32 and T > 0


def score_topk(Uq, It, sig_i, sig_u, c, mask_rowptr, mask_col, K, item_id_offset=0, out=None, prepared=None):
    """Fused score + mask + top-K. -> (ids [T,K] int32 global ids, scores [T,K] fp32).
    Catalogues of at least TC_MIN_ITEMS items go to the tcgen05 path, smaller ones to the exact
    fp32 kernel; both give the same bits.  prepared: optional `TcItems` (tensor-core path only)."""
    if uses_tc(It.shape[0], K, Uq.shape[0]):
        return score_topk_tc(Uq, It, sig_i, sig_u, c, mask_rowptr, mask_col, K, item_id_offset, out=out,
                             prepared=prepared)
    return score_topk_exact(Uq, It, sig_i, sig_u, c, mask_rowptr, mask_col, K, item_id_offset, out=out)


def score_topk_exact(Uq, It, sig_i, sig_u, c, mask_rowptr, mask_col, K, item_id_offset=0, out=None):
    """Exact fp32 CUDA-core kernel (score.cu)."""
    T, n_items = Uq.shape[0], It.shape[0]
    dev = Uq.device
    ids, sc = _topk_out(out, T, K, dev)
    nbytes = lib().macr_score_topk_workspace_bytes(T, n_items, K)
    ws = torch.empty(nbytes, dtype=torch.uint8, device=dev)
    check(lib().macr_score_topk(_f(Uq), T, _f(It), n_items, Uq.shape[1], _f(sig_i), _f(sig_u), c,
                                ptr(mask_rowptr), ptr(mask_col), K, item_id_offset, ptr(ids),
                                ptr(sc), ptr(ws), nbytes, stream_ptr()), "macr_score_topk")
    return ids, sc


def score_matrix(Uq, It, sig_i, sig_u, c):
    out = torch.empty((Uq.shape[0], It.shape[0]), dtype=torch.float32, device=Uq.device)
    check(lib().macr_score_matrix(_f(Uq), Uq.shape[0], _f(It), It.shape[0], Uq.shape[1], _f(sig_i),
                                  _f(sig_u), c, _f(out), stream_ptr()), "macr_score_matrix")
    return out


def topk_merge(ids, scores):
    G, T, K = ids.shape
    oi = torch.empty((T, K), dtype=torch.int32, device=ids.device)
    os_ = torch.empty((T, K), dtype=torch.float32, device=ids.device)
    check(lib().macr_topk_merge(_i(ids), _f(scores), T, K, G, ptr(oi), ptr(os_), stream_ptr()),
          "macr_topk_merge")
    return oi, os_


def topk_rows(scores, K):
    rows, cols = scores.shape
    out = torch.empty((rows, K), dtype=torch.int32, device=scores.device)
    check(lib().macr_topk_rows(_f(scores), cols, rows, K, ptr(out), stream_ptr()), "macr_topk_rows")
    return out


def inv_log2_table(K):
    """1/log2(k+2) in double from the HOST libm (evaluate_foldout.h:70-95 uses log2 on the host)."""
    return np.array([1.0 / math.log2(k + 2) for k in range(K)], dtype=np.float64)


def foldout_metrics(topk_ids, truth_rowptr, truth_col):
    T, K = topk_ids.shape
    out = torch.empty((T, 5 * K), dtype=torch.float32, device=topk_ids.device)
    tab = torch.from_numpy(inv_log2_table(K)).to(topk_ids.device)
    check(lib().macr_foldout_metrics(_i(topk_ids), T, K, _i(truth_rowptr), _i(truth_col), ptr(tab),
                                     _f(out), stream_ptr()), "macr_foldout_metrics")
    return out


class _Tables:
    """The four trainable tensors + Adam slots of one model, resident in HBM (row-major fp32)."""

    def __init__(self, U, I, w, wu, device):
        def t(a):  # host arrays are copied to the device; resident fp32 CUDA tensors are adopted as they are
            if isinstance(a, torch.Tensor) and a.is_cuda:
                return _cuda(a, torch.float32)
            return torch.as_tensor(np.ascontiguousarray(a, dtype=np.float32)).to(device).contiguous()

        self.U, self.I = t(U), t(I)
        self.w, self.wu = t(np.asarray(w).reshape(-1)), t(np.asarray(wu).reshape(-1))
        z = torch.zeros_like
        self.mU, self.vU, self.mI, self.vI = z(self.U), z(self.U), z(self.I), z(self.I)
        self.mw, self.vw, self.mwu, self.vwu = z(self.w), z(self.w), z(self.wu), z(self.wu)

    def all(self):
        return [self.U, self.mU, self.vU, self.I, self.mI, self.vI, self.w, self.mw, self.vw,
                self.wu, self.mwu, self.vwu]

    def state_dict(self):
        names = ["U", "mU", "vU", "I", "mI", "vI", "w", "mw", "vw", "wu", "mwu", "vwu"]
        return {k: getattr(self, k).detach().cpu().numpy() for k in names}


class MFTrainer:
    """Handle around macr_mf_trainer_* (one `--train rubibceboth` model)."""

    RUBIBCEBOTH, NORMALBCE, RUBIBCE = 0, 1, 2

    def set_mode(self, mode):
        """RUBIBCEBOTH (default), NORMALBCE (`--train normalbce`, the README's baseline) or
        RUBIBCE (`--train rubibce`: item gate only, w_user untouched)."""
        check(lib().macr_mf_trainer_set_mode(self._h, int(mode)), "macr_mf_trainer_set_mode")

    def _run_host(self, n, B, ids_ptr, loss_ptr):
        return lib().macr_mf_trainer_run_host(self._h, ids_ptr, n, B, loss_ptr)

    def __init__(self, U, I, w, wu, hp, max_batch, device="cuda:0"):
        self.dev = torch.device(device)
        torch.cuda.set_device(self.dev)
        self.tab = _Tables(U, I, w, wu, self.dev)
        self.hp = hp
        self.max_batch = max_batch
        self.stream = torch.cuda.current_stream(self.dev)
        self._h = C.c_void_p()
        t = self.tab
        check(lib().macr_mf_trainer_create(C.byref(self._h), _f(t.U), _f(t.mU), _f(t.vU),
                                           t.U.shape[0], _f(t.I), _f(t.mI), _f(t.vI), t.I.shape[0],
                                           _f(t.w), _f(t.mw), _f(t.vw), _f(t.wu), _f(t.mwu),
                                           _f(t.vwu), t.U.shape[1], max_batch, C.byref(hp),
                                           stream_ptr(self.stream)), "macr_mf_trainer_create")
        self._loss_dev = torch.zeros(4, dtype=torch.float32, device=self.dev)
        self._pin_ids = torch.empty(3 * max_batch, dtype=torch.int32).pin_memory()
        self._host_loss = (C.c_float * 3)()

    def step_device(self, users, pos, neg):
        """ids already in HBM (int32). Returns a device tensor [loss, mf, reg, L_ori]; no sync."""
        B = users.numel()
        check(lib().macr_mf_trainer_step(self._h, _i(users), _i(pos), _i(neg), B,
                                         ptr(self._loss_dev)), "macr_mf_trainer_step")
        return self._loss_dev

    def step_host(self, users, pos, neg):
        """ids as host sequences (what train.py feeds). Returns (loss, mf, reg) python floats."""
        B = len(users)
        if B > self.max_batch:
            raise MacrError(f"batch {B} > max_batch {self.max_batch}")
        buf = self._pin_ids.numpy()
        buf[:B] = users
        buf[B:2 * B] = pos
        buf[2 * B:3 * B] = neg
        base = self._pin_ids.data_ptr()
        check(lib().macr_mf_trainer_step_host(self._h, C.c_void_p(base), C.c_void_p(base + 4 * B),
                                              C.c_void_p(base + 8 * B), B, self._host_loss),
              "macr_mf_trainer_step_host")
        return float(self._host_loss[0]), float(self._host_loss[1]), float(self._host_loss[2])

    def step_pinned(self, ids3):
        """ids3: pinned host int32 tensor [3,B] (users | pos | neg). Returns (loss, mf, reg)."""
        B = ids3.shape[1]
        base = ids3.data_ptr()
        check(lib().macr_mf_trainer_step_host(self._h, C.c_void_p(base), C.c_void_p(base + 4 * B),
                                              C.c_void_p(base + 8 * B), B, self._host_loss),
              "macr_mf_trainer_step_host")
        return float(self._host_loss[0]), float(self._host_loss[1]), float(self._host_loss[2])

    def run(self, batches, losses=None):
        """Epoch mode: batches int32 device [n_steps,3,B]; returns device losses [n_steps,4]."""
        n, three, B = batches.shape
        assert three == 3
        if losses is None:
            losses = torch.empty((n, 4), dtype=torch.float32, device=self.dev)
        check(lib().macr_mf_trainer_run(self._h, _i(batches), n, B, _f(losses)),
              "macr_mf_trainer_run")
        return losses

    def run_host(self, batches_host, losses_host=None):
        """Epoch call with HOST buffers: `batches_host` int32 [n,3,B] (pinned recommended) ->
        float32 [n,4] host losses; one H2D, n graph replays, one D2H, synchronised."""
        n, three, B = batches_host.shape
        assert three == 3 and batches_host.dtype == torch.int32 and not batches_host.is_cuda
        if losses_host is None:
            losses_host = torch.empty((n, 4), dtype=torch.float32)
        check(self._run_host(n, B, C.c_void_p(batches_host.data_ptr()),
                             C.c_void_p(losses_host.data_ptr())), "trainer_run_host")
        return losses_host

    # ---- row-partitioned mode (include/macr_b200.h: macr_mf_trainer_shard) ----
    def ipc_export(self):
        """IPC handle (64 bytes) of the trainer's barrier flags."""
        buf = C.create_string_buffer(64)
        check(lib().macr_mf_trainer_ipc_export(self._h, buf), "macr_mf_trainer_ipc_export")
        return buf.raw

    def shard(self, desc, peer_u_ghost, peer_i_ghost, peer_flags):
        """peer_*: ctypes arrays (c_void_p * world): every peer's ghost bases / flags mapped here.
        From now on step / run / run_host take GLOBAL ids (identical on every rank)."""
        check(lib().macr_mf_trainer_shard(self._h, C.byref(desc), peer_u_ghost, peer_i_ghost, peer_flags),
              "macr_mf_trainer_shard")

    def peer_error(self):
        out = C.c_int(0)
        check(lib().macr_mf_trainer_peer_error(self._h, C.byref(out)), "macr_mf_trainer_peer_error")
        return out.value

    @property
    def launches_per_step(self):
        return int(lib().macr_mf_trainer_launches_per_step(self._h))

    @property
    def steps_done(self):
        return int(lib().macr_mf_trainer_steps_done(self._h))

    def set_steps_done(self, t):
        check(lib().macr_mf_trainer_set_steps_done(self._h, int(t)), "set_steps_done")

    def close(self):
        if self._h:
            lib().macr_mf_trainer_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


class LGCNTrainer:
    """Handle around macr_lgcn_trainer_* (one `--loss bceboth` LightGCN model)."""

    RUBIBCEBOTH, NORMALBCE, RUBIBCE = 0, 1, 2

    def set_mode(self, mode):
        """RUBIBCEBOTH (`--loss bceboth`, default), NORMALBCE (`--loss bce`, the baseline) or
        RUBIBCE (`--loss bce1`: item gate only)."""
        check(lib().macr_lgcn_trainer_set_mode(self._h, int(mode)), "macr_lgcn_trainer_set_mode")

    def _run_host(self, n, B, ids_ptr, loss_ptr, train=True):
        return lib().macr_lgcn_trainer_run_host(self._h, ids_ptr, n, B, 1 if train else 0, loss_ptr)

    def __init__(self, rowptr, col, val, U, I, w, wu, n_layers, hp, max_batch, device="cuda:0"):
        self.dev = torch.device(device)
        torch.cuda.set_device(self.dev)
        self.tab = _Tables(U, I, w, wu, self.dev)
        ti = lambda a: torch.as_tensor(np.ascontiguousarray(a, dtype=np.int32)).to(self.dev)
        self.rowptr, self.col = ti(rowptr), ti(col)
        self.val = torch.as_tensor(np.ascontiguousarray(val, dtype=np.float32)).to(self.dev)
        self.hp, self.n_layers, self.max_batch = hp, n_layers, max_batch
        self.stream = torch.cuda.current_stream(self.dev)
        self._h = C.c_void_p()
        t = self.tab
        check(lib().macr_lgcn_trainer_create(C.byref(self._h), _i(self.rowptr), _i(self.col),
                                             _f(self.val), _f(t.U), _f(t.mU), _f(t.vU),
                                             t.U.shape[0], _f(t.I), _f(t.mI), _f(t.vI),
                                             t.I.shape[0], _f(t.w), _f(t.mw), _f(t.vw), _f(t.wu),
                                             _f(t.mwu), _f(t.vwu), t.U.shape[1], n_layers,
                                             max_batch, C.byref(hp), stream_ptr(self.stream)),
              "macr_lgcn_trainer_create")
        self._loss_dev = torch.zeros(4, dtype=torch.float32, device=self.dev)
        self._pin_ids = torch.empty(3 * max_batch, dtype=torch.int32).pin_memory()
        self._host_loss = (C.c_float * 3)()

    def step_device(self, users, pos, neg, train=True):
        check(lib().macr_lgcn_trainer_step(self._h, _i(users), _i(pos), _i(neg), users.numel(),
                                           1 if train else 0, ptr(self._loss_dev)),
              "macr_lgcn_trainer_step")
        return self._loss_dev

    def step_host(self, users, pos, neg, train=True):
        B = len(users)
        if B > self.max_batch:
            raise MacrError(f"batch {B} > max_batch {self.max_batch}")
        buf = self._pin_ids.numpy()
        buf[:B] = users
        buf[B:2 * B] = pos
        buf[2 * B:3 * B] = neg
        base = self._pin_ids.data_ptr()
        check(lib().macr_lgcn_trainer_step_host(self._h, C.c_void_p(base), C.c_void_p(base + 4 * B),
                                                C.c_void_p(base + 8 * B), B, 1 if train else 0,
                                                self._host_loss), "macr_lgcn_trainer_step_host")
        return float(self._host_loss[0]), float(self._host_loss[1]), float(self._host_loss[2])

    def run(self, batches, train=True, losses=None):
        n, three, B = batches.shape
        assert three == 3
        if losses is None:
            losses = torch.empty((n, 4), dtype=torch.float32, device=self.dev)
        check(lib().macr_lgcn_trainer_run(self._h, _i(batches), n, B, 1 if train else 0,
                                          _f(losses)), "macr_lgcn_trainer_run")
        return losses

    def embeddings(self):
        """Propagated tables of the current parameters: (users [U,64], items [I,64]) views."""
        out = C.c_void_p()
        check(lib().macr_lgcn_trainer_embeddings(self._h, C.byref(out)),
              "macr_lgcn_trainer_embeddings")
        nu, ni, d = self.tab.U.shape[0], self.tab.I.shape[0], self.tab.U.shape[1]
        E = _wrap_device_ptr(out.value, (nu + ni, d), self.dev)
        return E[:nu], E[nu:]

    def run_host(self, batches_host, losses_host=None):
        """Epoch call with HOST buffers: `batches_host` int32 [n,3,B] (pinned recommended) ->
        float32 [n,4] host losses; one H2D, n graph replays, one D2H, synchronised."""
        n, three, B = batches_host.shape
        assert three == 3 and batches_host.dtype == torch.int32 and not batches_host.is_cuda
        if losses_host is None:
            losses_host = torch.empty((n, 4), dtype=torch.float32)
        check(self._run_host(n, B, C.c_void_p(batches_host.data_ptr()),
                             C.c_void_p(losses_host.data_ptr())), "trainer_run_host")
        return losses_host

    def _run_host_train(self, batches_host, train=True):
        """run_host with the train / loss-only switch -> float32 [n,4] numpy losses."""
        n, three, B = batches_host.shape
        assert three == 3 and batches_host.dtype == torch.int32 and not batches_host.is_cuda
        losses_host = torch.empty((n, 4), dtype=torch.float32)
        check(self._run_host(n, B, C.c_void_p(batches_host.data_ptr()), C.c_void_p(losses_host.data_ptr()),
                             train), "macr_lgcn_trainer_run_host")
        return losses_host.numpy()

    # ---- row-partitioned mode (include/macr_b200.h: macr_lgcn_trainer_shard) ----
    def ipc_export(self):
        """IPC handles of the trainer's {E_mean, layer buffers, flags}: 3 x 64 bytes."""
        buf = C.create_string_buffer(3 * 64)
        check(lib().macr_lgcn_trainer_ipc_export(self._h, buf), "macr_lgcn_trainer_ipc_export")
        return [buf.raw[64 * k:64 * (k + 1)] for k in range(3)]

    def shard(self, desc, peer_u, peer_i, peer_e, peer_t, peer_f):
        """peer_*: ctypes arrays (c_void_p * world) of every peer's buffers mapped into this process."""
        check(lib().macr_lgcn_trainer_shard(self._h, C.byref(desc), peer_u, peer_i, peer_e, peer_t, peer_f),
              "macr_lgcn_trainer_shard")

    def peer_error(self):
        out = C.c_int(0)
        check(lib().macr_lgcn_trainer_peer_error(self._h, C.byref(out)), "macr_lgcn_trainer_peer_error")
        return out.value

    @property
    def launches_per_step(self):
        return int(lib().macr_lgcn_trainer_launches_per_step(self._h))

    @property
    def steps_done(self):
        return int(lib().macr_lgcn_trainer_steps_done(self._h))

    def set_steps_done(self, t):
        check(lib().macr_lgcn_trainer_set_steps_done(self._h, int(t)), "set_steps_done")

    def close(self):
        if self._h:
            lib().macr_lgcn_trainer_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


# ------------------------------------------------------------------------------------------------
# row-partitioned tables (csrc/shard.cu): peer-mappable memory, the fused exchange kernels
# ------------------------------------------------------------------------------------------------
class IpcBuffer:
    """cudaMalloc'ed, zero-filled device memory with its CUDA IPC handle (64 opaque bytes) so the
    other ranks of the box can map it (`IpcBuffer.open_peer`)."""

    def __init__(self, nbytes, device):
        self.dev = torch.device(device)
        torch.cuda.set_device(self.dev)
        out, handle = C.c_void_p(), C.create_string_buffer(64)
        check(lib().macr_ipc_alloc(int(nbytes), C.byref(out), handle), "macr_ipc_alloc")
        self.ptr, self.handle, self.nbytes = out.value, handle.raw, int(nbytes)
        self.as_f32((self.nbytes // 4,)).zero_()

    def as_f32(self, shape):
        return _wrap_device_ptr(self.ptr, tuple(shape), self.dev)

    @staticmethod
    def open_peer(handle):
        out = C.c_void_p()
        check(lib().macr_ipc_open(handle, C.byref(out)), "macr_ipc_open")
        return out.value

    @staticmethod
    def close_peer(p):
        if p:
            lib().macr_ipc_close(C.c_void_p(p))

    def free(self):
        if self.ptr:
            lib().macr_ipc_free(C.c_void_p(self.ptr))
            self.ptr = None


def shard_table(rows, device, peer_mappable=False):
    """Zero-filled [rows, 64] fp32 table; peer_mappable -> (tensor view, IpcBuffer) else (tensor, None)."""
    if peer_mappable:
        buf = IpcBuffer(rows * D * 4, device)
        return buf.as_f32((rows, D)), buf
    return torch.zeros((rows, D), dtype=torch.float32, device=device), None


def shard_pack(U_local, I_local, desc, ids3, B, parity, local3, ex):
    check(lib().macr_shard_pack(_f(U_local), _f(I_local), C.byref(desc), _i(ids3), B, parity, _i(local3),
                                _f(ex), stream_ptr()), "macr_shard_pack")


def shard_unpack(U_local, I_local, desc, ex, B, parity):
    check(lib().macr_shard_unpack(_f(U_local), _f(I_local), C.byref(desc), _f(ex), B, parity, stream_ptr()),
          "macr_shard_unpack")


def shard_push(U_local, I_local, desc, ids3, B, parity, local3, peer_u, peer_i):
    """peer_u / peer_i: ctypes arrays (c_void_p * world) of the peers' ghost bases in this process."""
    check(lib().macr_shard_push(_f(U_local), _f(I_local), C.byref(desc), _i(ids3), B, parity, _i(local3),
                                peer_u, peer_i, stream_ptr()), "macr_shard_push")


def topk_merge_peers(cand_ids, cand_sc, out_ids, out_sc, world, K, row0, rows):
    """cand_* / out_*: ctypes arrays (c_void_p * world) of every rank's buffers mapped into this process."""
    check(lib().macr_topk_merge_peers(cand_ids, cand_sc, out_ids, out_sc, world, K, row0, rows, stream_ptr()),
          "macr_topk_merge_peers")


def shard_barrier(peer_flags, rank, world, epoch, err_flag):
    check(lib().macr_shard_barrier(peer_flags, rank, world, epoch, _i(err_flag), stream_ptr()),
          "macr_shard_barrier")


class _CudaArrayView:
    def __init__(self, p, shape):
        n = int(np.prod(shape))
        self.__cuda_array_interface__ = {"shape": (n,), "typestr": "<f4", "data": (p, False),
                                         "version": 3}


def _wrap_device_ptr(p, shape, device):
    """Zero-copy torch view of library-owned device memory (float32)."""
    with torch.cuda.device(device):
        flat = torch.as_tensor(_CudaArrayView(p, shape), device=device)
    return flat.view(*shape)
