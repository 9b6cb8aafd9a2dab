"""ctypes front-end of oracle/libmacr_oracle.so (numpy in, numpy out).  TEST INFRASTRUCTURE ONLY."""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None

f32p = C.POINTER(C.c_float)
i32p = C.POINTER(C.c_int32)
f64p = C.POINTER(C.c_double)


class HParams(C.Structure):
    """Mirror of oracle_hparams / macr_hparams (include/macr_b200.h)."""

    _fields_ = [
        ("lr", C.c_float),
        ("beta1", C.c_float),
        ("beta2", C.c_float),
        ("eps", C.c_float),
        ("alpha", C.c_float),
        ("beta", C.c_float),
        ("decay", C.c_float),
        ("batch_size_flag", C.c_int32),
    ]

    @classmethod
    def make(cls, lr=1e-3, alpha=1e-3, beta=1e-3, decay=1e-5, batch_size=1024, beta1=0.9,
             beta2=0.999, eps=1e-8):
        return cls(lr, beta1, beta2, eps, alpha, beta, decay, batch_size)


def build(force=False):
    """Compile the C restatement (and oracle/_ref when /root/reference is present)."""
    so = os.path.join(_HERE, "libmacr_oracle.so")
    src = os.path.join(_HERE, "macr_oracle.c")
    if force or not os.path.exists(so) or os.path.getmtime(so) < os.path.getmtime(src):
        subprocess.run(["make", "-C", _HERE, "libmacr_oracle.so"], check=True,
                       stdout=subprocess.DEVNULL)
    ref_so = os.path.join(_HERE, "_ref", "libmacr_ref_eval.so")
    if os.path.isdir("/root/reference") and (force or not os.path.exists(ref_so)):
        subprocess.run(["make", "-C", _HERE, "ref"], check=True, stdout=subprocess.DEVNULL)
    return so


def lib():
    global _LIB
    if _LIB is None:
        _LIB = C.CDLL(build())
        _LIB.oracle_adam_lr_t.restype = C.c_float
        _LIB.oracle_adam_lr_t.argtypes = [C.c_float, C.c_float, C.c_float]
        _LIB.oracle_get_threads.restype = C.c_int
    return _LIB


def _f(a):
    assert a.dtype == np.float32 and a.flags.c_contiguous, (a.dtype, a.flags)
    return a.ctypes.data_as(f32p)


def _i(a):
    assert a.dtype == np.int32 and a.flags.c_contiguous
    return a.ctypes.data_as(i32p)


def _ids(x):
    return np.ascontiguousarray(np.asarray(x, dtype=np.int32))


def set_threads(n):
    lib().oracle_set_threads(C.c_int(int(n)))


def get_threads():
    return int(lib().oracle_get_threads())


def gather_dots(Ue, Ie, Ur, Ir, w, wu, u, p, n):
    u, p, n = _ids(u), _ids(p), _ids(n)
    B, d = len(u), Ue.shape[1]
    out = [np.empty(B, np.float32) for _ in range(6)]
    lib().oracle_gather_dots(_f(Ue), _f(Ie), _f(Ur), _f(Ir), _f(w.reshape(-1)),
                             _f(wu.reshape(-1)), _i(u), _i(p), _i(n), C.c_int(B), C.c_int(d),
                             *[_f(o) for o in out])
    return tuple(out)  # yp, yn, sp, sn, su, regsq


def grid_bce(yp, yn, sp, sn, su, alpha, beta, want_grad=True):
    B = len(yp)
    losses = np.empty(3, np.float32)
    g = [np.empty(B, np.float32) for _ in range(5)]
    lib().oracle_grid_bce(_f(yp), _f(yn), _f(sp), _f(sn), _f(su), C.c_int(B), C.c_float(alpha),
                          C.c_float(beta), _f(losses),
                          *([_f(x) for x in g] if want_grad else [None] * 5))
    return losses, (tuple(g) if want_grad else None)


def adam_lr_t(lr, b1p, b2p):
    return float(lib().oracle_adam_lr_t(C.c_float(lr), C.c_float(b1p), C.c_float(b2p)))


def adam_sparse(var, m, v, idx, grad_rows, lr_t, beta1=0.9, beta2=0.999, eps=1e-8):
    idx = _ids(idx)
    rows, d = var.shape
    lib().oracle_adam_sparse(_f(var), _f(m), _f(v), C.c_int64(rows), C.c_int(d), _i(idx),
                             _f(grad_rows), C.c_int(len(idx)), C.c_float(lr_t),
                             C.c_float(beta1), C.c_float(beta2), C.c_float(eps))


def adam_dense_vec(var, m, v, g, lr_t, beta1=0.9, beta2=0.999, eps=1e-8):
    lib().oracle_adam_dense_vec(_f(var), _f(m), _f(v), _f(g), C.c_int(var.size),
                                C.c_float(lr_t), C.c_float(beta1), C.c_float(beta2),
                                C.c_float(eps))


class MFState:
    """Host copy of everything one MF / LightGCN model owns (tables + Adam slots)."""

    def __init__(self, U, I, w, wu):
        c = lambda a: np.ascontiguousarray(a, dtype=np.float32).copy()
        self.U, self.I, self.w, self.wu = c(U), c(I), c(w).reshape(-1), c(wu).reshape(-1)
        z = np.zeros_like
        self.mU, self.vU, self.mI, self.vI = z(self.U), z(self.U), z(self.I), z(self.I)
        self.mw, self.vw, self.mwu, self.vwu = z(self.w), z(self.w), z(self.wu), z(self.wu)
        self.pw = np.array([0.9, 0.999], np.float32)  # beta1_power, beta2_power
        self.t = 0

    def tables(self):
        return [self.U, self.mU, self.vU, self.I, self.mI, self.vI, self.w, self.mw, self.vw,
                self.wu, self.mwu, self.vwu]


def mf_step(st, u, p, n, hp):
    """One `--train rubibceboth` step on MFState `st` (in place). Returns (loss, mf, reg, L_ori)."""
    u, p, n = _ids(u), _ids(p), _ids(n)
    if st.t == 0:
        st.pw[:] = (hp.beta1, hp.beta2)
    losses = np.empty(4, np.float32)
    lib().oracle_mf_step(_f(st.U), _f(st.mU), _f(st.vU), C.c_int64(st.U.shape[0]), _f(st.I),
                         _f(st.mI), _f(st.vI), C.c_int64(st.I.shape[0]), _f(st.w), _f(st.mw),
                         _f(st.vw), _f(st.wu), _f(st.mwu), _f(st.vwu), C.c_int(st.U.shape[1]),
                         _i(u), _i(p), _i(n), C.c_int(len(u)), C.byref(hp), _f(st.pw),
                         _f(losses))
    st.t += 1
    return losses


def mf_step_normal(st, u, p, n, hp):
    """One `--train normalbce` step (model.py:277-287) on MFState `st` (in place)."""
    u, p, n = _ids(u), _ids(p), _ids(n)
    if st.t == 0:
        st.pw[:] = (hp.beta1, hp.beta2)
    losses = np.empty(4, np.float32)
    lib().oracle_mf_step_normal(_f(st.U), _f(st.mU), _f(st.vU), C.c_int64(st.U.shape[0]), _f(st.I),
                                _f(st.mI), _f(st.vI), C.c_int64(st.I.shape[0]),
                                C.c_int(st.U.shape[1]), _i(u), _i(p), _i(n), C.c_int(len(u)),
                                C.byref(hp), _f(st.pw), _f(losses))
    st.t += 1
    return losses


def mf_step_item(st, u, p, n, hp):
    """One `--train rubibce` step (model.py:158-183: item gate only) on MFState `st` (in place)."""
    u, p, n = _ids(u), _ids(p), _ids(n)
    if st.t == 0:
        st.pw[:] = (hp.beta1, hp.beta2)
    losses = np.empty(4, np.float32)
    lib().oracle_mf_step_item(_f(st.U), _f(st.mU), _f(st.vU), C.c_int64(st.U.shape[0]), _f(st.I),
                              _f(st.mI), _f(st.vI), C.c_int64(st.I.shape[0]), _f(st.w), _f(st.mw),
                              _f(st.vw), _f(st.wu), C.c_int(st.U.shape[1]), _i(u), _i(p), _i(n),
                              C.c_int(len(u)), C.byref(hp), _f(st.pw), _f(losses))
    st.t += 1
    return losses


def grid_bce_item(yp, yn, sp, sn, alpha, want_grad=True):
    """model.py:172-177 grid + item branch; returns (losses3, d_yp, d_yn, d_sp, d_sn)."""
    B = len(yp)
    c = lambda a: np.ascontiguousarray(a, dtype=np.float32)
    yp, yn, sp, sn = c(yp), c(yn), c(sp), c(sn)
    l3 = np.zeros(3, np.float32)
    outs = [np.zeros(B, np.float32) for _ in range(4)]
    lib().oracle_grid_bce_item(_f(yp), _f(yn), _f(sp), _f(sn), C.c_int(B), C.c_float(alpha), _f(l3),
                               *( [_f(o) for o in outs] if want_grad else [None] * 4))
    return (l3, *outs)


def spmm_csr(rowptr, col, val, X):
    n, d = X.shape
    Y = np.empty_like(X)
    lib().oracle_spmm_csr(_i(rowptr), _i(col), _f(val), C.c_int64(n), _f(X), C.c_int(d), _f(Y))
    return Y


def lgcn_propagate(rowptr, col, val, U, I, n_layers):
    d = U.shape[1]
    E = np.empty((U.shape[0] + I.shape[0], d), np.float32)
    lib().oracle_lgcn_propagate(_i(rowptr), _i(col), _f(val), _f(U), C.c_int64(U.shape[0]),
                                _f(I), C.c_int64(I.shape[0]), C.c_int(d), C.c_int(n_layers),
                                _f(E))
    return E


def lgcn_step(st, rowptr, col, val, n_layers, u, p, n, hp, train=True):
    u, p, n = _ids(u), _ids(p), _ids(n)
    if st.t == 0:
        st.pw[:] = (hp.beta1, hp.beta2)
    losses = np.empty(4, np.float32)
    lib().oracle_lgcn_step(_i(rowptr), _i(col), _f(val), _f(st.U), _f(st.mU), _f(st.vU),
                           C.c_int64(st.U.shape[0]), _f(st.I), _f(st.mI), _f(st.vI),
                           C.c_int64(st.I.shape[0]), _f(st.w), _f(st.mw), _f(st.vw), _f(st.wu),
                           _f(st.mwu), _f(st.vwu), C.c_int(st.U.shape[1]), C.c_int(n_layers),
                           _i(u), _i(p), _i(n), C.c_int(len(u)), C.c_int(1 if train else 0),
                           C.byref(hp), _f(st.pw), _f(losses))
    if train:
        st.t += 1
    return losses


def lgcn_step_normal(st, rowptr, col, val, n_layers, u, p, n, hp, train=True):
    """One `--loss bce` LightGCN step (LightGCN.py:415-429,:186) on MFState `st` (in place)."""
    u, p, n = _ids(u), _ids(p), _ids(n)
    if st.t == 0:
        st.pw[:] = (hp.beta1, hp.beta2)
    losses = np.empty(4, np.float32)
    lib().oracle_lgcn_step_normal(_i(rowptr), _i(col), _f(val), _f(st.U), _f(st.mU), _f(st.vU),
                                  C.c_int64(st.U.shape[0]), _f(st.I), _f(st.mI), _f(st.vI),
                                  C.c_int64(st.I.shape[0]), _f(st.w), _f(st.wu),
                                  C.c_int(st.U.shape[1]), C.c_int(n_layers), _i(u), _i(p), _i(n),
                                  C.c_int(len(u)), C.c_int(1 if train else 0), C.byref(hp),
                                  _f(st.pw), _f(losses))
    if train:
        st.t += 1
    return losses


def lgcn_step_item(st, rowptr, col, val, n_layers, u, p, n, hp, train=True):
    """One `--loss bce1` LightGCN step (LightGCN.py:431-461) on MFState `st` (in place)."""
    u, p, n = _ids(u), _ids(p), _ids(n)
    if st.t == 0:
        st.pw[:] = (hp.beta1, hp.beta2)
    losses = np.empty(4, np.float32)
    lib().oracle_lgcn_step_item(_i(rowptr), _i(col), _f(val), _f(st.U), _f(st.mU), _f(st.vU),
                                C.c_int64(st.U.shape[0]), _f(st.I), _f(st.mI), _f(st.vI),
                                C.c_int64(st.I.shape[0]), _f(st.w), _f(st.mw), _f(st.vw),
                                _f(st.wu), C.c_int(st.U.shape[1]), C.c_int(n_layers), _i(u), _i(p),
                                _i(n), C.c_int(len(u)), C.c_int(1 if train else 0), C.byref(hp),
                                _f(st.pw), _f(losses))
    if train:
        st.t += 1
    return losses


def score_gates(rows, wvec):
    sig = np.empty(rows.shape[0], np.float32)
    lib().oracle_score_gates(_f(rows), C.c_int64(rows.shape[0]), C.c_int(rows.shape[1]),
                             _f(wvec.reshape(-1)), _f(sig))
    return sig


def score_matrix(Uq, It, sig_i, sig_u, c):
    out = np.empty((Uq.shape[0], It.shape[0]), np.float32)
    lib().oracle_score_matrix(_f(Uq), C.c_int(Uq.shape[0]), _f(It), C.c_int64(It.shape[0]),
                              C.c_int(Uq.shape[1]), _f(sig_i), _f(sig_u), C.c_float(c), _f(out))
    return out


def score_topk(Uq, It, sig_i, sig_u, c, mask_rowptr, mask_col, K, item_id_offset=0):
    T = Uq.shape[0]
    ids = np.empty((T, K), np.int32)
    sc = np.empty((T, K), np.float32)
    lib().oracle_score_topk(_f(Uq), C.c_int(T), _f(It), C.c_int64(It.shape[0]),
                            C.c_int(Uq.shape[1]), _f(sig_i), _f(sig_u), C.c_float(c),
                            _i(mask_rowptr) if mask_rowptr is not None else None,
                            _i(mask_col) if mask_col is not None else None, C.c_int(K),
                            C.c_int32(item_id_offset), _i(ids), _f(sc))
    return ids, sc


def topk_rows(scores, K):
    rows, cols = scores.shape
    out = np.empty((rows, K), np.int32)
    lib().oracle_topk_rows(_f(scores), C.c_int(cols), C.c_int(rows), C.c_int(K), _i(out))
    return out


def topk_merge(ids, scores):
    G, T, K = ids.shape
    oi = np.empty((T, K), np.int32)
    os_ = np.empty((T, K), np.float32)
    lib().oracle_topk_merge(_i(ids), _f(scores), C.c_int(T), C.c_int(K), C.c_int(G), _i(oi),
                            _f(os_))
    return oi, os_


def foldout_metrics(topk_ids, truth_rowptr, truth_col):
    T, K = topk_ids.shape
    out = np.empty((T, 5 * K), np.float32)
    lib().oracle_foldout_metrics(_i(topk_ids), C.c_int(T), C.c_int(K), _i(truth_rowptr),
                                 _i(truth_col), _f(out))
    return out


def inv_log2_table(K):
    out = np.empty(K, np.float64)
    lib().oracle_inv_log2_table(C.c_int(K), out.ctypes.data_as(f64p))
    return out
