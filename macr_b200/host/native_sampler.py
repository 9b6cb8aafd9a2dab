"""Native (C) batch samplers behind `Data.sample()` -- same Mersenne-Twister streams as the
reference's pure-Python samplers, ~100x faster (macr_b200/csrc/sampler.cu, include/macr_b200.h).

The interpreter-side generators are the source of truth: every call takes `random.getstate()`
(and `np.random.get_state()` for LightGCN), lets the C code advance the 624-word states, and
puts them back, so anything that draws from `random` / `np.random` afterwards sees exactly the
stream position the reference's sampler would have left.
"""
import ctypes as C
import random

import numpy as np

from .._lib import check, lib


def _p(a):
    return C.c_void_p(a.ctypes.data)


class ListCSR:
    """Per-user id lists as int64 CSR over the user id: `order` keeps list order (what
    `choice` / `randint` index), `sorted` the same ids ascending (membership tests)."""

    def __init__(self, lists, n_users):
        self.rowptr = np.zeros(n_users + 1, np.int64)
        for u, items in lists.items():
            if 0 <= u < n_users:
                self.rowptr[u + 1] = len(items)
        np.cumsum(self.rowptr, out=self.rowptr)
        self.order = np.zeros(max(1, int(self.rowptr[-1])), np.int32)
        self.sorted = np.zeros_like(self.order)
        for u, items in lists.items():
            if 0 <= u < n_users and len(items):
                lo = self.rowptr[u]
                a = np.asarray(items, np.int32)
                self.order[lo:lo + len(a)] = a
                self.sorted[lo:lo + len(a)] = np.sort(a)
        self._tags = None

    def pair_tags(self, log2=None):
        """Hashed (user, id) pair set over `sorted` for the epoch samplers (built once):
        -> (uint16 tags [buckets, 8], log2 of the bucket count), ~2 pairs per bucket unless
        `log2` says otherwise (a crowded table costs exact look-ups, never a wrong answer)."""
        if self._tags is None or log2 is not None:
            n = int(self.rowptr[-1])
            if log2 is None:
                log2 = max(1, int(np.ceil(np.log2(max(1, n) / 2.0))))
            raw = np.empty((8 << log2) + 8, np.uint16)
            off = (-raw.ctypes.data % 16) // 2          # 16-byte aligned view
            tags = raw[off:off + (8 << log2)]
            check(lib().macr_pairset_build(_p(self.rowptr), _p(self.sorted), len(self.rowptr) - 1,
                                           _p(tags), log2), "macr_pairset_build")
            self._tags = (tags, log2)
        return self._tags


def _py_state():
    version, internal, gauss = random.getstate()
    return np.array(internal, dtype=np.uint32), (version, gauss)


def _py_restore(buf, meta):
    random.setstate((meta[0], tuple(int(x) for x in buf), meta[1]))


def _np_state():
    name, keys, pos, has_gauss, cached = np.random.get_state()
    buf = np.empty(625, np.uint32)
    buf[:624] = keys
    buf[624] = pos
    return buf, (name, has_gauss, cached)


def _np_restore(buf, meta):
    np.random.set_state((meta[0], buf[:624].copy(), int(buf[624]), meta[1], meta[2]))


def _check_population(B, n_users, n_pop):
    """`random.sample(pop, B)` of the reference (taken when B <= n_users) raises ValueError when
    the population is smaller than B (users without a train line are not in it): same error here,
    before any generator state moves."""
    if B <= n_users and B > n_pop:
        raise ValueError("Sample larger than population or is negative")


def sample_mf(users_pop, n_users, n_items, csr, B):
    """-> (users, pos, neg) int32 arrays; advances `random` exactly like load_data.py:543-566."""
    _check_population(B, n_users, len(users_pop))
    st, meta = _py_state()
    out = np.empty((3, B), np.int32)
    check(lib().macr_sample_mf(_p(st), _p(users_pop), len(users_pop), n_users, n_items,
                               _p(csr.rowptr), _p(csr.order), _p(csr.sorted), B, _p(out[0]),
                               _p(out[1]), _p(out[2])), "macr_sample_mf")
    _py_restore(st, meta)
    return out[0], out[1], out[2]


def sample_mf_epoch(users_pop, n_users, n_items, csr, B, n_batches, out=None):
    """n_batches consecutive `sample()` calls in one native call -> int32 [n_batches, 3, B] (the
    layout `MFTrainer.run_host` takes); speculative chunked draws (macr_sample_mf_epoch)."""
    _check_population(B, n_users, len(users_pop))
    st, meta = _py_state()
    if out is None:
        out = np.empty((n_batches, 3, B), np.int32)
    assert out.dtype == np.int32 and out.flags.c_contiguous and out.shape == (n_batches, 3, B)
    tags, log2 = csr.pair_tags()
    check(lib().macr_sample_mf_epoch(_p(st), _p(users_pop), len(users_pop), n_users, n_items,
                                     _p(csr.rowptr), _p(csr.order), _p(csr.sorted), _p(tags), log2,
                                     B, n_batches, _p(out)), "macr_sample_mf_epoch")
    _py_restore(st, meta)
    return out


def sample_lgcn_epoch(users_pop, n_users, n_items, pos_csr, ban_csr, B, n_batches, out=None):
    _check_population(B, n_users, len(users_pop))
    st, meta = _py_state()
    nst, nmeta = _np_state()
    if out is None:
        out = np.empty((n_batches, 3, B), np.int32)
    assert out.dtype == np.int32 and out.flags.c_contiguous and out.shape == (n_batches, 3, B)
    tags, log2 = ban_csr.pair_tags()
    check(lib().macr_sample_lgcn_epoch(_p(st), _p(nst), _p(users_pop), len(users_pop), n_users,
                                       n_items, _p(pos_csr.rowptr), _p(pos_csr.order),
                                       _p(ban_csr.rowptr), _p(ban_csr.sorted), _p(tags), log2, B,
                                       n_batches, _p(out)), "macr_sample_lgcn_epoch")
    _py_restore(st, meta)
    _np_restore(nst, nmeta)
    return out


def sample_lgcn(users_pop, n_users, n_items, pos_csr, ban_csr, B):
    """-> (users, pos, neg); advances `random` and `np.random` like utility/load_data.py:174-212."""
    _check_population(B, n_users, len(users_pop))
    st, meta = _py_state()
    nst, nmeta = _np_state()
    out = np.empty((3, B), np.int32)
    check(lib().macr_sample_lgcn(_p(st), _p(nst), _p(users_pop), len(users_pop), n_users, n_items,
                                 _p(pos_csr.rowptr), _p(pos_csr.order), _p(ban_csr.rowptr),
                                 _p(ban_csr.sorted), B, _p(out[0]), _p(out[1]), _p(out[2])),
          "macr_sample_lgcn")
    _py_restore(st, meta)
    _np_restore(nst, nmeta)
    return out[0], out[1], out[2]
