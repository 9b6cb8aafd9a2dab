"""LightGCN data set, normalised adjacency and triple sampler -- host mirror of
macr_lightgcn/utility/load_data.py (reference lines cited).

The adjacency is built straight into CSR with numpy (the reference goes dok -> lil -> dok ->
csr, load_data.py:126-139, minutes at scale); values reproduce its float32 arithmetic
`D^-1/2 . A . D^-1/2` evaluated as `(d[r] * 1) * d[c]` (load_data.py:112-121), zero-degree
nodes -> 0.  The sampler consumes `random` and legacy `np.random` call for call like
`Data.sample()` (load_data.py:174-212).
"""
import random

import numpy as np
import scipy.sparse as sp

from .data_mf import cached_sorted_csr, lists_to_csr


class Data:
    def __init__(self, path, batch_size, args=None):
        self.path = path
        self.batch_size = batch_size
        valid_set = getattr(args, "valid_set", "test")
        train_file = path + "/train.txt"
        test_file = path + ("/test.txt" if valid_set == "test" else "/valid.txt")  # :19-23

        self.n_users = self.n_items = 0
        self.n_train = self.n_test = 0
        self.exist_users = []          # FILE order: random.sample() indexes into it (:36,176)
        self.train_items, self.test_set = {}, {}
        rows, cols = [], []
        with open(train_file) as f:
            for raw in f:
                tok = raw.strip("\n").split(" ")
                if tok == [""]:
                    continue
                uid, items = int(tok[0]), [int(t) for t in tok[1:] if t != ""]
                self.exist_users.append(uid)
                self.n_users = max(self.n_users, uid)
                if items:
                    self.n_items = max(self.n_items, max(items))
                    self.train_items[uid] = items
                    rows.append(np.full(len(items), uid, np.int64))
                    cols.append(np.asarray(items, np.int64))
                self.n_train += len(items)
        with open(test_file) as f:
            for raw in f:
                tok = raw.strip("\n").split(" ")
                try:
                    ids = [int(t) for t in tok]
                except ValueError:     # the reference skips unparsable lines (:45-48)
                    continue
                if len(ids) < 2:
                    continue
                self.test_set[ids[0]] = ids[1:]
                self.n_items = max(self.n_items, max(ids[1:]))
                self.n_test += len(ids) - 1
        self.n_users += 1
        self.n_items += 1
        self.print_statistics()
        r = np.concatenate(rows) if rows else np.zeros(0, np.int64)
        c = np.concatenate(cols) if cols else np.zeros(0, np.int64)
        # R[uid, i] = 1. (dok assignment: duplicates collapse to a single 1, :66-67)
        R = sp.csr_matrix((np.ones(r.size, np.float32), (r, c)), shape=(self.n_users, self.n_items))
        R.data[:] = 1.0
        R.sum_duplicates()
        R.data[:] = 1.0
        self.R = R
        self._train_sets = {}

    def print_statistics(self):
        print("n_users=%d, n_items=%d" % (self.n_users, self.n_items))
        print("n_interactions=%d" % (self.n_train + self.n_test))
        print("n_train=%d, n_test=%d, sparsity=%.5f" % (
            self.n_train, self.n_test, (self.n_train + self.n_test) / (self.n_users * self.n_items)))

    # ---- adjacency: load_data.py:95-164 -----------------------------------------------------------
    def plain_adj(self):
        """A = [[0, R], [R^T, 0]] as float32 CSR with sorted indices (create_adj_mat, :126-139)."""
        A = sp.bmat([[None, self.R], [self.R.T, None]], format="csr", dtype=np.float32)
        A.sort_indices()
        return A

    def pre_adj(self):
        """`pre` = D^-1/2 A D^-1/2 (get_adj_mat, :112-123): float32, inf -> 0."""
        A = self.plain_adj()
        rowsum = np.asarray(A.sum(1), dtype=np.float32).flatten()
        with np.errstate(divide="ignore"):
            d_inv = np.power(rowsum, np.float32(-0.5)).astype(np.float32)
        d_inv[np.isinf(d_inv)] = 0.0
        row_of = np.repeat(np.arange(A.shape[0]), np.diff(A.indptr))
        data = (d_inv[row_of] * A.data).astype(np.float32) * d_inv[A.indices]
        return sp.csr_matrix((data.astype(np.float32), A.indices.copy(), A.indptr.copy()), shape=A.shape)

    def get_adj_mat(self):
        """Same 4-tuple as the reference (plain, norm, mean, pre); `norm`/`mean` are the
        row-normalised variants (:141-160).  Unlike the reference nothing is written to disk."""
        A = self.plain_adj()

        def row_normalised(M):
            rs = np.asarray(M.sum(1), dtype=np.float32).flatten()
            with np.errstate(divide="ignore"):
                inv = np.power(rs, np.float32(-1.0)).astype(np.float32)
            inv[np.isinf(inv)] = 0.0
            return sp.diags(inv).dot(M).tocsr()

        norm = row_normalised(A + sp.eye(A.shape[0], dtype=np.float32))
        mean = row_normalised(A)
        return A, norm, mean, self.pre_adj()

    def adj_csr(self, adj_type="pre"):
        """(rowptr, col, val) int32/int32/float32 of the adjacency the run uses
        (LightGCN.py:666-681: plain | norm | gcmc(mean) | pre | default mean+I)."""
        plain, norm, mean, pre = self.get_adj_mat()
        if adj_type == "plain":
            M = plain
        elif adj_type == "norm":
            M = norm
        elif adj_type == "gcmc":
            M = mean
        elif adj_type == "pre":
            M = pre
        else:
            M = (mean + sp.eye(mean.shape[0], dtype=np.float32)).tocsr()
        M = M.tocsr().astype(np.float32)
        M.sort_indices()
        return M.indptr.astype(np.int32), M.indices.astype(np.int32), M.data.astype(np.float32)

    # ---- sampler: load_data.py:174-212 ------------------------------------------------------------
    def _train_set(self, u):
        s = self._train_sets.get(u)
        if s is None:
            s = self._train_sets[u] = frozenset(self.train_items[u])
        return s

    def _draw(self, users, pos_lists, exclude):
        pos_items, neg_items = [], []
        for u in users:
            mine = pos_lists[u]
            pos_items.append(mine[np.random.randint(low=0, high=len(mine), size=1)[0]])
            banned = exclude(u)
            while True:
                cand = np.random.randint(low=0, high=self.n_items, size=1)[0]
                if cand not in banned:
                    neg_items.append(cand)
                    break
        return users, pos_items, neg_items

    def sample(self):
        """utility/load_data.py:174-212 through the native sampler (same `random` and
        `np.random` streams, bit-identical triples); `sample_py` is the Python restatement."""
        from . import native_sampler as ns

        if getattr(self, "_ns_train", None) is None:
            self._ns_train = ns.ListCSR(self.train_items, self.n_users)
            self._ns_pop = np.asarray(self.exist_users, np.int32)
        u, p, n = ns.sample_lgcn(self._ns_pop, self.n_users, self.n_items, self._ns_train,
                                 self._ns_train, self.batch_size)
        return u.tolist(), p.tolist(), n.tolist()

    def sample_epoch(self, n_batches, out=None):
        """`n_batches` consecutive sample() calls -> int32 [n_batches, 3, B]."""
        from . import native_sampler as ns

        if getattr(self, "_ns_train", None) is None:
            self._ns_train = ns.ListCSR(self.train_items, self.n_users)
            self._ns_pop = np.asarray(self.exist_users, np.int32)
        return ns.sample_lgcn_epoch(self._ns_pop, self.n_users, self.n_items, self._ns_train,
                                    self._ns_train, self.batch_size, n_batches, out)

    def sample_py(self):
        B = self.batch_size
        if B <= self.n_users:
            users = random.sample(self.exist_users, B)
        else:
            users = [random.choice(self.exist_users) for _ in range(B)]
        return self._draw(users, self.train_items, self._train_set)

    def sample_test(self):
        """load_data.py:214-257 (the test-loss pass of LightGCN.py:799-819) through the native
        sampler.  The reference passes dict_keys to random.sample -- a TypeError on Python >=
        3.11; list(keys) draws the same stream."""
        from . import native_sampler as ns

        self._build_test_csr()
        u, p, n = ns.sample_lgcn(self._ns_test_pop, self.n_users, self.n_items, self._ns_test,
                                 self._ns_test_ban, self.batch_size)
        return u.tolist(), p.tolist(), n.tolist()

    def _build_test_csr(self):
        from . import native_sampler as ns

        if getattr(self, "_ns_test", None) is None:
            self._ns_test = ns.ListCSR(self.test_set, self.n_users)
            both = {u: list(set(v) | set(self.train_items.get(u, ()))) for u, v in self.test_set.items()}
            self._ns_test_ban = ns.ListCSR(both, self.n_users)
            self._ns_test_pop = np.asarray(list(self.test_set.keys()), np.int32)

    def sample_test_epoch(self, n_batches, out=None):
        """`n_batches` consecutive sample_test() calls -> int32 [n_batches, 3, B]."""
        from . import native_sampler as ns

        self._build_test_csr()
        return ns.sample_lgcn_epoch(self._ns_test_pop, self.n_users, self.n_items, self._ns_test,
                                    self._ns_test_ban, self.batch_size, n_batches, out)

    def sample_test_py(self):
        B = self.batch_size
        keys = list(self.test_set.keys())
        if B <= self.n_users:
            users = random.sample(keys, B)
        else:
            users = [random.choice(keys) for _ in range(B)]

        def banned(u):
            return frozenset(self.test_set[u]) | frozenset(self.train_items.get(u, ()))

        return self._draw(users, self.test_set, banned)

    # ---- CSR views for the device-side evaluation -----------------------------------------------
    def train_csr(self, users):
        return cached_sorted_csr(self, "_train_sorted", self.train_items, users)

    def truth_csr(self, users):
        return lists_to_csr([self.test_set.get(u, []) for u in users], len(users), sort_unique=False)
