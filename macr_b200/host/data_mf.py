"""MF data set + triple sampler -- host mirror of macr_mf/load_data.py (reference lines cited).

File format (`data/<dataset>/{train,test,valid}.txt`): one line per user, `uid iid iid ...`.
The sampler consumes Python's `random` stream call for call like the reference's
`Data.sample()` (load_data.py:543-566), so a run seeded with `random.seed(12345)` draws the
same (user, pos, neg) triples bit for bit; only the membership test is done on a set instead
of a list (same truth value, O(1)).
"""
import collections
import os
import random

import numpy as np


def read_interactions(path):
    """-> list of (uid, [items]) in file order; lines with no item are dropped
    (load_data.py:51-60)."""
    out = []
    with open(path) as f:
        for raw in f:
            tok = raw.split()
            if len(tok) < 2:
                continue
            ids = [int(t) for t in tok]
            out.append((ids[0], ids[1:]))
    return out


def lists_to_csr(lists, n_rows, sort_unique=True):
    """dict/list of per-row id lists -> (rowptr int32 [n_rows+1], col int32)."""
    rowptr = np.zeros(n_rows + 1, np.int64)
    cols = []
    for r in range(n_rows):
        items = lists.get(r, ()) if isinstance(lists, dict) else lists[r]
        a = np.asarray(items, np.int32)
        if sort_unique:
            a = np.unique(a)
        cols.append(a)
        rowptr[r + 1] = rowptr[r] + a.size
    col = np.concatenate(cols) if cols else np.zeros(0, np.int32)
    return rowptr.astype(np.int32), np.ascontiguousarray(col, np.int32)


def cached_sorted_csr(owner, attr, lists, users):
    """CSR of `np.unique(lists[u])` for `users`; the per-user sorted arrays are computed once per
    Data object (every evaluation re-reads the same train lists)."""
    cache = owner.__dict__.setdefault(attr, {})
    cols, rowptr = [], np.zeros(len(users) + 1, np.int64)
    for r, u in enumerate(users):
        a = cache.get(u)
        if a is None:
            a = cache[u] = np.unique(np.asarray(lists.get(u, ()), np.int32))
        cols.append(a)
        rowptr[r + 1] = rowptr[r] + a.size
    col = np.concatenate(cols) if cols else np.zeros(0, np.int32)
    return rowptr.astype(np.int32), np.ascontiguousarray(col, np.int32)


class Data:
    """`Data(args)` of macr_mf/load_data.py:504-541 restricted to the hot path's configuration
    (`--data_type ori --source normal --model mf`, load_data.py:26-118)."""

    def __init__(self, args):
        if getattr(args, "data_type", "ori") != "ori" or getattr(args, "source", "normal") != "normal" \
                or getattr(args, "model", "mf") not in ("mf", "biasmf"):
            raise NotImplementedError(
                "only --data_type ori --source normal --model mf is on the MACR hot path "
                "(imbalance / dice / CausalE loaders are out of scope, DESIGN.md section 8)")
        # the reference ignores --data_path for this loader and reads ./data/<dataset>/
        # (load_data.py:27); fall back to --data_path when that directory does not exist
        self.path = "./data/{}/".format(args.dataset)
        if not os.path.isdir(self.path):
            self.path = os.path.join(args.data_path, args.dataset) + "/"
        self.batch_size = args.batch_size
        self.n_users = self.n_items = 0
        self.n_train = self.n_test = self.n_valid = 0
        self.train_user_list = collections.defaultdict(list)
        self.test_user_list = collections.defaultdict(list)
        self.valid_user_list = collections.defaultdict(list)
        self.train_item_list = collections.defaultdict(list)
        self.test_item_list = collections.defaultdict(list)
        self.valid_item_list = collections.defaultdict(list)
        self.valid_items = set()

        self.n_train = self._load("train.txt", self.train_user_list, self.train_item_list)
        if args.valid_set == "valid":  # load_data.py:67-85
            self.n_valid = self._load("valid.txt", self.valid_user_list, self.valid_item_list)
            for items in self.valid_user_list.values():
                self.valid_items.update(items)
        if args.valid_set == "test":   # load_data.py:86-102
            self.n_test = self._load("test.txt", self.test_user_list, self.test_item_list)
        self.test_users = set(self.test_user_list.keys())
        self.valid_users = list(self.valid_user_list.keys())
        self.valid_items = list(self.valid_items)
        self.n_users += 1              # ids are 0-based: count = max id + 1 (load_data.py:103-104)
        self.n_items += 1
        self.users = list(range(self.n_users))
        self.items = list(range(self.n_items))
        self._train_sets = {}
        print("n_items:", self.n_items, "n_users:", self.n_users)
        print("sparsity:", 1.0 * self.n_train / self.n_items / self.n_users)

    def _load(self, name, user_lists, item_lists):
        total = 0
        for uid, items in read_interactions(self.path + name):
            user_lists[uid] = items  # a repeated uid keeps its last line, like the reference
            for it in items:
                item_lists[it].append(uid)
            self.n_users = max(self.n_users, uid)
            self.n_items = max(self.n_items, max(items))
            total += len(items)
        return total

    # ---- sampler: load_data.py:543-566 --------------------------------------------------------
    def _train_set(self, user):
        s = self._train_sets.get(user)
        if s is None:
            s = self._train_sets[user] = frozenset(self.train_user_list[user])
        return s

    def sample(self):
        """load_data.py:543-566 through the native sampler (same `random` stream, bit-identical
        triples, ~100x faster); `sample_py` is the line-by-line Python restatement."""
        from . import native_sampler as ns

        if getattr(self, "_ns_csr", None) is None:
            self._ns_csr = ns.ListCSR(self.train_user_list, self.n_users)
            self._ns_pop = np.asarray(self.users, np.int32)
        u, p, n = ns.sample_mf(self._ns_pop, self.n_users, self.n_items, self._ns_csr,
                               self.batch_size)
        return u.tolist(), p.tolist(), n.tolist()

    def sample_epoch(self, n_batches, out=None):
        """`n_batches` consecutive sample() calls -> int32 [n_batches, 3, B] (what
        `MFTrainer.run_host` consumes): the epoch loop of train.py:470-499 without per-step
        interpreter work."""
        from . import native_sampler as ns

        if getattr(self, "_ns_csr", None) is None:
            self._ns_csr = ns.ListCSR(self.train_user_list, self.n_users)
            self._ns_pop = np.asarray(self.users, np.int32)
        return ns.sample_mf_epoch(self._ns_pop, self.n_users, self.n_items, self._ns_csr,
                                  self.batch_size, n_batches, out)

    def sample_py(self):
        B = self.batch_size
        if B <= self.n_users:
            users = random.sample(self.users, B)
        else:
            users = [random.choice(self.users) for _ in range(B)]
        pos_items, neg_items = [], []
        for user in users:
            mine = self.train_user_list[user]
            pos_items.append(random.choice(mine) if mine else 0)
            seen = self._train_set(user)
            while True:
                cand = random.choice(self.items)
                if cand not in seen:
                    neg_items.append(cand)
                    break
        return users, pos_items, neg_items

    # ---- CSR views for the device-side evaluation -----------------------------------------------
    def train_csr(self, users):
        """train items of `users` (sorted unique global item ids): the top-K exclusion mask,
        i.e. `all_items - set(training_items)` of train.py:132-133."""
        return cached_sorted_csr(self, "_train_sorted", self.train_user_list, users)

    def truth_csr(self, users, valid_set="test"):
        src = self.test_user_list if valid_set == "test" else self.valid_user_list
        return lists_to_csr([src.get(u, []) for u in users], len(users), sort_unique=False)
