"""Shared synthetic-input builders for the parity tests (seeded, shape-matched to SURVEY 8d)."""
import numpy as np


def xavier(rng, rows, d=64):
    lim = np.sqrt(6.0 / (rows + d))
    return rng.uniform(-lim, lim, size=(rows, d)).astype(np.float32)


def make_model(seed, n_users, n_items, d=64, scale=1.0):
    """Tables as the reference initialises them: Xavier-uniform (macr_mf/model.py:107-115),
    w / w_user Xavier over [64,1] (:59-60). `scale` > 1 gives non-degenerate scores."""
    rng = np.random.RandomState(seed)
    U = xavier(rng, n_users, d) * scale
    I = xavier(rng, n_items, d) * scale
    lim = np.sqrt(6.0 / (d + 1))
    w = rng.uniform(-lim, lim, size=d).astype(np.float32)
    wu = rng.uniform(-lim, lim, size=d).astype(np.float32)
    return U.astype(np.float32), I.astype(np.float32), w, wu


def make_batch(rng, n_users, n_items, B, zipf=True):
    """users without replacement when B <= n_users (like random.sample), popular positives."""
    if B <= n_users:
        u = rng.permutation(n_users)[:B]
    else:
        u = rng.randint(0, n_users, B)
    if zipf:
        ranks = np.arange(1, n_items + 1, dtype=np.float64)
        pz = (1.0 / ranks) / np.sum(1.0 / ranks)
        p = rng.choice(n_items, size=B, p=pz)
    else:
        p = rng.randint(0, n_items, B)
    n = rng.randint(0, n_items, B)
    return u.astype(np.int32), p.astype(np.int32), n.astype(np.int32)


def make_interactions(seed, n_users, n_items, avg_deg):
    """Random bipartite train lists (sorted unique item ids per user), at least one per user."""
    rng = np.random.RandomState(seed)
    lists = []
    for u in range(n_users):
        k = max(1, int(rng.poisson(avg_deg)))
        k = min(k, n_items)
        lists.append(np.sort(rng.choice(n_items, size=k, replace=False)).astype(np.int32))
    return lists


def lists_to_csr(lists):
    rowptr = np.zeros(len(lists) + 1, np.int32)
    for i, l in enumerate(lists):
        rowptr[i + 1] = rowptr[i] + len(l)
    col = np.concatenate([np.asarray(l, np.int32) for l in lists]) if lists else np.zeros(0, np.int32)
    return rowptr, np.ascontiguousarray(col, dtype=np.int32)


def norm_adj_csr(lists, n_users, n_items):
    """D^-1/2 A D^-1/2 of the bipartite graph as float32 CSR (utility/load_data.py:112-124)."""
    import scipy.sparse as sp

    rows = np.concatenate([np.full(len(l), u, np.int64) for u, l in enumerate(lists)])
    cols = np.concatenate(lists).astype(np.int64)
    R = sp.csr_matrix((np.ones(len(rows), np.float32), (rows, cols)), shape=(n_users, n_items))
    A = sp.bmat([[None, R], [R.T, None]], format="csr", dtype=np.float32)
    rowsum = np.array(A.sum(1)).flatten()
    with np.errstate(divide="ignore"):
        d_inv = np.power(rowsum, -0.5)
    d_inv[np.isinf(d_inv)] = 0.0
    Dm = sp.diags(d_inv)
    N = Dm.dot(A).dot(Dm).tocsr().astype(np.float32)
    N.sort_indices()
    return N.indptr.astype(np.int32), N.indices.astype(np.int32), N.data.astype(np.float32)
