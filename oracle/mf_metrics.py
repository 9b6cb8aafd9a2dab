"""numpy restatement of the MF evaluation arithmetic, macr_mf/train.py:32-117,286-290.
TEST INFRASTRUCTURE ONLY.  (`np.asfarray` of the reference is gone in numpy 2; `np.asarray(..,
float)` is its definition.)  Pinned against the reference's own functions run here:
tests/golden/mf_metrics.npz (tests/golden/make_golden.py:run_mf_metrics), tests/test_host_logic.py."""
import numpy as np


def get_performance(user_pos_test, r, Ks):
    """train.py:106-117 for one user; r = hit list of the top max(Ks) items."""
    precision, recall, ndcg, hit = [], [], [], []
    n_pos = len(user_pos_test)
    for K in Ks:
        rk = np.asarray(r, dtype=float)[:K]
        precision.append(np.mean(np.asarray(r)[:K]))                  # :32-42
        recall.append(np.sum(rk) / n_pos)                             # :77-79
        tp = 1.0 / np.log2(np.arange(2, K + 2))                       # :69
        dcg_max = tp[: min(n_pos, K)].sum()                           # :70
        dcg = np.sum(rk / np.log2(np.arange(2, rk.size + 2))) if rk.size else 0.0  # :56
        ndcg.append(dcg / dcg_max if dcg_max else 0.0)                # :71-73
        hit.append(1.0 if np.sum(np.asarray(r)[:K]) > 0 else 0.0)     # :82-87
    return {"recall": np.array(recall), "precision": np.array(precision),
            "ndcg": np.array(ndcg), "hit_ratio": np.array(hit)}


def evaluate(topk_ids, test_lists, Ks):
    """train.py:286-290: sum over users of per-user metrics / n_test_users."""
    n = len(test_lists)
    res = {k: np.zeros(len(Ks)) for k in ("precision", "recall", "ndcg", "hit_ratio")}
    for ids, pos in zip(topk_ids, test_lists):
        pos_set = set(pos)
        r = [1 if int(i) in pos_set else 0 for i in ids]             # :97-102
        re = get_performance(pos, r, Ks)
        for k in res:
            res[k] += re[k] / n
    return res
