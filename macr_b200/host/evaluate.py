"""Evaluation drivers: full-catalogue counterfactual score -> train-item mask -> top-K -> metrics.

  MFEvaluator.test        macr_mf/train.py:162-311 (`test`) with `test_one_user` (:119-138),
                          `ranklist_by_sorted` (:89-104) and `get_performance` (:106-117)
  LGCNEvaluator.test      macr_lightgcn/utility/batch_test.py:26-162
  eval_score_matrix_foldout
                          macr_lightgcn/evaluator/cpp/evaluate_foldout.py:12-18 (drop-in)

Default (`eval_mode="fused"`): the [B_u, I] score matrix never exists -- one fused kernel
scores, masks and keeps K ids per user, and only those ids (MF) or the [T, 5K] metric curves
(LightGCN) come back to the host.  `eval_mode="matrix"` runs the literal sequence of the
reference (fetch `rubi_ratings_both`, mask, rank) through the same C ABI and exists to show the
two agree.  The O(B*I) Python side loops of the reference that compute values nobody reads
(train.py:261-277, batch_test.py:94-115 incl. the `Lightgcn_macr.txt` dump) are not reproduced.
"""
import numpy as np
import torch

from .. import ops
from .data_mf import lists_to_csr

_FETCH_OF = {"rubi_both": "rubi_ratings_both", "rubiboth": "rubi_ratings_both", "o": "batch_ratings",
             "normal": "batch_ratings",
             "rubi_c": "rubi_ratings",    # MF `--train rubibce --test rubi` (train.py:241,551)
             "rubi1": "rubi_ratings1"}    # LightGCN `--test rubi1` (batch_test.py:66-74)
_HEAD_OF = {"rubi_ratings_both": "both", "rubi_ratings": "item", "rubi_ratings1": "item", "batch_ratings": "plain"}


def _batches(seq, size):
    # the reference walks n // size + 1 batches (the last one may be empty); empty ones are skipped
    for s in range(0, len(seq), size):
        yield seq[s:s + size]


# ------------------------------------------------------------------------------------------------
# MF: precision / recall / ndcg / hit_ratio @Ks in float64 (train.py:32-117)
# ------------------------------------------------------------------------------------------------
def mf_metrics_from_hits(hits, n_pos, Ks, n_ranked=None):
    """hits [T, Kmax] 0/1 in rank order, n_pos [T] = len(user_pos_test) -> dict of sums over users
    of the per-user metrics (train.py:106-117; dcg at :44-57, ndcg_at_k method 1 at :60-75).
    n_ranked [T]: length of the user's rank list (< Kmax when fewer unmasked items exist);
    `precision_at_k` is `np.mean(r[:k])` (train.py:32-36), i.e. divides by min(k, len(r))."""
    hits = np.asarray(hits, dtype=np.float64)
    n_pos = np.asarray(n_pos, dtype=np.float64)
    n_ranked = np.full(hits.shape[0], hits.shape[1], np.float64) if n_ranked is None else \
        np.asarray(n_ranked, dtype=np.float64)
    out = {k: np.zeros(len(Ks)) for k in ("precision", "recall", "ndcg", "hit_ratio")}
    Kmax = hits.shape[1]
    discount = 1.0 / np.log2(np.arange(2, Kmax + 2))
    for j, K in enumerate(Ks):
        h = hits[:, :K]
        got = h.sum(1)
        out["precision"][j] = np.sum(got / np.maximum(np.minimum(n_ranked, K), 1.0))
        out["recall"][j] = np.sum(got / n_pos)
        dcg = (h * discount[:K]).sum(1)
        ideal = np.minimum(n_pos, K).astype(np.int64)
        idcg = np.concatenate([[0.0], np.cumsum(discount[:K])])[ideal]
        out["ndcg"][j] = np.sum(np.where(idcg > 0, dcg / np.where(idcg > 0, idcg, 1.0), 0.0))
        out["hit_ratio"][j] = np.sum(got > 0)
    return out


def _truth_keys(truth_lists):
    """ragged truth lists -> (lens int64 [T], flat item ids int64) for `_hits`"""
    lens = np.fromiter((len(t) for t in truth_lists), np.int64, len(truth_lists))
    flat = np.fromiter((i for t in truth_lists for i in t), np.int64, int(lens.sum()))
    return lens, flat


def _hits(topk_ids, truth_lists, keys=None):
    """r[t, k] = 1.0 if the k-th ranked item of user t is in the user's test list (train.py:94-99,
    `if i in user_pos_test`), for all users at once: one sorted membership test over (row, item)
    keys instead of one `np.isin` per user (0.09 s per evaluation on addressa, 1.2 s on gowalla)."""
    ids = np.asarray(topk_ids).astype(np.int64)
    T = ids.shape[0]
    lens, flat = _truth_keys(truth_lists) if keys is None else keys
    stride = max(int(ids.max(initial=0)), int(flat.max(initial=0))) + 2  # ids of -1 (padding) hit nothing
    rows = np.arange(T, dtype=np.int64)
    return np.isin(rows[:, None] * stride + ids, np.repeat(rows, lens) * stride + flat).astype(np.float64)


class MFEvaluator:
    def __init__(self, data, Ks, batch_size, eval_mode="fused"):
        self.data, self.Ks, self.batch_size, self.eval_mode = data, list(Ks), batch_size, eval_mode
        self._per_batch = {}  # (valid_set, users of the batch) -> train CSR, truth lists as arrays

    def _batch_lists(self, user_batch, valid_set):
        """train-item mask CSR and test lists of one user batch; every evaluation of a run walks the
        same batches (train.py:174-180), so the ragged Python lists are flattened once"""
        key = (valid_set, tuple(user_batch))
        got = self._per_batch.get(key)
        if got is None:
            truth_of = self.data.test_user_list if valid_set == "test" else self.data.valid_user_list
            truth = [truth_of[u] for u in user_batch]
            got = self._per_batch[key] = (self.data.train_csr(user_batch), truth, _truth_keys(truth),
                                          np.array([len(t) for t in truth], np.float64))
        return got

    def test(self, sess, model, test_users, batch_test_flag=False, model_type="o", valid_set="test"):
        if model_type not in _FETCH_OF:
            raise NotImplementedError(f"model_type {model_type!r} is outside the MACR hot path")
        Kmax = max(self.Ks)
        sums = {k: np.zeros(len(self.Ks)) for k in ("precision", "recall", "ndcg", "hit_ratio")}
        n_test_users, count = len(test_users), 0
        head = _HEAD_OF[_FETCH_OF[model_type]]
        prep = {}  # item gates + tensor-core item operands: prepared by the first batch, shared by the rest
        for user_batch in _batches(test_users, self.batch_size):
            (mrp, mcol), truth, keys, n_pos = self._batch_lists(user_batch, valid_set)  # all_items - train_items, train.py:132-133
            if self.eval_mode == "fused":
                ids, _ = model.topk(user_batch, Kmax, mrp, mcol, head=head, prep=prep)
                ids = ids.cpu().numpy()
            else:  # literal: fetch the matrix, mask, rank on the host (heapq.nlargest order)
                rate = sess.run(getattr(model, _FETCH_OF[model_type]),
                                {model.users: user_batch, model.pos_items: range(self.data.n_items)})
                ids = host_topk(rate, mrp, mcol, Kmax)
            part = mf_metrics_from_hits(_hits(ids, truth, keys), n_pos, self.Ks,
                                        n_ranked=(np.asarray(ids) >= 0).sum(1))
            for k in sums:
                sums[k] += part[k]
            count += len(user_batch)
        assert count == n_test_users  # train.py:309
        return {k: v / n_test_users for k, v in sums.items()}


def host_topk(rate, mask_rowptr, mask_col, K):
    """Rank a host score matrix: masked items removed, score descending, lower id first."""
    rate = np.array(rate, dtype=np.float32, copy=True)
    T, n = rate.shape
    out = np.full((T, K), -1, np.int32)
    for t in range(T):
        row = rate[t]
        row[mask_col[mask_rowptr[t]:mask_rowptr[t + 1]]] = -np.inf
        order = np.lexsort((np.arange(n), -row))[:K]
        order = order[~np.isneginf(row[order])]  # fewer than K unmasked items: pad with -1
        out[t, :len(order)] = order
    return out


# ------------------------------------------------------------------------------------------------
# LightGCN: fold-out metric curves (evaluate_foldout.h:16-195) + the hr rewrite of batch_test.py
# ------------------------------------------------------------------------------------------------
class LGCNEvaluator:
    def __init__(self, data, batch_size, eval_mode="fused"):
        self.data, self.batch_size, self.eval_mode = data, batch_size, eval_mode

    def test(self, sess, model, users_to_test, drop_flag=False, train_set_flag=0, method="normal"):
        if method not in _FETCH_OF:
            raise NotImplementedError(f"method {method!r} is outside the MACR hot path")
        top_show = np.sort(model.Ks)
        max_top = int(max(top_show))
        head = _HEAD_OF[_FETCH_OF[method]]
        all_result, count = [], 0
        prep = {}  # item gates + tensor-core item operands of this evaluation
        for user_batch in _batches(users_to_test, self.batch_size):
            mrp, mcol = self.data.train_csr(user_batch) if train_set_flag == 0 else \
                (np.zeros(len(user_batch) + 1, np.int32), np.zeros(0, np.int32))
            trp, tcol = self.data.truth_csr(user_batch)
            if self.eval_mode == "fused":
                ids, _ = model.topk(user_batch, max_top, mrp, mcol, head=head, prep=prep)
                with torch.cuda.device(model.dev):
                    res = ops.foldout_metrics(ids, model._ids(trp), model._ids(tcol)).cpu().numpy()
            else:
                rate = sess.run(getattr(model, _FETCH_OF[method]),
                                {model.users: user_batch, model.pos_items: range(self.data.n_items)})
                rate = np.array(rate, dtype=np.float32, copy=True)
                for idx in range(len(user_batch)):  # batch_test.py:124-129
                    rate[idx][mcol[mrp[idx]:mrp[idx + 1]]] = -np.inf
                res = eval_score_matrix_foldout(rate, [self.data.test_set[u] for u in user_batch],
                                                max_top, device=model.dev)
            all_result.append(res)
            count += len(res)
        assert count == len(users_to_test)  # batch_test.py:139
        return lgcn_result_from_curves(np.concatenate(all_result, axis=0), top_show)


def lgcn_result_from_curves(all_result, top_show):
    """batch_test.py:141-161: per-user fold-out curves float32 [n, 5*K] -> the reported dict.  Slot 2
    (ap in the C++ evaluator) is overwritten with `recall@k != 0`, i.e. the hit ratio; then the
    user mean, taken at the cut-offs `top_show` (= sorted Ks)."""
    all_result = np.array(all_result, dtype=np.float32, copy=True)
    top_show = np.asarray(top_show)
    K = all_result.shape[1] // 5
    all_result[:, 2 * K:3 * K] = (all_result[:, K:2 * K] != 0).astype(np.float32)  # :143-149
    final = np.mean(all_result, axis=0).reshape(5, K)[:, top_show - 1]            # :151-157
    return {"hr": final[2].copy(), "recall": final[1].copy(), "ndcg": final[3].copy()}


def eval_score_matrix_foldout(score_matrix, test_items, top_k=20, thread_num=None, device="cuda:0"):
    """Drop-in for evaluator/cpp/evaluate_foldout.py:12-18: float32 [rows, 5*top_k] laid out
    [precision | recall | ap | ndcg | mrr] x top_k.  `thread_num` is accepted and ignored (the
    reference builds a 5 x cpu_count thread pool per call; here one kernel ranks every row).
    Ties: lower item id first (std::partial_sort_copy leaves tie order unspecified)."""
    if len(score_matrix) != len(test_items):
        raise ValueError("The lengths of score_matrix and test_items are not equal.")
    dev = torch.device(device)
    with torch.cuda.device(dev):
        scores = torch.as_tensor(np.ascontiguousarray(score_matrix, dtype=np.float32)).to(dev)
        if scores.shape[0] == 0:
            return np.zeros((0, 5 * top_k), np.float32)
        rk = ops.topk_rows(scores, top_k)
        trp, tcol = lists_to_csr(list(test_items), len(test_items), sort_unique=False)
        tcol_d = torch.as_tensor(tcol if len(tcol) else np.zeros(1, np.int32)).to(dev)
        return ops.foldout_metrics(rk, torch.as_tensor(trp).to(dev), tcol_d).cpu().numpy()
