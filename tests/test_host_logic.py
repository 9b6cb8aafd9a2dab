"""Host-side logic that needs no GPU: metric arithmetic against the oracle's restatement of
train.py:32-117, host ranking, the session dispatch rules, checkpoint RNG round trip."""
import random
import types

import numpy as np
import pytest

from macr_b200.host import evaluate, session
from oracle import mf_metrics


def test_mf_metrics_from_hits_matches_reference_arithmetic():
    rng = np.random.RandomState(0)
    T, n_items, Ks = 40, 90, [5, 10, 20]
    topk = np.stack([rng.permutation(n_items)[:20] for _ in range(T)])
    truth = [list(rng.choice(n_items, size=int(rng.randint(1, 30)), replace=False)) for _ in range(T)]
    want = mf_metrics.evaluate(topk, truth, Ks)
    got = evaluate.mf_metrics_from_hits(evaluate._hits(topk, truth), [len(t) for t in truth], Ks)
    for k in want:
        np.testing.assert_allclose(got[k] / T, want[k], rtol=1e-12, atol=1e-15, err_msg=k)


def test_mf_precision_divides_by_the_rank_list_length_for_short_lists():
    """train.py:32-36: precision_at_k = np.mean(r[:k]); a user with fewer than K unmasked items has a
    shorter r (padding ids -1 are not ranks), so the divisor is min(k, len(r))."""
    ids = np.array([[4, 7, 9, -1, -1], [1, 2, 3, 5, 6]], np.int32)
    truth = [[7, 9], [1]]
    hits = evaluate._hits(ids, truth)
    got = evaluate.mf_metrics_from_hits(hits, [2, 1], [2, 5], n_ranked=(ids >= 0).sum(1))
    want_p = [np.mean([0, 1]) + np.mean([1, 0]), np.mean([0, 1, 1]) + np.mean([1, 0, 0, 0, 0])]
    np.testing.assert_allclose(got["precision"], want_p, rtol=1e-15)
    np.testing.assert_allclose(got["recall"], [0.5 + 1.0, 1.0 + 1.0], rtol=1e-15)


def test_mf_evaluation_matches_the_references_own_functions(oracle):
    """tests/golden/mf_metrics.npz holds outputs of macr_mf/train.py's OWN ranklist_by_sorted +
    get_performance (:32-117), compiled from the reference file by tests/golden/make_golden.py.
    Pins, on the same score matrix: the ranking rule (train items removed, score descending,
    heapq.nlargest keeps the lower id on a tie), the oracle's metric restatement and the product's
    host-side metric arithmetic."""
    import os

    g = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "mf_metrics.npz"))
    rating, Ks = g["rating"], g["Ks"].tolist()
    split = lambda flat, lens: np.split(flat, np.cumsum(lens)[:-1])
    train, test = split(g["train"], g["train_len"]), [t.tolist() for t in split(g["test"], g["test_len"])]
    mrp = np.concatenate([[0], np.cumsum(g["train_len"])]).astype(np.int32)
    ids = evaluate.host_topk(rating, mrp, g["train"], max(Ks))                 # product: literal eval path
    np.testing.assert_array_equal(evaluate._hits(ids, test).astype(np.int8), g["hits"])
    assert rating[3, 10] == rating[3, 200] and ids[3, 0] == 10 and 200 in ids[3, :2]  # the tie: lower id first
    masked = rating.copy()
    for u, t in enumerate(train):
        masked[u, t] = -np.inf
    np.testing.assert_array_equal(oracle.topk_rows(masked, max(Ks)), ids)      # oracle ranking rule
    want = {k: g["res_" + k] for k in ("precision", "recall", "ndcg", "hit_ratio")}
    got_o = mf_metrics.evaluate(ids, test, Ks)                                 # oracle restatement
    got_p = evaluate.mf_metrics_from_hits(evaluate._hits(ids, test), [len(t) for t in test], Ks)
    for k in want:
        np.testing.assert_allclose(got_o[k], want[k], rtol=1e-12, atol=1e-15, err_msg=k)
        np.testing.assert_allclose(got_p[k] / len(test), want[k], rtol=1e-12, atol=1e-15, err_msg=k)


def test_lgcn_evaluation_matches_the_references_own_test_function(oracle):
    """tests/golden/lgcn_test.npz holds the outputs of macr_lightgcn/utility/batch_test.py's OWN
    test() (:26-162), run by tests/golden/make_golden.py with the reference's own Data object and
    C++ evaluator on seeded score matrices of the `tiny` data set.  Pins the whole evaluation
    chain around the score matrix: train items := -inf, top-K, fold-out curves (oracle), the
    hit-ratio rewrite and the user mean at the cut-offs (product: lgcn_result_from_curves), with
    the product's own loader supplying the train / test lists."""
    import os
    import types

    from macr_b200.host.data_lgcn import Data

    gold = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
    g = np.load(os.path.join(gold, "lgcn_test.npz"))
    data = Data(os.path.join(gold, "tiny"), 16, types.SimpleNamespace(valid_set="test"))
    users = g["users_to_test"].tolist()
    assert users == list(data.test_set.keys())  # same users, same order as the reference's loader
    Ks = np.sort(g["Ks"])
    K = int(Ks.max())
    mrp, mcol = data.train_csr(users)
    trp, tcol = data.truth_csr(users)
    for method, fetch in (("normal", "batch_ratings"), ("rubiboth", "rubi_ratings_both")):
        rate = g[fetch][users].copy()
        for i in range(len(users)):  # batch_test.py:124-129
            rate[i, mcol[mrp[i]:mrp[i + 1]]] = -np.inf
        curves = oracle.foldout_metrics(oracle.topk_rows(rate, K), trp, tcol)
        got = evaluate.lgcn_result_from_curves(curves, Ks)
        for k in ("hr", "recall", "ndcg"):
            np.testing.assert_allclose(got[k], g[f"{method}_{k}"], rtol=1e-6, atol=1e-7, err_msg=f"{method} {k}")
    assert got["hr"].dtype == np.float32 and got["hr"].shape == (2,)


def test_host_topk_masks_and_breaks_ties_by_id():
    rate = np.array([[1.0, 3.0, 3.0, 2.0, 0.5], [5.0, 4.0, 3.0, 2.0, 1.0]], np.float32)
    mrp, mcol = np.array([0, 1, 4], np.int32), np.array([1, 0, 1, 2], np.int32)
    ids = evaluate.host_topk(rate, mrp, mcol, 3)
    assert ids[0].tolist() == [2, 3, 0]
    assert ids[1].tolist() == [3, 4, -1]  # only two unmasked items left


class _FakeModel:
    def __init__(self):
        self.users = session.Placeholder(self, "users")
        self.loss = session.Fetch(self, "loss")
        self.opt = session.Fetch(self, "opt")
        self.other = session.Unsupported(self, "opt_bce")
        self.calls = []

    def _run(self, names, feeds):
        self.calls.append((names, feeds))
        return [None if n == "opt" else 1.5 for n in names]


def test_session_dispatch():
    m, sess = _FakeModel(), session.Session()
    assert sess.run(session.global_variables_initializer()) is None
    out = sess.run([m.opt, m.loss], feed_dict={m.users: [1, 2]})
    assert out == [None, 1.5]
    assert m.calls[-1] == (["opt", "loss"], {"users": [1, 2]})
    assert sess.run(m.loss, {m.users: [3]}) == 1.5
    with pytest.raises(NotImplementedError):
        sess.run([m.other], {m.users: [1]})
    with pytest.raises(ValueError):
        sess.run([m.loss, _FakeModel().loss])


def test_checkpoint_restores_rng_streams(tmp_path):
    from macr_b200.host import checkpoint

    class M:
        def __init__(self):
            self.sd = {"U": np.arange(6, dtype=np.float32).reshape(3, 2), "steps_done": np.int64(7)}

        def state_dict(self):
            return dict(self.sd)

        def load_state_dict(self, sd):
            self.loaded = sd

    random.seed(5)
    np.random.seed(5)
    random.random(), np.random.rand()
    m = M()
    path = str(tmp_path / "ck" / "3_ckpt.npz")
    checkpoint.save(path, m, {"epoch": 3})
    want = (random.random(), np.random.rand())
    random.seed(99)
    np.random.seed(99)
    extra = checkpoint.load(path, m)
    assert (random.random(), np.random.rand()) == want
    assert int(extra["epoch"]) == 3 and int(m.loaded["steps_done"]) == 7
    np.testing.assert_array_equal(m.loaded["U"], m.sd["U"])


def test_drivers_reject_modes_outside_the_hot_path():
    from macr_b200.cli import lightgcn, train_mf

    with pytest.raises(SystemExit):
        train_mf.main(["--dataset", "tiny", "--train", "rubi"])  # BPR two-branch variant: not implemented
    with pytest.raises(SystemExit):
        lightgcn.main(["--dataset", "tiny", "--loss", "bpr"])


def test_mf_evaluator_walks_user_batches_like_the_reference_and_reuses_its_lists():
    """MFEvaluator.test (train.py:162-311) with a stand-in model whose `topk` ranks a fixed score
    matrix on the host: the batch walk, the train-item mask, the vectorised hit test and the metric
    sums equal the oracle's restatement evaluated over all users at once -- on the first call and
    on the second (per-batch lists cached), for the test and the valid split, padding ids included."""
    import torch

    rng = np.random.RandomState(3)
    n_users, n_items, Ks = 75, 40, [5, 20]
    lists = lambda lo, hi: {u: rng.choice(n_items, size=rng.randint(lo, hi), replace=False).tolist()
                            for u in range(n_users)}
    data = types.SimpleNamespace(n_items=n_items, train_user_list=lists(1, 30), test_user_list=lists(1, 8),
                                 valid_user_list=lists(1, 5))
    data.train_user_list[4] = list(range(n_items - 7))  # fewer than K unmasked items: ids padded with -1
    calls = []

    def train_csr(users):
        calls.append(len(users))
        rows = [np.unique(np.asarray(data.train_user_list[u], np.int32)) for u in users]
        rp = np.zeros(len(users) + 1, np.int32)
        rp[1:] = np.cumsum([len(r) for r in rows])
        return rp, np.concatenate(rows).astype(np.int32)

    data.train_csr = train_csr
    rating = rng.rand(n_users, n_items).astype(np.float32)

    class Model:
        def topk(self, users, K, mrp, mcol, head="both", prep=None):
            assert head == "both" and isinstance(prep, dict)
            ids = evaluate.host_topk(rating[np.asarray(users)], mrp, mcol, K)
            return torch.from_numpy(ids), None

    users = list(range(n_users))
    ev = evaluate.MFEvaluator(data, Ks, 32, "fused")
    for valid_set, truth_of in (("test", data.test_user_list), ("valid", data.valid_user_list), ("test", data.test_user_list)):
        got = ev.test(None, Model(), users, model_type="rubi_both", valid_set=valid_set)
        mrp, mcol = train_csr(users)
        ids = evaluate.host_topk(rating, mrp, mcol, max(Ks))
        assert (ids[4] == -1).sum() == max(Ks) - 7
        truth = [truth_of[u] for u in users]
        want = evaluate.mf_metrics_from_hits(evaluate._hits(ids, truth), [len(t) for t in truth], Ks,
                                             n_ranked=(ids >= 0).sum(1))
        ref = mf_metrics.evaluate(ids, truth, Ks)
        for k in want:
            np.testing.assert_allclose(got[k], want[k] / n_users, rtol=1e-12, err_msg=k)
            if k != "precision":  # the oracle's precision divides by K (no padded rows in its contract)
                np.testing.assert_allclose(got[k], ref[k], rtol=1e-12, err_msg=k)
    # three batches per walk, flattened once per split: train_csr ran for 2 splits x 3 batches + 3 checks
    assert calls.count(32) == 4 and calls.count(11) == 2
