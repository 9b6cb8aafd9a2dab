"""Pins the C oracle (oracle/macr_oracle.c) on the CPU:
  * losses + closed-form gradients against torch autograd of the LITERAL restatement of the
    TF graph (oracle/literal_torch.py keeps the [B]*[B,1] -> [B,B] broadcast of model.py:204);
  * whole MF / LightGCN steps (loss -> gradients -> TF-1.14 Adam, duplicates included) against
    the literal graph + TFAdam, several steps;
  * top-K + fold-out metric curves against golden outputs of the REFERENCE's own C++ evaluator
    (tests/golden/ref_evaluator.npz) and, when oracle/_ref is present, against it live;
  * MF metric arithmetic (train.py:32-117) on hand-checked cases.
"""
import os

import numpy as np
import pytest
import torch

from helpers import lists_to_csr, make_batch, make_interactions, make_model, norm_adj_csr
from oracle import literal_torch as lit
from oracle import mf_metrics

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def t64(a):
    return torch.tensor(np.asarray(a), dtype=torch.float64)


def test_broadcast_shape_is_a_grid():
    """The loss of model.py:204-211 is a mean over a [B,B] grid, not over B pairs."""
    B, d = 8, 64
    g = torch.Generator().manual_seed(0)
    u, p, n = (torch.randn(B, d, generator=g, dtype=torch.float64) for _ in range(3))
    w, wu = torch.randn(d, 1, generator=g, dtype=torch.float64), torch.randn(d, 1, generator=g, dtype=torch.float64)
    pos = torch.sum(u * p, 1)
    grid = pos * torch.sigmoid(p @ w) * torch.sigmoid(u @ wu)
    assert grid.shape == (B, B)
    i, j = 3, 5
    want = pos[j] * torch.sigmoid(p[i] @ w) * torch.sigmoid(u[i] @ wu)
    assert torch.allclose(grid[i, j], want[0])


@pytest.mark.parametrize("B,scale", [(1, 1.0), (7, 1.0), (64, 3.0), (200, 10.0)])
def test_grid_bce_matches_autograd(oracle, B, scale):
    rng = np.random.RandomState(B)
    vecs = [(rng.randn(B) * s).astype(np.float32) for s in (scale, scale, 2.0, 2.0, 2.0)]
    alpha, beta = 1e-2, 1e-3
    losses, grads = oracle.grid_bce(*vecs, alpha, beta)
    # saturated logits (scale 10): "1 - sigmoid(N) + 1e-10" only equals the reference's value in
    # the reference's own precision (fp32: 1 - sigmoid rounds to 0), so the literal graph is
    # evaluated in fp32 there; elsewhere fp64 gives the tighter check.
    dt = torch.float32 if scale >= 10 else torch.float64
    tol = 50.0 if scale >= 10 else 1.0
    yp, yn, sp, sn, su = [torch.tensor(v, dtype=dt).requires_grad_(True) for v in vecs]
    P = yp * torch.sigmoid(sp[:, None]) * torch.sigmoid(su[:, None])
    N = yn * torch.sigmoid(sn[:, None]) * torch.sigmoid(su[:, None])
    l_ori = torch.mean(-torch.log(torch.sigmoid(P) + 1e-10) - torch.log(1 - torch.sigmoid(N) + 1e-10))
    l_item = torch.mean(-torch.log(torch.sigmoid(sp) + 1e-10) - torch.log(1 - torch.sigmoid(sn) + 1e-10))
    l_user = torch.mean(-torch.log(torch.sigmoid(su) + 1e-10) - torch.log(1 - torch.sigmoid(su) + 1e-10))
    (l_ori + alpha * l_item + beta * l_user).backward()
    np.testing.assert_allclose(losses, [l_ori.item(), l_item.item(), l_user.item()], rtol=2e-5 * tol,
                               atol=1e-6)
    for got, ref in zip(grads, (yp, yn, sp, sn, su)):
        ref = ref.grad.numpy()
        np.testing.assert_allclose(got, ref, rtol=2e-4 * tol,
                                   atol=2e-6 * tol * max(1e-12, np.abs(ref).max()))
    l_only, none = oracle.grid_bce(*vecs, alpha, beta, want_grad=False)
    assert none is None
    np.testing.assert_array_equal(l_only, losses)


def _literal_mf_steps(U, I, w, wu, batches, hp_kw):
    P = [t64(U).requires_grad_(True), t64(I).requires_grad_(True),
         t64(w).reshape(-1, 1).requires_grad_(True), t64(wu).reshape(-1, 1).requires_grad_(True)]
    opt = lit.TFAdam(P, hp_kw["lr"])
    out = []
    for u, p, n in batches:
        for q in P:
            q.grad = None
        ui, pi, ni = (torch.as_tensor(x, dtype=torch.long) for x in (u, p, n))
        mf, reg, *_ = lit.bce_two_branch_both(P[0][ui], P[1][pi], P[1][ni], P[2], P[3],
                                              hp_kw["alpha"], hp_kw["beta"], hp_kw["decay"],
                                              hp_kw["batch_size"])
        (mf + reg).backward()
        opt.step([q.grad for q in P], [True, True, False, False])
        out.append(((mf + reg).item(), mf.item(), reg.item()))
    return P, opt, out


def test_mf_step_matches_literal_graph(oracle):
    """3 steps incl. duplicate users/items in a batch (segment-sum dedup) and rows that only
    move on momentum (TF's dense Adam)."""
    n_users, n_items, B = 60, 40, 96  # B > n_users: users repeat inside a batch
    U, I, w, wu = make_model(5, n_users, n_items, scale=4.0)
    hp_kw = dict(lr=1e-2, alpha=1e-2, beta=1e-3, decay=1e-3, batch_size=B)
    rng = np.random.RandomState(6)
    batches = [make_batch(rng, n_users, n_items, B) for _ in range(3)]
    st = oracle.MFState(U, I, w, wu)
    hp = oracle.HParams.make(**hp_kw)
    got = [oracle.mf_step(st, *b, hp) for b in batches]
    P, opt, want = _literal_mf_steps(U, I, w, wu, batches, hp_kw)
    np.testing.assert_allclose(np.array(got)[:, :3], np.array(want), rtol=1e-5, atol=1e-6)
    for name, a, b in (("U", st.U, P[0]), ("I", st.I, P[1]), ("w", st.w, P[2]), ("wu", st.wu, P[3]),
                       ("mU", st.mU, opt.m[0]), ("vI", st.vI, opt.v[1])):
        ref = b.detach().numpy().reshape(a.shape)
        np.testing.assert_allclose(a, ref, rtol=2e-3, atol=2e-6, err_msg=name)
    # rows never touched by any batch did not move (m = v = 0 there)
    touched = np.unique(np.concatenate([b[0] for b in batches]))
    rest = np.setdiff1d(np.arange(n_users), touched)
    np.testing.assert_array_equal(st.U[rest], U[rest])


def test_dense_adam_moves_untouched_rows_on_momentum(oracle):
    """TF-1.14 _apply_sparse_shared decays m, v of EVERY row and moves every row."""
    rng = np.random.RandomState(0)
    var = rng.randn(10, 64).astype(np.float32)
    m = rng.randn(10, 64).astype(np.float32) * 1e-2
    v = rng.rand(10, 64).astype(np.float32) * 1e-4
    v0, m0, var0 = v.copy(), m.copy(), var.copy()
    g = rng.randn(2, 64).astype(np.float32)
    oracle.adam_sparse(var, m, v, np.array([3, 3], np.int32), g, 1e-3)
    np.testing.assert_array_equal(m[0], m0[0] * np.float32(0.9))
    np.testing.assert_array_equal(v[0], v0[0] * np.float32(0.999))
    assert np.all(var[0] != var0[0])
    gs = g[0] + g[1]  # duplicates are summed first
    np.testing.assert_array_equal(m[3], m0[3] * np.float32(0.9) + gs * np.float32(1 - np.float32(0.9)))


def test_lgcn_step_matches_literal_graph(oracle):
    n_users, n_items, B, L = 50, 30, 64, 2
    lists = make_interactions(3, n_users, n_items, 5)
    rowptr, col, val = norm_adj_csr(lists, n_users, n_items)
    U, I, w, wu = make_model(9, n_users, n_items, scale=4.0)
    hp_kw = dict(lr=1e-2, alpha=1e-2, beta=1e-3, decay=1e-3, batch_size=B)
    rng = np.random.RandomState(10)
    batches = [make_batch(rng, n_users, n_items, B) for _ in range(3)]
    st = oracle.MFState(U, I, w, wu)
    hp = oracle.HParams.make(**hp_kw)
    N = n_users + n_items
    A = torch.zeros(N, N, dtype=torch.float64)
    for r in range(N):
        A[r, col[rowptr[r]:rowptr[r + 1]]] = t64(val[rowptr[r]:rowptr[r + 1]])
    P = [t64(U).requires_grad_(True), t64(I).requires_grad_(True),
         t64(w).reshape(-1, 1).requires_grad_(True), t64(wu).reshape(-1, 1).requires_grad_(True)]
    opt = lit.TFAdam(P, hp_kw["lr"])
    # propagate agrees first
    E = oracle.lgcn_propagate(rowptr, col, val, U, I, L)
    ua, ia = lit.lightgcn_embed(A, P[0], P[1], L)
    np.testing.assert_allclose(E, torch.cat([ua, ia]).detach().numpy(), rtol=1e-5, atol=1e-7)
    for u, p, n in batches:
        lo_eval = oracle.lgcn_step(st, rowptr, col, val, L, u, p, n, hp, train=False)
        lo = oracle.lgcn_step(st, rowptr, col, val, L, u, p, n, hp, train=True)
        np.testing.assert_allclose(lo_eval, lo, rtol=1e-6)
        for q in P:
            q.grad = None
        ui, pi, ni = (torch.as_tensor(x, dtype=torch.long) for x in (u, p, n))
        ua, ia = lit.lightgcn_embed(A, P[0], P[1], L)
        mf, emb, *_ = lit.bce_two_branch_both(ua[ui], ia[pi], ia[ni], P[2], P[3], hp_kw["alpha"],
                                              hp_kw["beta"], hp_kw["decay"], B,
                                              reg_rows=(P[0][ui], P[1][pi], P[1][ni]))
        (mf + emb).backward()
        # every table element has a gradient here -> the sparse formula over all rows
        opt.step([q.grad for q in P], [True, True, False, False])
        np.testing.assert_allclose(lo[:3], [(mf + emb).item(), mf.item(), emb.item()], rtol=1e-5, atol=1e-6)
    for name, a, b in (("U", st.U, P[0]), ("I", st.I, P[1]), ("w", st.w, P[2]), ("wu", st.wu, P[3])):
        np.testing.assert_allclose(a, b.detach().numpy().reshape(a.shape), rtol=2e-3, atol=2e-6, err_msg=name)


def test_scoring_matches_literal_and_topk_is_sorted(oracle):
    n_users, n_items, T, K = 90, 333, 40, 20
    U, I, w, wu = make_model(11, n_users, n_items, scale=6.0)
    q = np.random.RandomState(1).permutation(n_users)[:T]
    Uq = np.ascontiguousarray(U[q])
    si, su = oracle.score_gates(I, w), oracle.score_gates(Uq, wu)
    S = oracle.score_matrix(Uq, I, si, su, 40.0)
    want = lit.rubi_ratings_both(t64(Uq), t64(I), t64(w).reshape(-1, 1), t64(wu).reshape(-1, 1), 40.0)
    np.testing.assert_allclose(S, want.numpy(), rtol=2e-6, atol=1e-5)
    lists = make_interactions(2, T, n_items, 9)
    mrp, mcol = lists_to_csr(lists)
    ids, sc = oracle.score_topk(Uq, I, si, su, 40.0, mrp, mcol, K)
    for t in range(T):
        row = S[t].copy()
        row[lists[t]] = -np.inf
        order = np.lexsort((np.arange(n_items), -row))[:K]  # score desc, lower id first
        np.testing.assert_array_equal(ids[t], order)
        np.testing.assert_array_equal(sc[t], row[order])
        assert not set(ids[t]) & set(lists[t])
    # item shards + merge == one shot (the multi-GPU contract)
    parts = [oracle.score_topk(Uq, np.ascontiguousarray(I[lo:hi]), np.ascontiguousarray(si[lo:hi]), su,
                               40.0, mrp, mcol, K, item_id_offset=lo)
             for lo, hi in ((0, 100), (100, 101), (101, 333))]
    mi, ms = oracle.topk_merge(np.stack([p[0] for p in parts]), np.stack([p[1] for p in parts]))
    np.testing.assert_array_equal(mi, ids)
    np.testing.assert_array_equal(ms, sc)
    # fewer than K unmasked items -> padded with -1 / -inf
    few = oracle.score_topk(Uq[:2], np.ascontiguousarray(I[:5]), np.ascontiguousarray(si[:5]), su[:2],
                            40.0, None, None, K)
    assert np.all(few[0][:, 5:] == -1) and np.all(np.isinf(few[1][:, 5:]))


def _golden_eval():
    g = np.load(os.path.join(GOLD, "ref_evaluator.npz"))
    off = np.concatenate([[0], np.cumsum(g["truth_len"])]).astype(np.int32)
    return g, off


def test_topk_and_foldout_metrics_match_reference_evaluator_golden(oracle):
    """Golden = c_top_k_array_index + evaluate_foldout of the reference (tools.h:13-33,
    evaluate_foldout.h:16-195) compiled from /root/reference, on a tie-free seeded matrix."""
    g, off = _golden_eval()
    rk = oracle.topk_rows(np.ascontiguousarray(g["scores"]), 20)
    np.testing.assert_array_equal(rk, g["rankings"])
    res = oracle.foldout_metrics(rk, off, np.ascontiguousarray(g["truth"], np.int32))
    np.testing.assert_array_equal(res, g["results"])  # float32 curves, bit for bit


def test_live_reference_evaluator_if_built(oracle):
    from oracle import ref_eval

    if not ref_eval.available():
        pytest.skip("oracle/_ref not built (no /root/reference on this box)")
    rng = np.random.RandomState(123)
    scores = rng.randn(70, 517).astype(np.float32)
    truth = [np.sort(rng.choice(517, size=int(rng.randint(1, 40)), replace=False)).astype(np.int32)
             for _ in range(70)]
    want = ref_eval.eval_score_matrix_foldout(scores, truth, top_k=20, thread_num=3)
    rk = oracle.topk_rows(scores, 20)
    rp, col = lists_to_csr(truth)
    np.testing.assert_array_equal(oracle.foldout_metrics(rk, rp, col), want)
    with pytest.raises(ValueError):
        ref_eval.eval_score_matrix_foldout(scores, truth[:-1])


def test_mf_metrics_hand_checked():
    """train.py:32-117 on a case small enough to check by hand (K=5)."""
    r = [1, 0, 0, 1, 0]
    out = mf_metrics.get_performance([10, 11, 12], r, [5])
    assert out["precision"][0] == pytest.approx(2 / 5)
    assert out["recall"][0] == pytest.approx(2 / 3)
    dcg = 1 / np.log2(2) + 1 / np.log2(5)
    idcg = 1 / np.log2(2) + 1 / np.log2(3) + 1 / np.log2(4)
    assert out["ndcg"][0] == pytest.approx(dcg / idcg)
    assert out["hit_ratio"][0] == 1.0
    agg = mf_metrics.evaluate([[10, 1, 2, 11, 3], [4, 5, 6, 7, 8]], [[10, 11, 12], [99]], [5])
    assert agg["hit_ratio"][0] == pytest.approx(0.5)
    assert agg["recall"][0] == pytest.approx((2 / 3) / 2)


def test_normalbce_step_matches_literal_graph(oracle):
    """`--train normalbce` (model.py:277-287, :100): the oracle's closed-form step vs torch
    autograd of the literal graph + TF Adam; w / w_user are not part of this graph."""
    n_users, n_items, B = 60, 40, 96
    U, I, w, wu = make_model(15, n_users, n_items, scale=4.0)
    hp_kw = dict(lr=1e-2, alpha=1e-2, beta=1e-3, decay=1e-3, batch_size=B)
    rng = np.random.RandomState(16)
    batches = [make_batch(rng, n_users, n_items, B) for _ in range(3)]
    st = oracle.MFState(U, I, w, wu)
    hp = oracle.HParams.make(**hp_kw)
    got = [oracle.mf_step_normal(st, *b, hp) for b in batches]
    P = [t64(U).requires_grad_(True), t64(I).requires_grad_(True)]
    opt = lit.TFAdam(P, hp_kw["lr"])
    want = []
    for u, p, n in batches:
        for q in P:
            q.grad = None
        ui, pi, ni = (torch.as_tensor(x, dtype=torch.long) for x in (u, p, n))
        mf, reg = lit.bce_plain(P[0][ui], P[1][pi], P[1][ni], hp_kw["decay"], hp_kw["batch_size"])
        (mf + reg).backward()
        opt.step([q.grad for q in P], [True, True])
        want.append(((mf + reg).item(), mf.item(), reg.item()))
    np.testing.assert_allclose(np.array(got)[:, :3], np.array(want), rtol=1e-5, atol=1e-6)
    for name, a, b in (("U", st.U, P[0]), ("I", st.I, P[1]), ("mU", st.mU, opt.m[0]), ("vI", st.vI, opt.v[1])):
        np.testing.assert_allclose(a, b.detach().numpy().reshape(a.shape), rtol=2e-3, atol=2e-6, err_msg=name)
    np.testing.assert_array_equal(st.w, w)  # untouched
    np.testing.assert_array_equal(st.wu, wu)


def test_lgcn_bce_step_matches_literal_graph(oracle):
    """`--loss bce` (LightGCN.py:415-429,:186): element-wise BCE on the propagated rows, L2 on the
    raw rows; oracle step vs torch autograd through the propagation."""
    n_users, n_items, B, L = 50, 30, 64, 2
    lists = make_interactions(13, n_users, n_items, 5)
    rowptr, col, val = norm_adj_csr(lists, n_users, n_items)
    U, I, w, wu = make_model(19, n_users, n_items, scale=4.0)
    hp_kw = dict(lr=1e-2, alpha=1e-2, beta=1e-3, decay=1e-3, batch_size=B)
    rng = np.random.RandomState(20)
    batches = [make_batch(rng, n_users, n_items, B) for _ in range(3)]
    st = oracle.MFState(U, I, w, wu)
    hp = oracle.HParams.make(**hp_kw)
    N = n_users + n_items
    A = torch.zeros(N, N, dtype=torch.float64)
    for r in range(N):
        A[r, col[rowptr[r]:rowptr[r + 1]]] = t64(val[rowptr[r]:rowptr[r + 1]])
    P = [t64(U).requires_grad_(True), t64(I).requires_grad_(True)]
    opt = lit.TFAdam(P, hp_kw["lr"])
    for u, p, n in batches:
        lo_eval = oracle.lgcn_step_normal(st, rowptr, col, val, L, u, p, n, hp, train=False)
        lo = oracle.lgcn_step_normal(st, rowptr, col, val, L, u, p, n, hp, train=True)
        np.testing.assert_allclose(lo_eval, lo, rtol=1e-6)
        for q in P:
            q.grad = None
        ui, pi, ni = (torch.as_tensor(x, dtype=torch.long) for x in (u, p, n))
        ua, ia = lit.lightgcn_embed(A, P[0], P[1], L)
        ps, ns = torch.sum(ua[ui] * ia[pi], 1), torch.sum(ua[ui] * ia[ni], 1)
        mf = torch.mean(-torch.log(torch.sigmoid(ps) + 1e-9) - torch.log(1 - torch.sigmoid(ns) + 1e-9))  # :423
        l2 = lambda x: torch.sum(x * x) / 2
        emb = hp_kw["decay"] * (l2(P[0][ui]) + l2(P[1][pi]) + l2(P[1][ni])) / B                      # :419-425
        (mf + emb).backward()
        opt.step([q.grad for q in P], [True, True])
        np.testing.assert_allclose(lo[:3], [(mf + emb).item(), mf.item(), emb.item()], rtol=1e-5, atol=1e-6)
    for name, a, b in (("U", st.U, P[0]), ("I", st.I, P[1])):
        np.testing.assert_allclose(a, b.detach().numpy().reshape(a.shape), rtol=2e-3, atol=2e-6, err_msg=name)
    np.testing.assert_array_equal(st.w, w)


def test_rubibce_step_matches_literal_graph(oracle):
    """`--train rubibce` (model.py:158-183, :83-85): item gate only.  Oracle closed form vs torch
    autograd of the literal graph + TF Adam; w_user receives no gradient and is never written."""
    n_users, n_items, B = 60, 40, 96
    U, I, w, wu = make_model(25, n_users, n_items, scale=4.0)
    hp_kw = dict(lr=1e-2, alpha=1e-2, beta=1e-3, decay=1e-3, batch_size=B)
    rng = np.random.RandomState(26)
    batches = [make_batch(rng, n_users, n_items, B) for _ in range(3)]
    st = oracle.MFState(U, I, w, wu)
    hp = oracle.HParams.make(**hp_kw)
    got = [oracle.mf_step_item(st, *b, hp) for b in batches]
    P = [t64(U).requires_grad_(True), t64(I).requires_grad_(True), t64(w).reshape(-1, 1).requires_grad_(True)]
    opt = lit.TFAdam(P, hp_kw["lr"])
    want = []
    for u, p, n in batches:
        for q in P:
            q.grad = None
        ui, pi, ni = (torch.as_tensor(x, dtype=torch.long) for x in (u, p, n))
        mf, reg, l_ori, _ = lit.bce_two_branch(P[0][ui], P[1][pi], P[1][ni], P[2], hp_kw["alpha"],
                                               hp_kw["decay"], hp_kw["batch_size"])
        (mf + reg).backward()
        opt.step([q.grad for q in P], [True, True, False])
        want.append(((mf + reg).item(), mf.item(), reg.item(), l_ori.item()))
    np.testing.assert_allclose(np.array(got), np.array(want), rtol=1e-5, atol=1e-6)
    for name, a, b in (("U", st.U, P[0]), ("I", st.I, P[1]), ("w", st.w, P[2]), ("mU", st.mU, opt.m[0]),
                       ("vI", st.vI, opt.v[1])):
        np.testing.assert_allclose(a, b.detach().numpy().reshape(a.shape), rtol=2e-3, atol=2e-6, err_msg=name)
    np.testing.assert_array_equal(st.wu, wu)  # untouched
    assert np.abs(st.w - w).max() > 0
    # the stand-alone grid: with sig(su) = 1 and beta = 0 the two-gate grid reduces to this one
    yp, yn, sp, sn = (rng.randn(B).astype(np.float32) for _ in range(4))
    l3, dyp, dyn, dsp, dsn = oracle.grid_bce_item(yp, yn, sp, sn, 0.01)
    big = np.full(B, 80.0, np.float32)  # sigmoid(80) == 1.0f
    l3b, (dyp_b, dyn_b, dsp_b, dsn_b, dsu_b) = oracle.grid_bce(yp, yn, sp, sn, big, 0.01, 0.0)
    np.testing.assert_array_equal(l3[:2], l3b[:2])
    for a, b in ((dyp, dyp_b), (dyn, dyn_b), (dsp, dsp_b), (dsn, dsn_b)):
        np.testing.assert_array_equal(a, b)
    assert l3[2] == 0.0 and not dsu_b.any()


def test_lgcn_bce1_step_matches_literal_graph(oracle):
    """`--loss bce1` (LightGCN.py:431-461,:190-194): item gate only on the propagated rows, L2 on
    the raw rows; oracle step vs torch autograd through the propagation."""
    n_users, n_items, B, L = 50, 30, 64, 2
    lists = make_interactions(23, n_users, n_items, 5)
    rowptr, col, val = norm_adj_csr(lists, n_users, n_items)
    U, I, w, wu = make_model(29, n_users, n_items, scale=4.0)
    hp_kw = dict(lr=1e-2, alpha=1e-2, beta=1e-3, decay=1e-3, batch_size=B)
    rng = np.random.RandomState(30)
    batches = [make_batch(rng, n_users, n_items, B) for _ in range(3)]
    st = oracle.MFState(U, I, w, wu)
    hp = oracle.HParams.make(**hp_kw)
    N = n_users + n_items
    A = torch.zeros(N, N, dtype=torch.float64)
    for r in range(N):
        A[r, col[rowptr[r]:rowptr[r + 1]]] = t64(val[rowptr[r]:rowptr[r + 1]])
    P = [t64(U).requires_grad_(True), t64(I).requires_grad_(True), t64(w).reshape(-1, 1).requires_grad_(True)]
    opt = lit.TFAdam(P, hp_kw["lr"])
    for u, p, n in batches:
        lo_eval = oracle.lgcn_step_item(st, rowptr, col, val, L, u, p, n, hp, train=False)
        lo = oracle.lgcn_step_item(st, rowptr, col, val, L, u, p, n, hp, train=True)
        np.testing.assert_allclose(lo_eval, lo, rtol=1e-6)
        for q in P:
            q.grad = None
        ui, pi, ni = (torch.as_tensor(x, dtype=torch.long) for x in (u, p, n))
        ua, ia = lit.lightgcn_embed(A, P[0], P[1], L)
        mf, emb, *_ = lit.bce_two_branch(ua[ui], ia[pi], ia[ni], P[2], hp_kw["alpha"], hp_kw["decay"], B,
                                         reg_rows=(P[0][ui], P[1][pi], P[1][ni]))
        (mf + emb).backward()
        opt.step([q.grad for q in P], [True, True, False])
        np.testing.assert_allclose(lo[:3], [(mf + emb).item(), mf.item(), emb.item()], rtol=1e-5, atol=1e-6)
    for name, a, b in (("U", st.U, P[0]), ("I", st.I, P[1]), ("w", st.w, P[2])):
        np.testing.assert_allclose(a, b.detach().numpy().reshape(a.shape), rtol=2e-3, atol=2e-6, err_msg=name)
    np.testing.assert_array_equal(st.wu, wu)
