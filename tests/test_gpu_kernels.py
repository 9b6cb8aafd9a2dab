"""-m gpu: every CUDA entry point of the C ABI against the CPU oracle on seeded inputs."""
import numpy as np
import pytest

from helpers import lists_to_csr, make_batch, make_interactions, make_model, norm_adj_csr

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def T():
    import torch

    return torch


@pytest.fixture(scope="module")
def ops():
    from macr_b200 import ops as o

    return o


def dev(T, a):
    return T.as_tensor(np.ascontiguousarray(a)).cuda()


def test_library_loaded():
    from macr_b200 import _lib

    assert _lib.lib().macr_abi_version() == 1


@pytest.mark.parametrize("B", [1, 37, 1024])
def test_gather_dots(T, ops, oracle, B):
    U, I, w, wu = make_model(1, 500, 300, scale=5.0)
    rng = np.random.RandomState(B)
    u, p, n = make_batch(rng, 500, 300, B)
    want = oracle.gather_dots(U, I, U, I, w, wu, u, p, n)
    got = ops.gather_dots(dev(T, U), dev(T, I), dev(T, U), dev(T, I), dev(T, w), dev(T, wu),
                          dev(T, u), dev(T, p), dev(T, n))
    for a, b in zip(got, want):
        np.testing.assert_allclose(a.cpu().numpy(), b, rtol=2e-6, atol=2e-6)


@pytest.mark.parametrize("B,scale", [(1, 1.0), (64, 1.0), (100, 3.0), (1024, 1.0), (2048, 8.0),
                                     (2500, 1.0), (4096, 30.0)])
def test_grid_bce(T, ops, oracle, B, scale):
    """B x B broadcast grid (model.py:204-217): losses within 1e-4 (north_star tolerance),
    gradients within 1e-5 relative to their scale. scale>1 drives saturated logits through the
    literal (slow) path of the kernel."""
    rng = np.random.RandomState(B)
    yp, yn = (rng.randn(B) * scale).astype(np.float32), (rng.randn(B) * scale).astype(np.float32)
    sp, sn, su = [(rng.randn(B) * min(scale, 4.0)).astype(np.float32) for _ in range(3)]
    alpha, beta = 1e-2, 1e-3
    l_want, g_want = oracle.grid_bce(yp, yn, sp, sn, su, alpha, beta)
    l_got, g_got = ops.grid_bce(*[dev(T, x) for x in (yp, yn, sp, sn, su)], alpha, beta)
    np.testing.assert_allclose(l_got.cpu().numpy(), l_want, rtol=1e-4, atol=1e-4)
    for a, b in zip(g_got, g_want):
        tol = 1e-5 * max(1e-30, float(np.abs(b).max()))
        np.testing.assert_allclose(a.cpu().numpy(), b, rtol=2e-4, atol=tol)
    l_only, none = ops.grid_bce(*[dev(T, x) for x in (yp, yn, sp, sn, su)], alpha, beta,
                                want_grad=False)
    assert none is None
    np.testing.assert_allclose(l_only.cpu().numpy(), l_got.cpu().numpy(), rtol=1e-6)


@pytest.mark.parametrize("n_ids,rows", [(1, 10), (33, 7), (1024, 100), (8192, 40981), (16384, 500)])
def test_batch_plan(T, ops, n_ids, rows):
    rng = np.random.RandomState(n_ids)
    ids = rng.randint(0, rows, n_ids).astype(np.int32)
    bitmap = T.zeros((rows + 31) // 32, dtype=T.int32, device="cuda")
    uniq, seg_off, seg_pos = ops.batch_plan(dev(T, ids), rows, bitmap.view(T.int32))
    uniq, seg_off, seg_pos = uniq.cpu().numpy(), seg_off.cpu().numpy(), seg_pos.cpu().numpy()
    want_uniq = np.unique(ids)  # ascending rows; the order of the unique rows enters no result
    np.testing.assert_array_equal(uniq, want_uniq)
    assert seg_off[0] == 0 and seg_off[-1] == n_ids
    for k, r in enumerate(want_uniq):
        pos = seg_pos[seg_off[k]:seg_off[k + 1]]
        np.testing.assert_array_equal(pos, np.nonzero(ids == r)[0])  # ascending positions
    bits = np.unpackbits(bitmap.cpu().numpy().view(np.uint8), bitorder="little")[:rows]
    want_bits = np.zeros(rows, np.uint8)
    want_bits[np.unique(ids)] = 1
    np.testing.assert_array_equal(bits, want_bits)


def test_adam_kernels_bit_exact(T, ops, oracle):
    """TF-1.14 dense-semantics Adam: untouched-row sweep + touched rows == oracle, bit for bit."""
    rng = np.random.RandomState(3)
    rows, d = 3000, 64
    var = rng.randn(rows, d).astype(np.float32) * 0.1
    m = rng.randn(rows, d).astype(np.float32) * 1e-3
    v = (rng.rand(rows, d).astype(np.float32)) * 1e-5
    m[::7] = 0
    v[::7] = 0
    idx = rng.permutation(np.unique(rng.randint(0, rows, 500))).astype(np.int32)
    g = rng.randn(len(idx), d).astype(np.float32) * 1e-3
    lr_t = oracle.adam_lr_t(1e-3, 0.9 ** 5, 0.999 ** 5)
    ov, om, ovv = var.copy(), m.copy(), v.copy()
    oracle.adam_sparse(ov, om, ovv, idx, g, lr_t)
    dv, dm, dvv = dev(T, var), dev(T, m), dev(T, v)
    bitmap = T.zeros((rows + 31) // 32, dtype=T.int32, device="cuda")
    uniq, _, _ = ops.batch_plan(dev(T, idx), rows, bitmap)
    ops.adam_sweep_untouched(dv, dm, dvv, bitmap, lr_t)
    np.testing.assert_array_equal(uniq.cpu().numpy(), np.sort(idx))  # plan lists rows ascending
    ops.adam_rows(dv, dm, dvv, uniq, dev(T, g[np.argsort(idx)]), bitmap, lr_t)
    assert int(bitmap.abs().sum().item()) == 0  # bits cleared for the next step
    np.testing.assert_array_equal(dv.cpu().numpy(), ov)
    np.testing.assert_array_equal(dm.cpu().numpy(), om)
    np.testing.assert_array_equal(dvv.cpu().numpy(), ovv)
    # all-rows-have-a-gradient variant (LightGCN tables)
    gd = rng.randn(rows, d).astype(np.float32) * 1e-3
    ov, om, ovv = var.copy(), m.copy(), v.copy()
    oracle.adam_sparse(ov, om, ovv, np.arange(rows, dtype=np.int32), gd, lr_t)
    dv, dm, dvv = dev(T, var), dev(T, m), dev(T, v)
    ops.adam_dense(dv, dm, dvv, dev(T, gd), lr_t)
    np.testing.assert_array_equal(dv.cpu().numpy(), ov)
    np.testing.assert_array_equal(dm.cpu().numpy(), om)
    # ApplyAdam for w / w_user
    wv, wm, wvv, wg = [rng.randn(64).astype(np.float32) * s for s in (0.3, 1e-3, 0.0, 1e-2)]
    wvv = np.abs(wvv) + 1e-6
    o1, o2, o3 = wv.copy(), wm.copy(), wvv.astype(np.float32).copy()
    oracle.adam_dense_vec(o1, o2, o3, wg, lr_t)
    d1, d2, d3 = dev(T, wv), dev(T, wm), dev(T, wvv.astype(np.float32))
    ops.adam_vec(d1, d2, d3, dev(T, wg), lr_t)
    np.testing.assert_array_equal(d1.cpu().numpy(), o1)
    np.testing.assert_array_equal(d2.cpu().numpy(), o2)
    np.testing.assert_array_equal(d3.cpu().numpy(), o3)


def _run_mf(T, ops, oracle, n_users, n_items, B, steps, alpha, scale, seed, host=False):
    U, I, w, wu = make_model(seed, n_users, n_items, scale=scale)
    hp_o = oracle.HParams.make(lr=1e-3, alpha=alpha, beta=1e-3, decay=1e-5, batch_size=B)
    hp_g = ops.HParams.make(lr=1e-3, alpha=alpha, beta=1e-3, decay=1e-5, batch_size=B)
    st = oracle.MFState(U, I, w, wu)
    tr = ops.MFTrainer(U, I, w, wu, hp_g, max_batch=B)
    rng = np.random.RandomState(seed + 1)
    for s in range(steps):
        u, p, n = make_batch(rng, n_users, n_items, B)
        lo = oracle.mf_step(st, u, p, n, hp_o)
        if host:
            lg = np.array(tr.step_host(u.tolist(), p.tolist(), n.tolist()))
        else:
            lg = tr.step_device(dev(T, u), dev(T, p), dev(T, n)).cpu().numpy()[:3]
        np.testing.assert_allclose(lg, lo[:3], rtol=1e-4, atol=1e-4, err_msg=f"step {s}")
    T.cuda.synchronize()
    t = tr.tab
    for name, got, want in (("U", t.U, st.U), ("I", t.I, st.I), ("w", t.w, st.w),
                            ("wu", t.wu, st.wu), ("mU", t.mU, st.mU), ("vI", t.vI, st.vI)):
        np.testing.assert_allclose(got.cpu().numpy(), want, rtol=1e-4, atol=1e-5, err_msg=name)
    assert tr.steps_done == steps
    tr.close()


@pytest.mark.parametrize("B,steps", [(64, 3), (1000, 3), (1024, 10)])
def test_mf_trainer_steps(T, ops, oracle, B, steps):
    _run_mf(T, ops, oracle, 1500, 744, B, steps, 1e-3, 3.0, seed=B)


def test_mf_trainer_step_host(T, ops, oracle):
    _run_mf(T, ops, oracle, 1500, 744, 512, 4, 1e-2, 3.0, seed=11, host=True)


def test_mf_trainer_duplicates_and_big_batch(T, ops, oracle):
    """B > n_users (duplicate users, load_data.py:546-547) and the 128-tile grid (B=2048)."""
    _run_mf(T, ops, oracle, 700, 300, 2048, 2, 1e-3, 4.0, seed=5)


def test_mf_trainer_epoch_mode(T, ops, oracle):
    n_users, n_items, B, steps = 2000, 900, 256, 6
    U, I, w, wu = make_model(7, n_users, n_items, scale=3.0)
    hp_o = oracle.HParams.make(batch_size=B)
    hp_g = ops.HParams.make(batch_size=B)
    st = oracle.MFState(U, I, w, wu)
    tr = ops.MFTrainer(U, I, w, wu, hp_g, max_batch=B)
    rng = np.random.RandomState(8)
    batches = np.stack([np.stack(make_batch(rng, n_users, n_items, B)) for _ in range(steps)])
    want = np.stack([oracle.mf_step(st, *batches[s], hp_o) for s in range(steps)])
    got = tr.run(dev(T, batches.astype(np.int32))).cpu().numpy()
    np.testing.assert_allclose(got, want, rtol=1e-4, atol=1e-4)
    np.testing.assert_allclose(tr.tab.U.cpu().numpy(), st.U, rtol=1e-4, atol=1e-5)
    np.testing.assert_allclose(tr.tab.I.cpu().numpy(), st.I, rtol=1e-4, atol=1e-5)
    assert tr.launches_per_step >= 5
    tr.close()


def _graph(seed, n_users, n_items, deg):
    lists = make_interactions(seed, n_users, n_items, deg)
    return lists, norm_adj_csr(lists, n_users, n_items)


def test_spmm_and_propagate(T, ops, oracle):
    n_users, n_items = 700, 400
    _, (rowptr, col, val) = _graph(2, n_users, n_items, 12)
    U, I, _, _ = make_model(4, n_users, n_items, scale=3.0)
    X = np.concatenate([U, I], 0)
    want = oracle.spmm_csr(rowptr, col, val, X)
    got = ops.spmm_csr(dev(T, rowptr), dev(T, col), dev(T, val), dev(T, X)).cpu().numpy()
    np.testing.assert_allclose(got, want, rtol=2e-6, atol=1e-7)
    for L in (0, 1, 2, 3):
        wantE = oracle.lgcn_propagate(rowptr, col, val, U, I, L)
        gotE = ops.lgcn_propagate(dev(T, rowptr), dev(T, col), dev(T, val), dev(T, U), dev(T, I),
                                  L).cpu().numpy()
        np.testing.assert_allclose(gotE, wantE, rtol=3e-6, atol=1e-7, err_msg=f"L={L}")


@pytest.mark.parametrize("L,B,steps", [(2, 256, 4), (1, 100, 2), (3, 1024, 2)])
def test_lgcn_trainer_steps(T, ops, oracle, L, B, steps):
    n_users, n_items = 1200, 500
    _, (rowptr, col, val) = _graph(L, n_users, n_items, 10)
    U, I, w, wu = make_model(9, n_users, n_items, scale=4.0)
    hp_o = oracle.HParams.make(lr=1e-3, alpha=1e-2, beta=1e-3, decay=1e-5, batch_size=B)
    hp_g = ops.HParams.make(lr=1e-3, alpha=1e-2, beta=1e-3, decay=1e-5, batch_size=B)
    st = oracle.MFState(U, I, w, wu)
    tr = ops.LGCNTrainer(rowptr, col, val, U, I, w, wu, L, hp_g, max_batch=B)
    rng = np.random.RandomState(L)
    for s in range(steps):
        u, p, n = make_batch(rng, n_users, n_items, B)
        # loss-only pass first (LightGCN.py:616-647), must not move anything
        lo0 = oracle.lgcn_step(st, rowptr, col, val, L, u, p, n, hp_o, train=False)
        lg0 = tr.step_device(dev(T, u), dev(T, p), dev(T, n), train=False).cpu().numpy()
        np.testing.assert_allclose(lg0[:3], lo0[:3], rtol=1e-4, atol=1e-4)
        lo = oracle.lgcn_step(st, rowptr, col, val, L, u, p, n, hp_o, train=True)
        lg = np.array(tr.step_host(u.tolist(), p.tolist(), n.tolist()))
        np.testing.assert_allclose(lg, lo[:3], rtol=1e-4, atol=1e-4, err_msg=f"step {s}")
    t = tr.tab
    for name, got, want in (("U", t.U, st.U), ("I", t.I, st.I), ("w", t.w, st.w), ("wu", t.wu, st.wu)):
        np.testing.assert_allclose(got.cpu().numpy(), want, rtol=1e-4, atol=1e-5, err_msg=name)
    ue, ie = tr.embeddings()
    wantE = oracle.lgcn_propagate(rowptr, col, val, st.U, st.I, L)
    np.testing.assert_allclose(T.cat([ue, ie]).cpu().numpy(), wantE, rtol=1e-4, atol=1e-5)
    tr.close()


def _score_inputs(seed, T_users, n_items, n_all_users=None, mask_deg=30):
    n_all_users = n_all_users or T_users
    U, I, w, wu = make_model(seed, n_all_users, n_items, scale=10.0)
    rng = np.random.RandomState(seed)
    q = rng.permutation(n_all_users)[:T_users].astype(np.int32)
    lists = make_interactions(seed + 1, T_users, n_items, mask_deg)
    return U, I, w, wu, q, lists


@pytest.mark.parametrize("T_users,n_items,K,c", [(1, 5, 3, 40.0), (130, 744, 20, 40.0),
                                                 (300, 1000, 32, 0.0), (257, 4133, 20, 40.0)])
def test_score_topk_bit_exact(T, ops, oracle, T_users, n_items, K, c):
    """score + mask + top-K: ids AND scores bit-identical to the oracle (fp32 FMA chain)."""
    U, I, w, wu, q, lists = _score_inputs(T_users, T_users, n_items, mask_deg=min(30, n_items // 3))
    Uq = U[q]
    sig_i, sig_u = oracle.score_gates(I, w), oracle.score_gates(Uq, wu)
    mrp, mcol = lists_to_csr(lists)
    want_ids, want_sc = oracle.score_topk(Uq, I, sig_i, sig_u, c, mrp, mcol, K)
    dI, dq = dev(T, I), dev(T, q)
    dUq = ops.gather_rows(dev(T, U), dq)
    np.testing.assert_array_equal(dUq.cpu().numpy(), Uq)
    gsi, gsu = ops.score_gates(dI, dev(T, w)), ops.score_gates(dUq, dev(T, wu))
    np.testing.assert_array_equal(gsi.cpu().numpy(), sig_i)
    np.testing.assert_array_equal(gsu.cpu().numpy(), sig_u)
    ids, sc = ops.score_topk(dUq, dI, gsi, gsu, c, dev(T, mrp), dev(T, mcol), K)
    np.testing.assert_array_equal(ids.cpu().numpy(), want_ids)
    np.testing.assert_array_equal(sc.cpu().numpy(), want_sc)
    # no mask
    w2, _ = oracle.score_topk(Uq, I, sig_i, sig_u, c, None, None, K)
    g2, _ = ops.score_topk(dUq, dI, gsi, gsu, c, None, None, K)
    np.testing.assert_array_equal(g2.cpu().numpy(), w2)
    # dense matrix fetch (rubi_ratings_both) bit-exact too
    M = ops.score_matrix(dUq, dI, gsi, gsu, c).cpu().numpy()
    np.testing.assert_array_equal(M, oracle.score_matrix(Uq, I, sig_i, sig_u, c))


def test_score_topk_ties_and_short_rows(T, ops, oracle):
    """all-equal scores -> lowest ids win; fewer than K unmasked items -> -1 / -inf padding."""
    n_items, K = 50, 20
    Uq = np.zeros((3, 64), np.float32)
    I = np.zeros((n_items, 64), np.float32)
    sig_i = np.full(n_items, 0.5, np.float32)
    sig_u = np.full(3, 0.5, np.float32)
    lists = [np.arange(0, 45, dtype=np.int32), np.array([0, 2], np.int32), np.zeros(0, np.int32)]
    mrp, mcol = lists_to_csr(lists)
    want_ids, want_sc = oracle.score_topk(Uq, I, sig_i, sig_u, 40.0, mrp, mcol, K)
    ids, sc = ops.score_topk(dev(T, Uq), dev(T, I), dev(T, sig_i), dev(T, sig_u), 40.0,
                             dev(T, mrp), dev(T, mcol), K)
    np.testing.assert_array_equal(ids.cpu().numpy(), want_ids)
    np.testing.assert_array_equal(sc.cpu().numpy(), want_sc)
    assert list(want_ids[0][:5]) == [45, 46, 47, 48, 49] and want_ids[0][5] == -1
    assert list(want_ids[1][:3]) == [1, 3, 4]


def test_sharded_scoring_merge_equals_single(T, ops, oracle):
    """item-sharded scoring + macr_topk_merge == unsharded (what the N-GPU path does)."""
    U, I, w, wu, q, lists = _score_inputs(21, 200, 3000)
    Uq = U[q]
    sig_i, sig_u = oracle.score_gates(I, w), oracle.score_gates(Uq, wu)
    mrp, mcol = lists_to_csr(lists)
    K = 20
    want_ids, want_sc = oracle.score_topk(Uq, I, sig_i, sig_u, 40.0, mrp, mcol, K)
    dUq, dsu = dev(T, Uq), dev(T, sig_u)
    parts_i, parts_s = [], []
    G = 4
    bounds = np.linspace(0, 3000, G + 1).astype(int)
    for g in range(G):
        lo, hi = bounds[g], bounds[g + 1]
        ids, sc = ops.score_topk(dUq, dev(T, I[lo:hi]), dev(T, sig_i[lo:hi]), dsu, 40.0,
                                 dev(T, mrp), dev(T, mcol), K, item_id_offset=int(lo))
        parts_i.append(ids)
        parts_s.append(sc)
    mi, ms = ops.topk_merge(T.stack(parts_i).contiguous(), T.stack(parts_s).contiguous())
    np.testing.assert_array_equal(mi.cpu().numpy(), want_ids)
    np.testing.assert_array_equal(ms.cpu().numpy(), want_sc)
    oi, osc = oracle.topk_merge(T.stack(parts_i).cpu().numpy(), T.stack(parts_s).cpu().numpy())
    np.testing.assert_array_equal(oi, want_ids)


def test_topk_rows_and_foldout_metrics(T, ops, oracle):
    """drop-in for c_top_k_array_index + evaluate_foldout (tools.h:24-33, evaluate_foldout.h)."""
    rng = np.random.RandomState(0)
    rows, cols, K = 300, 777, 20
    scores = rng.randn(rows, cols).astype(np.float32)
    scores[5, :] = 1.0  # ties
    scores[6, :100] = -np.inf  # masked train items, batch_test.py:129
    truth = [np.sort(rng.choice(cols, size=rng.randint(1, 40), replace=False)).astype(np.int32)
             for _ in range(rows)]
    trp, tcol = lists_to_csr(truth)
    want_rk = oracle.topk_rows(scores, K)
    got_rk = ops.topk_rows(dev(T, scores), K)
    np.testing.assert_array_equal(got_rk.cpu().numpy(), want_rk)
    want = oracle.foldout_metrics(want_rk, trp, tcol)
    got = ops.foldout_metrics(got_rk, dev(T, trp), dev(T, tcol)).cpu().numpy()
    np.testing.assert_array_equal(got, want)  # bit-exact float32 curves


def test_errors_are_loud(T, ops):
    from macr_b200._lib import MacrError

    x = T.zeros((4, 32), dtype=T.float32, device="cuda")
    with pytest.raises(MacrError):
        ops.spmm_csr(T.zeros(5, dtype=T.int32, device="cuda"), T.zeros(1, dtype=T.int32, device="cuda"),
                     T.zeros(1, dtype=T.float32, device="cuda"), x)  # d != 64
    with pytest.raises(MacrError):
        ops.topk_rows(T.zeros((2, 100), dtype=T.float32, device="cuda"), 64)  # K > 32
    with pytest.raises(MacrError):
        ops.gather_dots(x.cpu(), x, x, x, x, x, x, x, x)  # host tensor


def test_normalbce_trainer_steps_match_oracle(T, ops, oracle):
    """`--train normalbce` (README.md:30; model.py:277-287,:100): element-wise BCE step on the
    same plan / row-gradient / TF-Adam machinery; w and w_user stay untouched."""
    n_users, n_items, B, steps = 900, 500, 256, 5
    U, I, w, wu = make_model(41, n_users, n_items, scale=4.0)
    hp_kw = dict(lr=1e-2, alpha=1e-2, beta=1e-3, decay=1e-4, batch_size=B)
    st = oracle.MFState(U, I, w, wu)
    tr = ops.MFTrainer(U, I, w, wu, ops.HParams.make(**hp_kw), max_batch=B)
    tr.set_mode(ops.MFTrainer.NORMALBCE)
    rng = np.random.RandomState(42)
    for s in range(steps):
        u, p, n = make_batch(rng, n_users, n_items, B)
        want = oracle.mf_step_normal(st, u, p, n, oracle.HParams.make(**hp_kw))
        got = tr.step_host(u.tolist(), p.tolist(), n.tolist())
        np.testing.assert_allclose(np.array(got), want[:3], rtol=1e-4, atol=1e-5, err_msg=f"step {s}")
    t = tr.tab
    for name, g, o in (("U", t.U, st.U), ("I", t.I, st.I), ("mU", t.mU, st.mU), ("vI", t.vI, st.vI)):
        np.testing.assert_allclose(g.cpu().numpy(), o, rtol=1e-4, atol=1e-5, err_msg=name)
    np.testing.assert_array_equal(t.w.cpu().numpy(), w)
    np.testing.assert_array_equal(t.wu.cpu().numpy(), wu)
    # switching back re-captures the MACR graph
    tr.set_mode(ops.MFTrainer.RUBIBCEBOTH)
    u, p, n = make_batch(rng, n_users, n_items, B)
    want = oracle.mf_step(st, u, p, n, oracle.HParams.make(**hp_kw))
    got = tr.step_host(u.tolist(), p.tolist(), n.tolist())
    np.testing.assert_allclose(np.array(got), want[:3], rtol=1e-4, atol=1e-4)
    tr.close()


def test_lgcn_bce_trainer_steps_match_oracle(T, ops, oracle):
    """`--loss bce` (README.md:59; LightGCN.py:415-429,:186) on the LightGCN trainer."""
    n_users, n_items, L, B, steps = 1200, 500, 2, 256, 4
    _, (rowptr, col, val) = _graph(L, n_users, n_items, 10)
    U, I, w, wu = make_model(29, n_users, n_items, scale=4.0)
    hp_kw = dict(lr=1e-3, alpha=1e-2, beta=1e-3, decay=1e-5, batch_size=B)
    st = oracle.MFState(U, I, w, wu)
    tr = ops.LGCNTrainer(rowptr, col, val, U, I, w, wu, L, ops.HParams.make(**hp_kw), max_batch=B)
    tr.set_mode(ops.LGCNTrainer.NORMALBCE)
    rng = np.random.RandomState(30)
    for s in range(steps):
        u, p, n = make_batch(rng, n_users, n_items, B)
        lo_eval = tr.step_host(u.tolist(), p.tolist(), n.tolist(), train=False)
        want = oracle.lgcn_step_normal(st, rowptr, col, val, L, u, p, n, oracle.HParams.make(**hp_kw))
        got = tr.step_host(u.tolist(), p.tolist(), n.tolist())
        np.testing.assert_allclose(np.array(got), want[:3], rtol=1e-4, atol=1e-5, err_msg=f"step {s}")
        np.testing.assert_allclose(np.array(lo_eval), want[:3], rtol=1e-4, atol=1e-5)
    t = tr.tab
    for name, g, o in (("U", t.U, st.U), ("I", t.I, st.I)):
        np.testing.assert_allclose(g.cpu().numpy(), o, rtol=1e-4, atol=1e-5, err_msg=name)
    np.testing.assert_array_equal(t.w.cpu().numpy(), w)
    tr.close()


def test_planned_spmm_with_hub_rows_matches_oracle(T, ops, oracle):
    """macr_spmm_plan_* / *_planned: rows far longer than one 64-nonzero segment (a hub item
    adjacent to every user) and empty rows; same results as the oracle up to the rounding of the
    segment-wise summation, and identical between two calls (deterministic combine order)."""
    import scipy.sparse as sp

    n_users, n_items = 900, 300
    rng = np.random.RandomState(3)
    R = sp.random(n_users, n_items, density=0.02, random_state=rng, format="lil", dtype=np.float32)
    R[:, 7] = 1.0          # hub: 900 nonzeros in one row of A
    R[5, :] = 0.0          # a user with no interaction -> empty row
    R = (R.tocsr() != 0).astype(np.float32)
    A = sp.bmat([[None, R], [R.T, None]], format="csr", dtype=np.float32)
    deg = np.asarray(A.sum(1)).ravel()
    with np.errstate(divide="ignore"):
        dinv = np.power(deg, -0.5)
    dinv[np.isinf(dinv)] = 0
    N = sp.diags(dinv).dot(A).dot(sp.diags(dinv)).tocsr().astype(np.float32)
    N.sort_indices()
    rowptr, col, val = N.indptr.astype(np.int32), N.indices.astype(np.int32), N.data.astype(np.float32)
    assert np.diff(rowptr).max() >= 890 and np.diff(rowptr).min() == 0
    U, I, _, _ = make_model(8, n_users, n_items, scale=3.0)
    X = np.concatenate([U, I], 0)
    d_rp, d_col, d_val = dev(T, rowptr), dev(T, col), dev(T, val)
    plan = ops.SpmmPlan(d_rp)
    want = oracle.spmm_csr(rowptr, col, val, X)
    got = ops.spmm_csr(d_rp, d_col, d_val, dev(T, X), plan=plan)
    np.testing.assert_allclose(got.cpu().numpy(), want, rtol=3e-6, atol=2e-7)
    again = ops.spmm_csr(d_rp, d_col, d_val, dev(T, X), plan=plan)
    assert T.equal(got, again)
    plain = ops.spmm_csr(d_rp, d_col, d_val, dev(T, X))  # stateless kernel, CTA-cooperative hub
    np.testing.assert_allclose(plain.cpu().numpy(), want, rtol=3e-6, atol=2e-7)
    for L in (1, 2):
        wantE = oracle.lgcn_propagate(rowptr, col, val, U, I, L)
        gotE = ops.lgcn_propagate(d_rp, d_col, d_val, dev(T, U), dev(T, I), L, plan=plan)
        np.testing.assert_allclose(gotE.cpu().numpy(), wantE, rtol=5e-6, atol=2e-7)
    plan.close()


def test_rubibce_trainer_steps_match_oracle(T, ops, oracle):
    """`--train rubibce` (model.py:158-183,:83-85): the B x B grid with the item gate only, on the
    same kernels (gather_dots writes gate 1 / user loss 0); w_user and its slots are never written,
    also after steps of the two-gate graph left a momentum in them."""
    n_users, n_items, B, steps = 900, 500, 256, 5
    U, I, w, wu = make_model(51, n_users, n_items, scale=4.0)
    hp_kw = dict(lr=1e-2, alpha=1e-2, beta=1e-3, decay=1e-4, batch_size=B)
    st = oracle.MFState(U, I, w, wu)
    tr = ops.MFTrainer(U, I, w, wu, ops.HParams.make(**hp_kw), max_batch=B)
    tr.set_mode(ops.MFTrainer.RUBIBCE)
    rng = np.random.RandomState(52)
    for s in range(steps):
        u, p, n = make_batch(rng, n_users, n_items, B)
        want = oracle.mf_step_item(st, u, p, n, oracle.HParams.make(**hp_kw))
        got = tr.step_host(u.tolist(), p.tolist(), n.tolist())
        np.testing.assert_allclose(np.array(got), want[:3], rtol=1e-4, atol=1e-5, err_msg=f"step {s}")
    t = tr.tab
    for name, g, o in (("U", t.U, st.U), ("I", t.I, st.I), ("w", t.w, st.w), ("mU", t.mU, st.mU),
                       ("vI", t.vI, st.vI), ("mw", t.mw, st.mw)):
        np.testing.assert_allclose(g.cpu().numpy(), o, rtol=1e-4, atol=1e-5, err_msg=name)
    np.testing.assert_array_equal(t.wu.cpu().numpy(), wu)
    assert not t.mwu.any().item() and not t.vwu.any().item()
    assert np.abs(t.w.cpu().numpy() - w).max() > 0
    # one two-gate step gives w_user a momentum; the item-only graph must then leave all three alone
    tr.set_mode(ops.MFTrainer.RUBIBCEBOTH)
    u, p, n = make_batch(rng, n_users, n_items, B)
    want = oracle.mf_step(st, u, p, n, oracle.HParams.make(**hp_kw))
    got = tr.step_host(u.tolist(), p.tolist(), n.tolist())
    np.testing.assert_allclose(np.array(got), want[:3], rtol=1e-4, atol=1e-4)
    frozen = [x.clone() for x in (t.wu, t.mwu, t.vwu)]
    assert frozen[1].any().item()
    tr.set_mode(ops.MFTrainer.RUBIBCE)
    u, p, n = make_batch(rng, n_users, n_items, B)
    want = oracle.mf_step_item(st, u, p, n, oracle.HParams.make(**hp_kw))
    got = tr.step_host(u.tolist(), p.tolist(), n.tolist())
    np.testing.assert_allclose(np.array(got), want[:3], rtol=1e-4, atol=1e-4)
    for a, b in zip(frozen, (t.wu, t.mwu, t.vwu)):
        assert T.equal(a, b)
    np.testing.assert_allclose(t.U.cpu().numpy(), st.U, rtol=1e-4, atol=1e-5)
    tr.close()


def test_lgcn_bce1_trainer_steps_match_oracle(T, ops, oracle):
    """`--loss bce1` (LightGCN.py:431-461,:190-194) on the LightGCN trainer."""
    n_users, n_items, L, B, steps = 1200, 500, 2, 256, 4
    _, (rowptr, col, val) = _graph(L, n_users, n_items, 10)
    U, I, w, wu = make_model(39, n_users, n_items, scale=4.0)
    hp_kw = dict(lr=1e-3, alpha=1e-2, beta=1e-3, decay=1e-5, batch_size=B)
    st = oracle.MFState(U, I, w, wu)
    tr = ops.LGCNTrainer(rowptr, col, val, U, I, w, wu, L, ops.HParams.make(**hp_kw), max_batch=B)
    tr.set_mode(ops.LGCNTrainer.RUBIBCE)
    rng = np.random.RandomState(40)
    for s in range(steps):
        u, p, n = make_batch(rng, n_users, n_items, B)
        lo_eval = tr.step_host(u.tolist(), p.tolist(), n.tolist(), train=False)
        want = oracle.lgcn_step_item(st, rowptr, col, val, L, u, p, n, oracle.HParams.make(**hp_kw))
        got = tr.step_host(u.tolist(), p.tolist(), n.tolist())
        np.testing.assert_allclose(np.array(got), want[:3], rtol=1e-4, atol=1e-5, err_msg=f"step {s}")
        np.testing.assert_allclose(np.array(lo_eval), want[:3], rtol=1e-4, atol=1e-5)
    t = tr.tab
    for name, g, o in (("U", t.U, st.U), ("I", t.I, st.I), ("w", t.w, st.w)):
        np.testing.assert_allclose(g.cpu().numpy(), o, rtol=1e-4, atol=1e-5, err_msg=name)
    np.testing.assert_array_equal(t.wu.cpu().numpy(), wu)
    tr.close()


def test_item_gate_score_head_matches_literal(T, ops, oracle):
    """`rubi_ratings` (model.py:141; LightGCN.py:442): (y - c) * sig(i.w) -- the fused kernel with a
    user gate of ones is bit-identical to the oracle's two-gate score with sig_u = 1 and agrees
    with the literal restatement."""
    import torch

    from oracle import literal_torch as lit

    U, I, w, wu = make_model(61, 300, 5000, scale=6.0)
    si = oracle.score_gates(I, w)
    ones = np.ones(300, np.float32)
    want_ids, want_sc = oracle.score_topk(U, I, si, ones, 3.0, None, None, 20)
    dU, dI = dev(T, U), dev(T, I)
    ids, sc = ops.score_topk(dU, dI, ops.score_gates(dI, dev(T, w)), dev(T, ones), 3.0, None, None, 20)
    np.testing.assert_array_equal(ids.cpu().numpy(), want_ids)
    np.testing.assert_array_equal(sc.cpu().numpy(), want_sc)
    M = lit.rubi_ratings(torch.tensor(U, dtype=torch.float64), torch.tensor(I, dtype=torch.float64),
                         torch.tensor(w, dtype=torch.float64).reshape(-1, 1), 3.0).numpy()
    np.testing.assert_allclose(np.take_along_axis(M, want_ids.astype(np.int64), 1), want_sc, rtol=2e-5, atol=2e-5)
