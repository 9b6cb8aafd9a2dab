"""GPU parity of the tcgen05 / TMA scoring path (macr_b200/csrc/score_tc.cu) through the C ABI.

The contract is the one of `macr_score_topk` (model.py:45,199;
train.py:89-104,133; batch_test.py:124-134): ids AND scores bit-identical to the oracle's fp32
FMA chain, ties -> lower id, train items excluded, short rows padded with -1 / -inf.  The
tensor-core passes only select candidates; every emitted score is re-computed exactly.
"""
import numpy as np
import pytest

from helpers import lists_to_csr, make_interactions, make_model

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def T():
    import torch

    assert torch.cuda.is_available()
    return torch


@pytest.fixture(scope="module")
def ops():
    from macr_b200 import ops as o

    return o


def dev(T, a):
    return T.from_numpy(np.ascontiguousarray(a)).cuda()


def _gates(T, ops, U, I, w, wu):
    dU, dI = dev(T, U), dev(T, I)
    return dU, dI, ops.score_gates(dI, dev(T, w)), ops.score_gates(dU, dev(T, wu))


@pytest.mark.parametrize("T_users,n_items,K,c,deg,scale", [
    (130, 4500, 20, 40.0, 30, 10.0),    # ragged last tiles on both sides
    (300, 6000, 32, 0.0, 30, 10.0),     # K = 32, c = 0
    (257, 8790, 20, 40.0, 71, 10.0),    # ml_10m catalogue, dense train lists
    (200, 5000, 1, -3.0, 5, 10.0),      # K = 1, negative c
    (1, 3000, 5, 40.0, 3, 10.0),        # a single query row
    (300, 6000, 20, 40.0, 0, 1.0),      # Xavier-scale tables (epoch-0 evaluation), no mask
])
def test_tc_bit_exact_vs_oracle(T, ops, oracle, T_users, n_items, K, c, deg, scale):
    U, I, w, wu = make_model(T_users + n_items, T_users, n_items, scale=scale)
    mrp = mcol = None
    if deg:
        mrp, mcol = lists_to_csr(make_interactions(7, T_users, n_items, deg))
    sig_i, sig_u = oracle.score_gates(I, w), oracle.score_gates(U, wu)
    want_ids, want_sc = oracle.score_topk(U, I, sig_i, sig_u, c, mrp, mcol, K)
    dU, dI, gsi, gsu = _gates(T, ops, U, I, w, wu)
    stats = T.zeros(2, dtype=T.int64, device="cuda")
    ids, sc = ops.score_topk_tc(dU, dI, gsi, gsu, c, None if mrp is None else dev(T, mrp),
                                None if mcol is None else dev(T, mcol), K, stats=stats)
    np.testing.assert_array_equal(ids.cpu().numpy(), want_ids)
    np.testing.assert_array_equal(sc.cpu().numpy(), want_sc)
    # the tensor-core path did the work: nothing was handed to the exact fallback kernel
    assert int(stats[0].item()) == 0
    assert K <= stats[1].item() / T_users < 6 * K + 40


def test_tc_full_size_equals_exact_kernel(T, ops):
    """BASELINE configs[1] scoring shape (15 424 x 40 981): tcgen05 path == exact fp32 kernel."""
    T_users, n_items, K = 15424, 40981, 20
    U, I, w, wu = make_model(3, T_users, n_items, scale=10.0)
    rng = np.random.RandomState(4)
    cnt = np.maximum(1, rng.poisson(27, T_users))
    mrp = np.zeros(T_users + 1, np.int32)
    mrp[1:] = np.cumsum(cnt)
    mcol = np.concatenate([np.sort(rng.choice(n_items, size=k, replace=False)) for k in cnt]).astype(np.int32)
    dU, dI, gsi, gsu = _gates(T, ops, U, I, w, wu)
    dm, dc = dev(T, mrp), dev(T, mcol)
    stats = T.zeros(2, dtype=T.int64, device="cuda")
    ti, ts = ops.score_topk_tc(dU, dI, gsi, gsu, 40.0, dm, dc, K, stats=stats)
    ei, es = ops.score_topk_exact(dU, dI, gsi, gsu, 40.0, dm, dc, K)
    assert bool((ti == ei).all().item()) and bool((ts == es).all().item())
    assert int(stats[0].item()) == 0
    # no train item is ever emitted
    got = ti.cpu().numpy()
    for r in rng.randint(0, T_users, 50):
        assert not set(got[r].tolist()) & set(mcol[mrp[r]:mrp[r + 1]].tolist())
    # dispatcher: this shape goes to the tcgen05 path and gives the same bits
    di, ds = ops.score_topk(dU, dI, gsi, gsu, 40.0, dm, dc, K)
    assert bool((di == ei).all().item()) and bool((ds == es).all().item())


def test_tc_overflow_rows_fall_back_to_exact(T, ops, oracle):
    """degenerate inputs: all scores equal (every item is a candidate) and users whose train list
    covers a third of the catalogue -> those rows are re-done by the exact kernel, results exact."""
    K = 20
    Z, ZI = np.zeros((150, 64), np.float32), np.zeros((5000, 64), np.float32)
    h, hu = np.full(5000, 0.5, np.float32), np.full(150, 0.5, np.float32)
    want_ids, want_sc = oracle.score_topk(Z, ZI, h, hu, 40.0, None, None, K)
    stats = T.zeros(2, dtype=T.int64, device="cuda")
    ids, sc = ops.score_topk_tc(dev(T, Z), dev(T, ZI), dev(T, h), dev(T, hu), 40.0, None, None, K,
                                stats=stats)
    np.testing.assert_array_equal(ids.cpu().numpy(), want_ids)
    np.testing.assert_array_equal(sc.cpu().numpy(), want_sc)
    assert list(want_ids[0][:3]) == [0, 1, 2] and int(stats[0].item()) == 150

    T_users, n_items = 120, 9000
    U, I, w, wu = make_model(5, T_users, n_items, scale=10.0)
    rng = np.random.RandomState(5)
    lists = [np.sort(rng.choice(n_items, size=(3000 if u % 7 == 0 else 40), replace=False)).astype(np.int32)
             for u in range(T_users)]
    mrp, mcol = lists_to_csr(lists)
    sig_i, sig_u = oracle.score_gates(I, w), oracle.score_gates(U, wu)
    want_ids, want_sc = oracle.score_topk(U, I, sig_i, sig_u, 40.0, mrp, mcol, K)
    dU, dI, gsi, gsu = _gates(T, ops, U, I, w, wu)
    stats.zero_()
    ids, sc = ops.score_topk_tc(dU, dI, gsi, gsu, 40.0, dev(T, mrp), dev(T, mcol), K, stats=stats)
    np.testing.assert_array_equal(ids.cpu().numpy(), want_ids)
    np.testing.assert_array_equal(sc.cpu().numpy(), want_sc)
    assert 0 < int(stats[0].item()) <= 30


def test_tc_train_items_ranked_on_top(T, ops, oracle):
    """What a trained model looks like: every user's train items are exactly its top-scoring items.
    The filter pass masks them in place (they never reach the 128-slot candidate lists), rows with
    a train list longer than 1/16 of the catalogue go straight to the exact kernel, and the result
    stays bit-identical to the oracle."""
    T_users, n_items, K, c = 400, 6000, 20, 40.0
    U, I, w, wu = make_model(21, T_users, n_items, scale=10.0)
    sig_i, sig_u = oracle.score_gates(I, w), oracle.score_gates(U, wu)
    S = oracle.score_matrix(U, I, sig_i, sig_u, c)
    order = np.argsort(-S, axis=1, kind="stable")
    lens = np.array([5, 40, 150, 500])[np.arange(T_users) % 4]  # 500 > 6000/16 + 64: straight to exact
    lists = [np.sort(order[t, :lens[t]]).astype(np.int32) for t in range(T_users)]
    mrp, mcol = lists_to_csr(lists)
    want_ids, want_sc = oracle.score_topk(U, I, sig_i, sig_u, c, mrp, mcol, K)
    assert not set(want_ids[2].tolist()) & set(lists[2].tolist())
    dU, dI, gsi, gsu = _gates(T, ops, U, I, w, wu)
    stats = T.zeros(2, dtype=T.int64, device="cuda")
    ids, sc = ops.score_topk_tc(dU, dI, gsi, gsu, c, dev(T, mrp), dev(T, mcol), K, stats=stats)
    np.testing.assert_array_equal(ids.cpu().numpy(), want_ids)
    np.testing.assert_array_equal(sc.cpu().numpy(), want_sc)
    n_direct = int((lens == 500).sum())
    assert n_direct <= int(stats[0].item()) <= n_direct + 20
    # the lists held the top-K survivors, not the (up to 150) train items above them
    assert stats[1].item() / (T_users - stats[0].item()) < 6 * K + 40


def test_tc_item_shards_merge_to_single(T, ops):
    """item-sharded tcgen05 scoring + macr_topk_merge == unsharded exact kernel (the N-GPU path)."""
    T_users, n_items, K, G = 500, 20000, 20, 4
    U, I, w, wu = make_model(11, T_users, n_items, scale=10.0)
    mrp, mcol = lists_to_csr(make_interactions(12, T_users, n_items, 25))
    dU, dI, gsi, gsu = _gates(T, ops, U, I, w, wu)
    dm, dc = dev(T, mrp), dev(T, mcol)
    ei, es = ops.score_topk_exact(dU, dI, gsi, gsu, 40.0, dm, dc, K)
    bounds = np.linspace(0, n_items, G + 1).astype(int)
    pi, ps = [], []
    for g in range(G):
        lo, hi = int(bounds[g]), int(bounds[g + 1])
        i_, s_ = ops.score_topk_tc(dU, dI[lo:hi].contiguous(), gsi[lo:hi].contiguous(), gsu, 40.0,
                                   dm, dc, K, item_id_offset=lo)
        pi.append(i_)
        ps.append(s_)
    mi, ms = ops.topk_merge(T.stack(pi).contiguous(), T.stack(ps).contiguous())
    assert bool((mi == ei).all().item()) and bool((ms == es).all().item())


def test_tc_argument_errors(T, ops):
    from macr_b200._lib import MacrError

    U = T.zeros((4, 64), device="cuda")
    I = T.zeros((100, 64), device="cuda")
    s = T.zeros(100, device="cuda")
    with pytest.raises(MacrError):  # catalogue too small for the tensor-core path
        ops.score_topk_tc(U, I, s, s[:4], 40.0, None, None, 20)
    ids, _ = ops.score_topk(U, I, s, s[:4], 40.0, None, None, 20)  # dispatcher -> exact kernel
    assert ids.shape == (4, 20)


def test_tc_prepared_items_give_the_same_bits(T, ops):
    """`macr_score_tc_prepare_items` + `macr_score_topk_tc_prepared` (item operands prepared once
    per evaluation, query blocks of different sizes scored against them) == `macr_score_topk_tc`
    bit for bit; operands prepared from other items / another c are refused."""
    n_items, K, c = 6000, 20, 40.0
    U, I, w, wu = make_model(11, 700, n_items, scale=10.0)
    mrp, mcol = lists_to_csr(make_interactions(3, 700, n_items, 25))
    dU, dI, gsi, gsu = _gates(T, ops, U, I, w, wu)
    prep = ops.TcItems(dI, gsi, c)
    for lo, hi in ((0, 300), (300, 301), (301, 700)):   # three blocks, three plans
        m = (mrp[lo:hi + 1] - mrp[lo]).astype(np.int32)
        mc = mcol[mrp[lo]:mrp[hi]]
        args = (dU[lo:hi].contiguous(), dI, gsi, gsu[lo:hi].contiguous(), c, dev(T, m), dev(T, mc), K)
        want_i, want_s = ops.score_topk_tc(*args)
        stats = T.zeros(2, dtype=T.int64, device="cuda")
        got_i, got_s = ops.score_topk(*args, prepared=prep)
        assert bool((got_i == want_i).all().item()) and bool((got_s == want_s).all().item())
        got_i, got_s = ops.score_topk_tc(*args, stats=stats, prepared=prep)
        assert bool((got_i == want_i).all().item()) and int(stats[0].item()) == 0
    with pytest.raises(ops.MacrError):
        ops.score_topk_tc(dU, dI, gsi, gsu, c + 1.0, None, None, K, prepared=prep)
    with pytest.raises(ops.MacrError):
        ops.score_topk_tc(dU, dI.clone(), gsi, gsu, c, None, None, K, prepared=prep)
