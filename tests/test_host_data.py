"""Host data path vs golden vectors produced by the REFERENCE's own loaders / samplers
(tests/golden/make_golden.py ran macr_mf/load_data.py and macr_lightgcn/utility/load_data.py
from /root/reference).  Bit-exact: sampled triples, adjacency CSR structure and values."""
import hashlib
import json
import os
import random
import types

import numpy as np
import pytest

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
REF_DATA = "/root/reference/data"


def _mf_args(data_path, dataset, batch_size):
    return types.SimpleNamespace(data_path=data_path, dataset=dataset, batch_size=batch_size,
                                 data_type="ori", model="mf", source="normal", valid_set="test")


def _sha1(u, p, n):
    return hashlib.sha1(np.array([u, p, n], np.int64).tobytes()).hexdigest()


def test_mf_loader_and_sampler_tiny(capsys):
    from macr_b200.host.data_mf import Data

    g = np.load(os.path.join(GOLD, "mf_sampler.npz"))
    data = Data(_mf_args(GOLD + "/", "tiny", 16))
    assert (data.n_users, data.n_items, data.n_train, data.n_test) == tuple(
        int(g[k]) for k in ("n_users", "n_items", "n_train", "n_test"))
    random.seed(12345)
    np.testing.assert_array_equal(np.array(data.sample(), np.int64), g["b1"])
    np.testing.assert_array_equal(np.array(data.sample(), np.int64), g["b2"])
    data.batch_size = 200  # > n_users: users drawn with replacement (load_data.py:546-547)
    b3 = np.array(data.sample(), np.int64)
    np.testing.assert_array_equal(b3, g["b3"])
    # user 7 has no train line -> positive item 0 (load_data.py:552-553)
    assert np.all(b3[1][b3[0] == 7] == 0)
    # negatives never come from the user's train list
    for u, n in zip(b3[0], b3[2]):
        assert n not in data.train_user_list[u]


def test_lgcn_loader_sampler_adjacency_tiny():
    from macr_b200.host.data_lgcn import Data

    g = np.load(os.path.join(GOLD, "lgcn_sampler.npz"))
    data = Data(os.path.join(GOLD, "tiny"), 16, types.SimpleNamespace(valid_set="test"))
    assert (data.n_users, data.n_items, data.n_train, data.n_test) == tuple(
        int(g[k]) for k in ("n_users", "n_items", "n_train", "n_test"))
    np.testing.assert_array_equal(np.array(data.exist_users), g["exist_users"])  # file order
    random.seed(12345)
    np.random.seed(12345)
    np.testing.assert_array_equal(np.array(data.sample(), np.int64), g["b1"])
    np.testing.assert_array_equal(np.array(data.sample(), np.int64), g["b2"])
    a = np.load(os.path.join(GOLD, "lgcn_adj_tiny.npz"))
    rowptr, col, val = data.adj_csr("pre")
    np.testing.assert_array_equal(rowptr, a["indptr"])
    np.testing.assert_array_equal(col, a["indices"])
    np.testing.assert_array_equal(val, a["data"])  # float32 values bit for bit
    # symmetric, zero row for the user without interactions (inf -> 0)
    import scipy.sparse as sp

    M = sp.csr_matrix((val, col, rowptr), shape=tuple(a["shape"]))
    assert abs(M - M.T).max() == 0
    assert M[7].nnz == 0


def test_lgcn_adjacency_variants_tiny():
    """--adj_type plain | norm | gcmc (LightGCN.py:666-681): the other three members of
    get_adj_mat()'s 4-tuple (load_data.py:126-164) against the reference's own, bit for bit in the
    float32 the session feeds."""
    from macr_b200.host.data_lgcn import Data

    g = np.load(os.path.join(GOLD, "lgcn_adj_variants_tiny.npz"))
    data = Data(os.path.join(GOLD, "tiny"), 16, types.SimpleNamespace(valid_set="test"))
    for ref_name, adj_type in (("plain", "plain"), ("norm", "norm"), ("mean", "gcmc")):
        rowptr, col, val = data.adj_csr(adj_type)
        np.testing.assert_array_equal(rowptr, g[ref_name + "_indptr"], err_msg=adj_type)
        np.testing.assert_array_equal(col, g[ref_name + "_indices"], err_msg=adj_type)
        np.testing.assert_array_equal(val, g[ref_name + "_data"], err_msg=adj_type)
    # the default branch (any other --adj_type): mean + I
    rowptr, col, val = data.adj_csr("anything")
    n = len(rowptr) - 1
    assert len(val) == len(g["mean_data"]) + n


def test_lgcn_sample_test_matches_reference():
    """Data.sample_test() (load_data.py:213-257) against the reference's own code (run by
    make_golden.py with random.sample's pre-3.11 handling of dict keys): native sampler and the
    Python restatement, both bit-exact, two consecutive batches."""
    from macr_b200.host.data_lgcn import Data

    g = np.load(os.path.join(GOLD, "lgcn_sample_test.npz"))
    data = Data(os.path.join(GOLD, "tiny"), 16, types.SimpleNamespace(valid_set="test"))
    for fn in (data.sample_test, data.sample_test_py):
        random.seed(4321)
        np.random.seed(4321)
        np.testing.assert_array_equal(np.array(fn(), np.int64), g["t1"])
        np.testing.assert_array_equal(np.array(fn(), np.int64), g["t2"])


def test_lgcn_sample_test_runs_and_excludes_known_items():
    from macr_b200.host.data_lgcn import Data

    data = Data(os.path.join(GOLD, "tiny"), 8, types.SimpleNamespace(valid_set="test"))
    random.seed(1)
    np.random.seed(1)
    users, pos, neg = data.sample_test()
    assert len(users) == len(pos) == len(neg) == 8
    for u, p, n in zip(users, pos, neg):
        assert p in data.test_set[u]
        assert n not in data.test_set[u] and n not in data.train_items.get(u, [])


@pytest.mark.skipif(not os.path.isdir(os.path.join(REF_DATA, "addressa")),
                    reason="reference data set not mounted")
def test_addressa_digests(monkeypatch):
    """Real data: first sampled batch and the `pre` adjacency hash equal the reference's."""
    from macr_b200.host import data_lgcn, data_mf

    with open(os.path.join(GOLD, "digests.json")) as f:
        dg = json.load(f)
    monkeypatch.chdir("/root/reference")  # the MF loader reads ./data/<dataset>/
    d = data_mf.Data(_mf_args("./data/", "addressa", 1024))
    want = dg["mf_addressa"]
    assert (d.n_users, d.n_items, d.n_train, d.n_test, len(d.test_users)) == (
        want["n_users"], want["n_items"], want["n_train"], want["n_test"], want["n_test_users"])
    random.seed(12345)
    u, p, n = d.sample()
    assert _sha1(u, p, n) == want["first_batch_sha1"]

    g = data_lgcn.Data(os.path.join(REF_DATA, "addressa"), 1024, types.SimpleNamespace(valid_set="test"))
    want = dg["lgcn_addressa"]
    random.seed(12345)
    np.random.seed(12345)
    u, p, n = g.sample()
    assert _sha1(u, [int(x) for x in p], [int(x) for x in n]) == want["first_batch_sha1"]
    rowptr, col, val = g.adj_csr("pre")
    h = hashlib.sha1()
    for a in (rowptr, col, val):
        h.update(a.tobytes())
    assert int(val.size) == want["adj_nnz"]
    assert h.hexdigest() == want["adj_sha1"]


def test_flag_defaults_match_reference_parsers():
    from macr_b200.host import flags

    with open(os.path.join(GOLD, "parser_defaults.json")) as f:
        want = json.load(f)
    got_mf = vars(flags.parse_mf_args([]))
    got_lg = vars(flags.parse_lgcn_args([]))
    for k, v in want["mf"].items():
        assert got_mf[k] == v, k
    for k, v in want["lgcn"].items():
        assert got_lg[k] == v, k
    # README commands parse (README.md:52,82)
    a = flags.parse_mf_args("--dataset addressa --batch_size 1024 --cuda 0 --saveID 1 --log_interval 10 "
                            "--lr 0.001 --train rubibceboth --test rubi --c 40 --alpha 1e-3 --beta 1e-3".split())
    assert (a.train, a.test, a.c, a.batch_size) == ("rubibceboth", "rubi", 40.0, 1024)
    b = flags.parse_lgcn_args("--dataset addressa --batch_size 1024 --layer_size [64,64] --loss bceboth "
                              "--test rubiboth --c 40 --alpha 1e-2 --beta 1e-3 --epoch 2000".split())
    assert flags.as_list(b.layer_size) == [64, 64] and b.loss == "bceboth"


@pytest.mark.skipif(not os.path.isdir(os.path.join(REF_DATA, "addressa")), reason="reference data not mounted")
@pytest.mark.parametrize("batch_size", [1024, 8192, 20000])
def test_native_mf_sampler_equals_python_restatement(batch_size):
    """C sampler (macr_sample_mf) vs the line-by-line Python sampler on the same `random` stream:
    identical triples AND identical stream position afterwards.  1024: set-based random.sample,
    8192: pool-based (n <= setsize), 20000 > n_users: choice with replacement."""
    from macr_b200.host.data_mf import Data

    data = Data(_mf_args(REF_DATA + "/", "addressa", batch_size))
    for seed in (12345, 7):
        random.seed(seed)
        want = [data.sample_py() for _ in range(3)]
        tail_py = random.random()
        random.seed(seed)
        got = [data.sample() for _ in range(3)]
        assert random.random() == tail_py
        for (wu, wp, wn), (gu, gp, gn) in zip(want, got):
            assert wu == gu and wp == gp and wn == gn


@pytest.mark.skipif(not os.path.isdir(os.path.join(REF_DATA, "addressa")), reason="reference data not mounted")
def test_native_lgcn_samplers_equal_python_restatement():
    """macr_sample_lgcn vs the Python sampler: `random` (users) and legacy `np.random.randint`
    (items) streams, train sampler and the test-loss sampler, stream positions included."""
    from macr_b200.host.data_lgcn import Data

    data = Data(os.path.join(REF_DATA, "addressa"), 1024, types.SimpleNamespace(valid_set="test"))
    for fn_py, fn_c in ((data.sample_py, data.sample), (data.sample_test_py, data.sample_test)):
        random.seed(99)
        np.random.seed(99)
        want = [fn_py() for _ in range(3)]
        tails = (random.random(), np.random.randint(0, 1 << 30))
        random.seed(99)
        np.random.seed(99)
        got = [fn_c() for _ in range(3)]
        assert (random.random(), np.random.randint(0, 1 << 30)) == tails
        for w, g in zip(want, got):
            assert [list(map(int, x)) for x in w] == [list(x) for x in g]


def test_native_numpy_randint_twin_small_ranges():
    """legacy np.random.randint(0, n, size=1): n = 1 draws nothing, powers of two never reject."""
    from macr_b200.host import native_sampler as ns

    for n_items in (1, 2, 3, 64, 65, 1000):
        lists = {0: list(range(0))}  # user 0 bans nothing
        pos = ns.ListCSR({0: [5]}, 1)
        ban = ns.ListCSR(lists, 1)
        random.seed(1)
        np.random.seed(5)
        want = [int(np.random.randint(low=0, high=1, size=1)[0]) * 0 + int(
            np.random.randint(low=0, high=n_items, size=1)[0]) for _ in range(50)]
        tail = np.random.randint(0, 1 << 30)
        np.random.seed(5)
        got = []
        for _ in range(50):
            _, p, n = ns.sample_lgcn(np.zeros(1, np.int32), 1, n_items, pos, ban, 1)
            assert p[0] == 5
            got.append(int(n[0]))
        assert got == want and np.random.randint(0, 1 << 30) == tail


def test_sample_epoch_equals_consecutive_samples():
    from macr_b200.host.data_lgcn import Data as LData
    from macr_b200.host.data_mf import Data

    data = Data(_mf_args(GOLD + "/", "tiny", 16))
    random.seed(3)
    want = np.array([data.sample() for _ in range(5)], np.int32)
    tail = random.random()
    random.seed(3)
    np.testing.assert_array_equal(data.sample_epoch(5), want)
    assert random.random() == tail
    ld = LData(os.path.join(GOLD, "tiny"), 16, types.SimpleNamespace(valid_set="test"))
    random.seed(4)
    np.random.seed(4)
    want = np.array([ld.sample() for _ in range(5)], np.int32)
    random.seed(4)
    np.random.seed(4)
    np.testing.assert_array_equal(ld.sample_epoch(5), want)


def _states():
    return random.getstate(), np.random.get_state()


def _same_states(a, b):
    return a[0] == b[0] and a[1][0] == b[1][0] and np.array_equal(a[1][1], b[1][1]) and a[1][2:] == b[1][2:]


@pytest.mark.parametrize("crowded", [False, True])
def test_epoch_samplers_redraw_after_rejected_candidates(crowded):
    """The epoch samplers draw chunks speculatively and go back in the word stream when a
    candidate turns out to be a list member.  Dense lists (every second draw rejected) make that
    the common case; triples AND generator states must still equal the per-batch samplers',
    whatever the chunking, the position inside the 624-word block, or the crowding of the hashed
    pair set (`crowded`: 2 buckets for ~2000 pairs -> every look-up settles on the exact list)."""
    from macr_b200.host import native_sampler as ns

    rng = np.random.RandomState(11)
    n_users, n_items = 70, 61
    lists = {u: rng.permutation(n_items)[:rng.randint(1, 50)].tolist() for u in range(n_users)}
    lists[5] = []          # MF: a user without train items gets positive 0, no draw
    both = {u: sorted(set(v) | set(rng.randint(0, n_items, 5).tolist())) for u, v in lists.items()}
    csr, ban = ns.ListCSR(lists, n_users), ns.ListCSR(both, n_users)
    if crowded:
        csr.pair_tags(log2=1), ban.pair_tags(log2=1)
    pop = np.arange(n_users, dtype=np.int32)
    pop_l = np.array([u for u in range(n_users) if lists[u]], np.int32)
    for B, n_b, skip in ((16, 9, 0), (64, 5, 623), (300, 3, 7), (1, 40, 0)):
        random.seed(B)
        np.random.seed(B)
        for _ in range(skip):        # start somewhere inside a block / right at its end
            random.getrandbits(32), np.random.randint(0, 1 << 30)
        start = _states()
        want = np.array([ns.sample_mf(pop, n_users, n_items, csr, B) for _ in range(n_b)])
        end = _states()
        random.setstate(start[0]), np.random.set_state(start[1])
        got = np.concatenate([ns.sample_mf_epoch(pop, n_users, n_items, csr, B, k)
                              for k in (n_b - 2, 0, 2)])
        np.testing.assert_array_equal(got, want)
        assert _same_states(_states(), end)
        # LightGCN: two streams, numpy's masked rejection, ban list != positive list
        random.setstate(start[0]), np.random.set_state(start[1])
        want = np.array([ns.sample_lgcn(pop_l, n_users, n_items, csr, ban, B) for _ in range(n_b)])
        end = _states()
        random.setstate(start[0]), np.random.set_state(start[1])
        got = np.concatenate([ns.sample_lgcn_epoch(pop_l, n_users, n_items, csr, ban, B, k)
                              for k in (1, n_b - 1)])
        np.testing.assert_array_equal(got, want)
        assert _same_states(_states(), end)
        for u, n in zip(got[:, 0].ravel(), got[:, 2].ravel()):
            assert n not in both[u]


@pytest.mark.skipif(not os.path.isdir(os.path.join(REF_DATA, "addressa")), reason="reference data not mounted")
def test_epoch_samplers_equal_per_batch_on_addressa():
    """Real lists (addressa, B=1024, one epoch = 111 batches): the epoch forms reproduce the
    per-batch twins (themselves pinned to the reference's sampler by sha1) and leave both
    generators where they do."""
    from macr_b200.host.data_lgcn import Data as LData
    from macr_b200.host.data_mf import Data

    data = Data(_mf_args(REF_DATA + "/", "addressa", 1024))
    n_b = data.n_train // 1024 + 1
    random.seed(12345)
    want = np.array([data.sample() for _ in range(n_b)], np.int32)
    end = random.getstate()
    random.seed(12345)
    np.testing.assert_array_equal(data.sample_epoch(n_b), want)
    assert random.getstate() == end
    ld = LData(os.path.join(REF_DATA, "addressa"), 1024, types.SimpleNamespace(valid_set="test"))
    for one, many in ((ld.sample, ld.sample_epoch), (ld.sample_test, ld.sample_test_epoch)):
        random.seed(7), np.random.seed(7)
        want = np.array([one() for _ in range(20)], np.int32)
        end = _states()
        random.seed(7), np.random.seed(7)
        np.testing.assert_array_equal(many(20), want)
        assert _same_states(_states(), end)


def test_sampler_word_stream_rewind_rebuilds_the_generator_state():
    """The epoch samplers read MT19937 output from a buffer they may re-read from an earlier
    position.  When the final position lies in a block older than the newest generated one, the
    state handed back to `random` is rebuilt from the buffered OUTPUT words (inverse tempering);
    at a block boundary the position is reported lazily (index 624 of the finished block), as
    CPython leaves it.  Checked against `random` itself, word for word."""
    import ctypes as C

    from macr_b200._lib import lib

    fn = lib().macr_sampler_stream_selftest
    fn.argtypes = [C.c_void_p, C.c_int64, C.c_int64, C.c_void_p]
    fn.restype = C.c_int
    for skip, advance, back in ((0, 10, 3), (5, 700, 200), (5, 1300, 1290), (100, 2000, 0), (100, 2000, 1376),
                                (7, 617, 0), (7, 617 + 624, 624), (7, 3000, 3000 - 617), (623, 1, 0), (623, 700, 699),
                                (0, 624 * 3, 624), (0, 0, 0)):
        random.seed(skip * 7 + advance)
        for _ in range(skip):
            random.getrandbits(32)
        version, internal, gauss = random.getstate()
        st = np.array(internal, dtype=np.uint32)
        x = np.zeros(1, np.uint32)
        assert fn(C.c_void_p(st.ctypes.data), advance, back, C.c_void_p(x.ctypes.data)) == 0
        want_x = 0
        start = (version, internal, gauss)
        for _ in range(advance):
            want_x ^= random.getrandbits(32)
        assert int(x[0]) == want_x
        random.setstate(start)
        for _ in range(advance - back):
            random.getrandbits(32)
        assert tuple(int(v) for v in st) == random.getstate()[1], (skip, advance, back)


def test_epoch_samplers_small_ranges_and_single_item_lists():
    """Word-driven draw loops vs the draw-by-draw twins where a draw may consume NO word (numpy's
    randint over one value; an MF user without train items) or a power-of-two range never rejects:
    catalogues of 1, 2, 3, 64, 65 items, lists of one item."""
    from macr_b200.host import native_sampler as ns

    n_users = 40
    pop = np.arange(n_users, dtype=np.int32)
    for n_items in (1, 2, 3, 64, 65):
        rng = np.random.RandomState(n_items)
        pos_lists = {u: rng.randint(0, n_items, rng.randint(1, 4)).tolist() for u in range(n_users)}
        pos_lists[3] = [0]
        ban_lists = {u: ([] if n_items < 3 else sorted(set(rng.randint(0, n_items, 2).tolist()) - {n_items - 1}))
                     for u in range(n_users)}
        pos, ban = ns.ListCSR(pos_lists, n_users), ns.ListCSR(ban_lists, n_users)
        mf_lists = dict(ban_lists)
        mf_lists[7] = []
        mf = ns.ListCSR(mf_lists, n_users)
        for B in (8, 40, 100):
            random.seed(n_items * 31 + B), np.random.seed(n_items * 31 + B)
            start = _states()
            want = np.array([ns.sample_lgcn(pop, n_users, n_items, pos, ban, B) for _ in range(6)])
            end = _states()
            random.setstate(start[0]), np.random.set_state(start[1])
            np.testing.assert_array_equal(ns.sample_lgcn_epoch(pop, n_users, n_items, pos, ban, B, 6), want)
            assert _same_states(_states(), end)
            random.setstate(start[0])
            want = np.array([ns.sample_mf(pop, n_users, n_items, mf, B) for _ in range(6)])
            end = random.getstate()
            random.setstate(start[0])
            np.testing.assert_array_equal(ns.sample_mf_epoch(pop, n_users, n_items, mf, B, 6), want)
            assert random.getstate() == end


@pytest.mark.parametrize("dense", [True, False])
def test_epoch_sampler_verifier_thread_equals_single_thread(dense, monkeypatch):
    """MACR_SAMPLER_THREADS=2 puts the verification pass on a second thread that runs behind the
    speculative draws and reports rejected candidates back (the drawing thread then rewinds).  Same
    triples and generator states as the single-threaded epoch form and as the per-batch twins --
    with dense lists (a rewind every other triple) and sparse ones (long speculative runs)."""
    from macr_b200.host import native_sampler as ns

    rng = np.random.RandomState(5)
    n_users, n_items = (300, 90) if dense else (1500, 30000)
    lists = {u: rng.choice(n_items, size=rng.randint(2, 40), replace=False).tolist() for u in range(n_users)}
    csr = ns.ListCSR(lists, n_users)
    pop = np.arange(n_users, dtype=np.int32)
    B, n_b = 256, 24
    results = []
    for mode in ("1", "2", "2"):
        monkeypatch.setenv("MACR_SAMPLER_THREADS", mode)
        random.seed(99), np.random.seed(99)
        a = ns.sample_mf_epoch(pop, n_users, n_items, csr, B, n_b)
        b = ns.sample_lgcn_epoch(pop, n_users, n_items, csr, csr, B, n_b)
        results.append((a, b, _states()))
    monkeypatch.delenv("MACR_SAMPLER_THREADS")
    random.seed(99), np.random.seed(99)
    want_a = np.array([ns.sample_mf(pop, n_users, n_items, csr, B) for _ in range(n_b)])
    want_b = np.array([ns.sample_lgcn(pop, n_users, n_items, csr, csr, B) for _ in range(n_b)])
    want_state = _states()
    for a, b, st in results:
        np.testing.assert_array_equal(a, want_a)
        np.testing.assert_array_equal(b, want_b)
        assert _same_states(st, want_state)
