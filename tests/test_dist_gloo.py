"""world_size-2 gloo test (CPU) of the multi-GPU scoring plumbing: shard bounds, id offsets, the
all-gather layout [G][T][K], and that shard-then-merge equals the unsharded ranking.  The
per-shard scoring and the merge are done by the CPU oracle here (the checker); on the GPU the
same plumbing feeds macr_score_topk / macr_topk_merge (tests/test_gpu_kernels.py,
bench.py --gpus N)."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from helpers import lists_to_csr, make_interactions, make_model


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, out_dir):
    import oracle
    from macr_b200.host import dist as mdist

    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        n_users, n_items, T, K = 120, 1001, 63, 20  # odd T: ragged user slices
        U, I, w, wu = make_model(3, n_users, n_items, scale=8.0)
        q = np.random.RandomState(4).permutation(n_users)[:T]
        Uq = np.ascontiguousarray(U[q])
        lists = make_interactions(5, T, n_items, 25)
        mrp, mcol = lists_to_csr(lists)
        sig_u = oracle.score_gates(Uq, wu)
        b = mdist.item_shard_bounds(n_items, world)
        assert b[0] == 0 and b[-1] == n_items and np.all(np.diff(b) > 0)
        lo, hi = int(b[rank]), int(b[rank + 1])
        It = np.ascontiguousarray(I[lo:hi])
        ids, sc = oracle.score_topk(Uq, It, oracle.score_gates(It, w), sig_u, 40.0, mrp, mcol, K,
                                    item_id_offset=lo)
        assert ids.min() >= lo and ids.max() < hi
        gi, gs = mdist.all_gather_candidates(torch.from_numpy(ids), torch.from_numpy(sc))
        assert gi.shape == (world, T, K)
        np.testing.assert_array_equal(gi[rank].numpy(), ids)  # shard order = rank order
        mi, ms = oracle.topk_merge(gi.numpy().copy(), gs.numpy().copy())
        want_i, want_s = oracle.score_topk(Uq, I, oracle.score_gates(I, w), sig_u, 40.0, mrp, mcol, K)
        np.testing.assert_array_equal(mi, want_i)
        np.testing.assert_array_equal(ms, want_s)
        # the product's exchange: all-to-all of row blocks, block merge, all-gather (odd T: one padding row)
        class _MergeOps:
            @staticmethod
            def topk_merge(gi_, gs_):
                a, b_ = oracle.topk_merge(gi_.numpy().copy(), gs_.numpy().copy())
                return torch.from_numpy(a), torch.from_numpy(b_)

        bi, bs = mdist.merge_shard_candidates(_MergeOps, torch.from_numpy(ids), torch.from_numpy(sc), world)
        np.testing.assert_array_equal(bi.numpy(), want_i)
        np.testing.assert_array_equal(bs.numpy(), want_s)
        # user-partitioned layout: ragged row slices -> the full [T,K] result on every rank
        ub = mdist.user_shard_bounds(T, world)
        ulo, uhi = int(ub[rank]), int(ub[rank + 1])
        sub_rp = (mrp[ulo:uhi + 1]).copy()  # absolute offsets into mcol stay valid
        li, lsc = oracle.score_topk(np.ascontiguousarray(Uq[ulo:uhi]), I, oracle.score_gates(I, w),
                                    np.ascontiguousarray(sig_u[ulo:uhi]), 40.0, sub_rp, mcol, K)
        full_i = mdist.all_gather_rows(torch.from_numpy(li), T, world, rank)
        full_s = mdist.all_gather_rows(torch.from_numpy(lsc), T, world, rank)
        np.testing.assert_array_equal(full_i.numpy(), want_i)
        np.testing.assert_array_equal(full_s.numpy(), want_s)
        # every rank ends with the same global list
        chk = torch.tensor([int(mi.astype(np.int64).sum())])
        both = [torch.zeros_like(chk) for _ in range(world)]
        dist.all_gather(both, chk)
        assert len({int(x) for x in both}) == 1
        open(os.path.join(out_dir, f"ok{rank}"), "w").write("ok")
    finally:
        dist.destroy_process_group()


def test_sharded_scoring_plumbing_world2(tmp_path, oracle):
    world = 2
    mp.spawn(_worker, args=(world, _free_port(), str(tmp_path)), nprocs=world, join=True)
    assert all(os.path.exists(tmp_path / f"ok{r}") for r in range(world))


def test_shard_bounds_cover_everything():
    from macr_b200.host.dist import item_shard_bounds

    for n, g in ((8790, 8), (40981, 8), (7, 8), (1_000_000, 4)):
        b = item_shard_bounds(n, g)
        assert b[0] == 0 and b[-1] == n and len(b) == g + 1 and np.all(np.diff(b) >= 0)


# ------------------------------------------------------------------------------------------------
# row-sharded MF training: ownership, ghost slots, the one all-reduce -- with the CPU oracle as the
# step engine (the checker) in place of macr_mf_trainer_step
# ------------------------------------------------------------------------------------------------
class _OracleOps:
    """Stand-in for `macr_b200.ops` on the CPU: same surface as far as RowShardedMFTrainer uses it
    (the oracle as the step engine, a torch restatement of csrc/shard.cu's pack / unpack)."""

    from macr_b200._lib import ShardDesc  # plain ctypes struct, no GPU needed

    class MFTrainer:
        def __init__(self, U, I, w, wu, hp, max_batch, device="cpu"):
            import oracle

            self.oracle, self.hp, self.dev = oracle, hp, torch.device("cpu")
            self.st = oracle.MFState(U.numpy(), I.numpy(), w, wu)

            class Tab:
                pass

            self.tab = Tab()
            for k in ("U", "mU", "vU", "I", "mI", "vI", "w", "wu"):
                setattr(self.tab, k, torch.from_numpy(getattr(self.st, k)))  # shares memory

        def step_device(self, u, p, n):
            return torch.from_numpy(self.oracle.mf_step(self.st, u.numpy(), p.numpy(), n.numpy(), self.hp))

        def close(self):
            pass

    @staticmethod
    def shard_table(rows, device, peer_mappable=False):
        return torch.zeros((rows, 64), dtype=torch.float32), None

    @staticmethod
    def _slots(desc, B, parity):
        q = torch.arange(3 * B)
        return torch.where(q < B, parity * desc.max_batch + q, parity * 2 * desc.max_batch + q - B)

    @classmethod
    def shard_pack(cls, U, I, desc, ids3, B, parity, local3, ex):
        ids = ids3.long()
        lo = torch.where(torch.arange(3 * B) < B, desc.u_lo, desc.i_lo)
        hi = torch.where(torch.arange(3 * B) < B, desc.u_hi, desc.i_hi)
        own = (ids >= lo) & (ids < hi)
        local3[: 3 * B] = torch.where(own, ids - lo, (hi - lo) + cls._slots(desc, B, parity)).int()
        rows = torch.cat([U[(ids[:B] - desc.u_lo).clamp(0, U.shape[0] - 1)],
                          I[(ids[B:] - desc.i_lo).clamp(0, I.shape[0] - 1)]])
        ex.copy_(torch.where(own[:, None], rows, torch.zeros(())))

    @classmethod
    def shard_unpack(cls, U, I, desc, ex, B, parity):
        n_lu, n_li, mb = desc.u_hi - desc.u_lo, desc.i_hi - desc.i_lo, desc.max_batch
        U[n_lu + parity * mb: n_lu + parity * mb + B] = ex[:B]
        I[n_li + parity * 2 * mb: n_li + parity * 2 * mb + 2 * B] = ex[B:]


def _train_worker(rank, world, port, out_dir):
    import oracle
    from helpers import make_batch
    from macr_b200.host import dist as mdist

    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        n_users, n_items, B, steps = 101, 77, 96, 4  # odd sizes: ragged shards; B > n_items: duplicates
        U, I, w, wu = make_model(7, n_users, n_items, scale=4.0)
        hp = oracle.HParams.make(lr=1e-2, alpha=1e-2, beta=1e-3, decay=1e-3, batch_size=B)
        rng = np.random.RandomState(8)
        batches = [make_batch(rng, n_users, n_items, B) for _ in range(steps)]
        sh = mdist.RowShardedMFTrainer(U, I, w, wu, hp, B, rank=rank, world=world, device="cpu",
                                       ops_module=_OracleOps)
        single = oracle.MFState(U, I, w, wu)
        for u, p, n in batches:
            want = oracle.mf_step(single, u, p, n, hp)
            got = sh.step_device(*(torch.from_numpy(np.asarray(x, np.int32)) for x in (u, p, n)))
            np.testing.assert_array_equal(got.numpy(), want)  # replicated part: identical everywhere
        loc = sh.local_tables()
        for k, full, lo, hi in (("U", single.U, sh.u_lo, sh.u_hi), ("mU", single.mU, sh.u_lo, sh.u_hi),
                                ("vU", single.vU, sh.u_lo, sh.u_hi), ("I", single.I, sh.i_lo, sh.i_hi),
                                ("mI", single.mI, sh.i_lo, sh.i_hi), ("vI", single.vI, sh.i_lo, sh.i_hi)):
            np.testing.assert_array_equal(loc[k].numpy(), full[lo:hi], err_msg=k)
        np.testing.assert_array_equal(loc["w"].numpy(), single.w)
        np.testing.assert_array_equal(loc["wu"].numpy(), single.wu)
        assert (sh.u_hi - sh.u_lo) in (50, 51) and sh.trainer.tab.U.shape[0] == sh.n_lu + 2 * B
        open(os.path.join(out_dir, f"train_ok{rank}"), "w").write("ok")
    finally:
        dist.destroy_process_group()


def test_row_sharded_training_plumbing_world2(tmp_path, oracle):
    world = 2
    mp.spawn(_train_worker, args=(world, _free_port(), str(tmp_path)), nprocs=world, join=True)
    assert all(os.path.exists(tmp_path / f"train_ok{r}") for r in range(world))


def test_row_sharded_trainer_world1_needs_no_process_group(oracle):
    """world = 1: no collective, the whole tables are local, ghost rows exist but are never read."""
    from helpers import make_batch
    from macr_b200.host import dist as mdist

    n_users, n_items, B = 37, 29, 48
    U, I, w, wu = make_model(17, n_users, n_items, scale=4.0)
    hp = oracle.HParams.make(lr=1e-2, alpha=1e-2, beta=1e-3, decay=1e-3, batch_size=B)
    sh = mdist.RowShardedMFTrainer(U, I, w, wu, hp, B, rank=0, world=1, device="cpu", ops_module=_OracleOps)
    single = oracle.MFState(U, I, w, wu)
    rng = np.random.RandomState(18)
    for _ in range(3):
        u, p, n = make_batch(rng, n_users, n_items, B)
        want = oracle.mf_step(single, u, p, n, hp)
        got = sh.step_device(*(torch.from_numpy(np.asarray(x, np.int32)) for x in (u, p, n)))
        np.testing.assert_array_equal(got.numpy(), want)
    loc = sh.local_tables()
    np.testing.assert_array_equal(loc["U"].numpy(), single.U)
    np.testing.assert_array_equal(loc["vI"].numpy(), single.vI)
    assert (sh.u_lo, sh.u_hi, sh.i_lo, sh.i_hi) == (0, n_users, 0, n_items)


def test_partition_adjacency_rows_cover_the_graph_exactly_once():
    """SURVEY 8e row 4 host logic: the per-rank CSRs hold disjoint row sets whose union is the
    adjacency; row pointers stay N+1 long, column ids stay global."""
    import scipy.sparse as sp

    from macr_b200.host import dist as mdist

    n_users, n_items = 37, 23
    rng = np.random.RandomState(1)
    R = sp.random(n_users, n_items, density=0.2, random_state=rng, format="csr", dtype=np.float32)
    A = sp.bmat([[None, R], [R.T, None]], format="csr", dtype=np.float32)
    A.sort_indices()
    world = 3
    ub, ib = mdist.user_shard_bounds(n_users, world), mdist.item_shard_bounds(n_items, world)
    total = sp.csr_matrix(A.shape, dtype=np.float32)
    for r in range(world):
        rp, cl, vl = mdist.partition_adjacency(A.indptr, A.indices, A.data, n_users, int(ub[r]), int(ub[r + 1]),
                                               int(ib[r]), int(ib[r + 1]))
        assert rp.dtype == np.int32 and len(rp) == A.shape[0] + 1 and rp[-1] == len(cl) == len(vl)
        part = sp.csr_matrix((vl, cl, rp), shape=A.shape)
        own = np.zeros(A.shape[0], bool)
        own[ub[r]:ub[r + 1]] = True
        own[n_users + ib[r]:n_users + ib[r + 1]] = True
        assert np.all(np.diff(rp)[~own] == 0)
        np.testing.assert_array_equal(part[own].toarray(), A[own].toarray())
        total = total + part
    np.testing.assert_array_equal(total.toarray(), A.toarray())


def test_balanced_bounds_split_weight_not_rows():
    from macr_b200.host import dist as mdist

    w = np.array([1000, 10, 10, 10, 500, 1, 1, 1, 1, 466], np.int64)  # total 2000
    b = mdist.balanced_bounds(w, 4)
    assert b[0] == 0 and b[-1] == len(w) and np.all(np.diff(b) >= 0)
    loads = [int(w[b[r]:b[r + 1]].sum()) for r in range(4)]
    assert sum(loads) == 2000 and max(loads) <= 1000  # the heavy head row sits alone
    assert list(mdist.balanced_bounds(np.ones(12, np.int64), 3)) == [0, 4, 8, 12]
    assert list(mdist.balanced_bounds(np.ones(5, np.int64), 1)) == [0, 5]
