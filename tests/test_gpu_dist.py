"""GPU test of the row-sharded MF trainer (macr_b200/host/dist.py:RowShardedMFTrainer): two ranks
share cuda:0 (gloo moves the CUDA exchange buffer, so one GPU is enough), each owns half of the
rows of both tables and their Adam slots, and after several steps the owned slices, w, w_user and
the losses are BIT-identical to a single-GPU `MFTrainer` run on the full tables (SURVEY 8e rows
"dense Adam sweep" / "gather + grid + row grads").  On a multi-GPU box the same class runs one
rank per GPU over NCCL (dev/bench_sharded_train.py)."""
import os
import socket

import numpy as np
import pytest
import torch.distributed as dist
import torch.multiprocessing as mp

from helpers import make_batch, make_model

pytestmark = pytest.mark.gpu

N_USERS, N_ITEMS, B, STEPS = 201, 745, 256, 5  # ragged shards; B > N_USERS: users repeat inside a batch
HP = dict(lr=1e-2, alpha=1e-2, beta=1e-3, decay=1e-4, batch_size=B)


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _batches():
    rng = np.random.RandomState(72)
    return [make_batch(rng, N_USERS, N_ITEMS, B) for _ in range(STEPS)]


def _worker(rank, world, port, out_dir):
    import torch

    from macr_b200 import ops
    from macr_b200.host import dist as mdist

    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        U, I, w, wu = make_model(71, N_USERS, N_ITEMS, scale=4.0)
        sh = mdist.RowShardedMFTrainer(U, I, w, wu, ops.HParams.make(**HP), B, rank=rank, world=world,
                                       device="cuda:0")
        losses = []
        for u, p, n in _batches():
            ids = [torch.from_numpy(np.asarray(x, np.int32)).cuda() for x in (u, p, n)]
            losses.append(sh.step_device(*ids).cpu().numpy().copy())
        out = {k: v.cpu().numpy() for k, v in sh.local_tables().items()}
        out["losses"] = np.stack(losses)
        out["bounds"] = np.array([sh.u_lo, sh.u_hi, sh.i_lo, sh.i_hi])
        np.savez(os.path.join(out_dir, f"rank{rank}.npz"), **out)
        sh.close()
    finally:
        dist.destroy_process_group()


def test_row_sharded_mf_training_equals_single_gpu(tmp_path):
    import torch

    from macr_b200 import ops

    world = 2
    mp.spawn(_worker, args=(world, _free_port(), str(tmp_path)), nprocs=world, join=True)
    U, I, w, wu = make_model(71, N_USERS, N_ITEMS, scale=4.0)
    tr = ops.MFTrainer(U, I, w, wu, ops.HParams.make(**HP), max_batch=B)
    want_losses = []
    for u, p, n in _batches():
        ids = [torch.from_numpy(np.asarray(x, np.int32)).cuda() for x in (u, p, n)]
        want_losses.append(tr.step_device(*ids).cpu().numpy().copy())
    t = tr.tab
    full = {k: getattr(t, k).cpu().numpy() for k in ("U", "mU", "vU", "I", "mI", "vI", "w", "wu")}
    covered_u = covered_i = 0
    for r in range(world):
        z = np.load(tmp_path / f"rank{r}.npz")
        u_lo, u_hi, i_lo, i_hi = (int(x) for x in z["bounds"])
        np.testing.assert_array_equal(z["losses"], np.stack(want_losses))
        for k in ("U", "mU", "vU"):
            np.testing.assert_array_equal(z[k], full[k][u_lo:u_hi], err_msg=f"rank {r} {k}")
        for k in ("I", "mI", "vI"):
            np.testing.assert_array_equal(z[k], full[k][i_lo:i_hi], err_msg=f"rank {r} {k}")
        np.testing.assert_array_equal(z["w"], full["w"])
        np.testing.assert_array_equal(z["wu"], full["wu"])
        covered_u += u_hi - u_lo
        covered_i += i_hi - i_lo
    assert covered_u == N_USERS and covered_i == N_ITEMS
    assert np.abs(full["U"] - U).max() > 0
    tr.close()


def test_row_partitioned_lightgcn_world1_equals_plain_trainer():
    """world = 1: the row-partitioned code path (range-restricted segment plan, owned-range Adam,
    the forward helper) on one GPU must reproduce the plain LightGCN trainer bit for bit."""
    import torch

    from helpers import make_interactions, norm_adj_csr
    from macr_b200 import ops
    from macr_b200.host import dist as mdist

    n_users, n_items, Bq, L = 601, 257, 128, 2
    lists = make_interactions(3, n_users, n_items, 40)
    rowptr, col, val = norm_adj_csr(lists, n_users, n_items)
    U, I, w, wu = make_model(9, n_users, n_items, scale=4.0)
    hp = dict(lr=1e-3, alpha=1e-2, beta=1e-3, decay=1e-4, batch_size=Bq)
    rng = np.random.RandomState(6)
    batches = np.stack([np.stack(make_batch(rng, n_users, n_items, Bq)) for _ in range(3)]).astype(np.int32)
    db = torch.from_numpy(batches).cuda()
    tr = ops.LGCNTrainer(rowptr, col, val, U, I, w, wu, L, ops.HParams.make(**hp), max_batch=Bq)
    sh = mdist.RowShardedLGCNTrainer(rowptr, col, val, U, I, w, wu, L, ops.HParams.make(**hp), Bq, rank=0, world=1)
    np.testing.assert_array_equal(sh.run(db).cpu().numpy(), tr.run(db).cpu().numpy())
    for k, v in sh.local_tables().items():
        np.testing.assert_array_equal(v.cpu().numpy(), getattr(tr.tab, k).cpu().numpy(), err_msg=k)
    for a, b in zip(sh.embeddings(), tr.embeddings()):
        np.testing.assert_array_equal(a.cpu().numpy(), b.cpu().numpy())
    sh.check_peers()
    sh.close()
    tr.close()


def test_topk_merge_peers_kernel_on_one_gpu(oracle):
    """`macr_topk_merge_peers` (the fused exchange + merge of item-sharded scoring) with three
    "ranks" whose buffers all live on this GPU: every rank's row block, merged from all shards'
    candidate lists and stored into every rank's result buffer, must equal the oracle's merge of
    the [G,T,K] block -- the same kernel then runs over NVLink peer mappings (test_gpu_multi.py)."""
    import ctypes as C

    import torch

    from macr_b200 import ops

    G, T, K = 3, 77, 20
    rng = np.random.RandomState(3)
    ids = np.full((G, T, K), -1, np.int32)
    sc = np.full((G, T, K), -np.inf, np.float32)
    for g in range(G):
        for t in range(T):
            n = int(rng.randint(0, K + 1))  # short lists are padded with -1 / -inf
            cand = rng.choice(np.arange(g * 1000, (g + 1) * 1000), size=n, replace=False)
            s_ = np.round(rng.randn(n), 1).astype(np.float32)  # coarse scores: ties across shards
            order = np.lexsort((cand, -s_))
            ids[g, t, :n], sc[g, t, :n] = cand[order], s_[order]
    want_i, want_s = oracle.topk_merge(ids.copy(), sc.copy())
    d_ids = [torch.from_numpy(ids[g].copy()).cuda() for g in range(G)]
    d_sc = [torch.from_numpy(sc[g].copy()).cuda() for g in range(G)]
    o_ids = [torch.full((T, K), -7, dtype=torch.int32, device="cuda") for _ in range(G)]
    o_sc = [torch.zeros((T, K), dtype=torch.float32, device="cuda") for _ in range(G)]
    tab = lambda ts: (C.c_void_p * G)(*[t.data_ptr() for t in ts])
    Tb = (T + G - 1) // G
    for r in range(G):  # "rank" r merges its row block and stores it into every rank's result
        row0 = min(T, r * Tb)
        ops.topk_merge_peers(tab(d_ids), tab(d_sc), tab(o_ids), tab(o_sc), G, K, row0, min(T, row0 + Tb) - row0)
    for r in range(G):
        np.testing.assert_array_equal(o_ids[r].cpu().numpy(), want_i)
        np.testing.assert_array_equal(o_sc[r].cpu().numpy(), want_s)
