"""Multi-GPU parity (skipped unless the box has >= 2 GPUs): one rank per GPU over NCCL.

* `RowShardedMFTrainer` with both exchange transports -- the NCCL all-reduce of the packed rows and
  the fused NVLink push (peer stores into the ghost rows + flag barrier, csrc/shard.cu) -- must
  leave owned slices, w, w_user and losses BIT-identical to a single-GPU `MFTrainer` run
  (SURVEY 8e rows "dense Adam sweep" / "gather + grid + row grads").
* item-partitioned and user-partitioned full-catalogue scoring (tcgen05 path) must return exactly
  the ids and scores of the unsharded call (SURVEY 8e row "full-catalogue scoring + top-K").
"""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from helpers import make_batch, make_model

pytestmark = pytest.mark.gpu

N_USERS, N_ITEMS, B, STEPS = 2001, 4745, 512, 6
HP = dict(lr=1e-2, alpha=1e-2, beta=1e-3, decay=1e-4, batch_size=B)


def _spawn(fn, args, world, out_dir):
    """mp.spawn with the workers' tracebacks kept: a failing rank writes `fail_rank<r>.txt`."""
    try:
        mp.spawn(_guard, args=(fn, out_dir) + args, nprocs=world, join=True)
    except Exception as e:  # noqa: BLE001
        logs = "".join(open(os.path.join(out_dir, f)).read() for f in sorted(os.listdir(out_dir))
                       if f.startswith("fail_rank"))
        raise AssertionError(f"a worker failed: {e}\n{logs}") from e


def _guard(rank, fn, out_dir, *args):
    import traceback

    try:
        fn(rank, *args)
    except BaseException:
        with open(os.path.join(out_dir, f"fail_rank{rank}.txt"), "w") as f:
            f.write(f"--- rank {rank} ---\n" + traceback.format_exc())
        raise


def _n_gpus():
    return torch.cuda.device_count() if torch.cuda.is_available() else 0


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _batches():
    rng = np.random.RandomState(72)
    return [make_batch(rng, N_USERS, N_ITEMS, B) for _ in range(STEPS)]


def _train_worker(rank, world, port, out_dir, exchange):
    from macr_b200 import ops
    from macr_b200.host import dist as mdist

    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    torch.cuda.set_device(rank)
    dev = torch.device("cuda", rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=dev)
    try:
        U, I, w, wu = make_model(71, N_USERS, N_ITEMS, scale=4.0)
        sh = mdist.RowShardedMFTrainer(U, I, w, wu, ops.HParams.make(**HP), B, rank=rank, world=world,
                                       device=dev, exchange=exchange)
        assert sh.exchange == exchange, (sh.exchange, exchange)
        losses = []
        bt = _batches()
        for u, p, n in bt[: STEPS // 2]:  # per-step device call
            ids = [torch.from_numpy(np.asarray(x, np.int32)).to(dev) for x in (u, p, n)]
            losses.append(sh.step_device(*ids).cpu().numpy().copy())
        host = torch.from_numpy(np.stack([np.stack(b) for b in bt[STEPS // 2:]]).astype(np.int32)).pin_memory()
        hl = sh.run_host(host)  # epoch call with host buffers
        losses += [x for x in hl.numpy()]
        sh.check_peers()
        out = {k: v.cpu().numpy() for k, v in sh.local_tables().items()}
        out["losses"] = np.stack(losses)
        out["bounds"] = np.array([sh.u_lo, sh.u_hi, sh.i_lo, sh.i_hi])
        np.savez(os.path.join(out_dir, f"{exchange}_rank{rank}.npz"), **out)
        sh.close()
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("exchange", ["allreduce", "push"])
def test_row_sharded_training_over_nccl_equals_single_gpu(tmp_path, exchange):
    if _n_gpus() < 2:
        pytest.skip("needs >= 2 GPUs")
    from macr_b200 import ops

    world = min(_n_gpus(), 4)
    _spawn(_train_worker, (world, _free_port(), str(tmp_path), exchange), world, str(tmp_path))
    U, I, w, wu = make_model(71, N_USERS, N_ITEMS, scale=4.0)
    tr = ops.MFTrainer(U, I, w, wu, ops.HParams.make(**HP), max_batch=B)
    want = []
    for u, p, n in _batches():
        ids = [torch.from_numpy(np.asarray(x, np.int32)).cuda() for x in (u, p, n)]
        want.append(tr.step_device(*ids).cpu().numpy().copy())
    t = tr.tab
    full = {k: getattr(t, k).cpu().numpy() for k in ("U", "mU", "vU", "I", "mI", "vI", "w", "wu")}
    for r in range(world):
        z = np.load(tmp_path / f"{exchange}_rank{r}.npz")
        u_lo, u_hi, i_lo, i_hi = (int(x) for x in z["bounds"])
        np.testing.assert_array_equal(z["losses"], np.stack(want))
        for k in ("U", "mU", "vU"):
            np.testing.assert_array_equal(z[k], full[k][u_lo:u_hi], err_msg=f"rank {r} {k}")
        for k in ("I", "mI", "vI"):
            np.testing.assert_array_equal(z[k], full[k][i_lo:i_hi], err_msg=f"rank {r} {k}")
        np.testing.assert_array_equal(z["w"], full["w"])
        np.testing.assert_array_equal(z["wu"], full["wu"])
    tr.close()


# ---- row-partitioned LightGCN --------------------------------------------------
LG_USERS, LG_ITEMS, LG_B, LG_L, LG_STEPS = 1201, 503, 256, 2, 4
LG_HP = dict(lr=1e-3, alpha=1e-2, beta=1e-3, decay=1e-4, batch_size=LG_B)


def _lg_inputs():
    from helpers import make_interactions, norm_adj_csr

    lists = make_interactions(3, LG_USERS, LG_ITEMS, 70)  # popular-enough rows: multi-segment rows exist
    rowptr, col, val = norm_adj_csr(lists, LG_USERS, LG_ITEMS)
    U, I, w, wu = make_model(9, LG_USERS, LG_ITEMS, scale=4.0)
    rng = np.random.RandomState(6)
    batches = np.stack([np.stack(make_batch(rng, LG_USERS, LG_ITEMS, LG_B)) for _ in range(LG_STEPS)]).astype(np.int32)
    return rowptr, col, val, U, I, w, wu, batches


def _lgcn_worker(rank, world, port, out_dir):
    from macr_b200 import ops
    from macr_b200.host import dist as mdist

    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    torch.cuda.set_device(rank)
    dev = torch.device("cuda", rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=dev)
    try:
        rowptr, col, val, U, I, w, wu, batches = _lg_inputs()
        sh = mdist.RowShardedLGCNTrainer(rowptr, col, val, U, I, w, wu, LG_L, ops.HParams.make(**LG_HP), LG_B,
                                         rank=rank, world=world, device=dev)
        assert sh.local_nnz < len(col)
        db = torch.from_numpy(batches).to(dev)
        l_eval = sh.run(db[:1], train=False).cpu().numpy()       # loss-only pass: moves nothing
        losses = sh.run(db[:2]).cpu().numpy()                    # epoch replay of the step graph
        losses = np.concatenate([losses, sh.run_host(torch.from_numpy(batches[2:]).pin_memory()).numpy()])
        ue, ie = sh.embeddings()
        sh.check_peers()
        out = {k: v.cpu().numpy() for k, v in sh.local_tables().items()}
        out.update(losses=losses, l_eval=l_eval, ue=ue.cpu().numpy(), ie=ie.cpu().numpy(),
                   bounds=np.array([sh.u_lo, sh.u_hi, sh.i_lo, sh.i_hi]))
        np.savez(os.path.join(out_dir, f"lgcn_rank{rank}.npz"), **out)
        sh.close()
    finally:
        dist.destroy_process_group()


def test_row_partitioned_lightgcn_equals_single_gpu(tmp_path):
    if _n_gpus() < 2:
        pytest.skip("needs >= 2 GPUs")
    from macr_b200 import ops

    world = min(_n_gpus(), 4)
    _spawn(_lgcn_worker, (world, _free_port(), str(tmp_path)), world, str(tmp_path))
    rowptr, col, val, U, I, w, wu, batches = _lg_inputs()
    tr = ops.LGCNTrainer(rowptr, col, val, U, I, w, wu, LG_L, ops.HParams.make(**LG_HP), max_batch=LG_B)
    db = torch.from_numpy(batches).cuda()
    l_eval = tr.run(db[:1], train=False).cpu().numpy()
    want = tr.run(db).cpu().numpy()
    ue, ie = tr.embeddings()
    t = tr.tab
    full = {k: getattr(t, k).cpu().numpy() for k in ("U", "mU", "vU", "I", "mI", "vI", "w", "wu")}
    assert np.abs(full["U"] - U).max() > 0
    for r in range(world):
        z = np.load(tmp_path / f"lgcn_rank{r}.npz")
        u_lo, u_hi, i_lo, i_hi = (int(x) for x in z["bounds"])
        np.testing.assert_array_equal(z["l_eval"], l_eval)
        np.testing.assert_array_equal(z["losses"], want)
        for k in ("U", "mU", "vU"):
            np.testing.assert_array_equal(z[k], full[k][u_lo:u_hi], err_msg=f"rank {r} {k}")
        for k in ("I", "mI", "vI"):
            np.testing.assert_array_equal(z[k], full[k][i_lo:i_hi], err_msg=f"rank {r} {k}")
        np.testing.assert_array_equal(z["w"], full["w"])
        np.testing.assert_array_equal(z["wu"], full["wu"])
        np.testing.assert_array_equal(z["ue"], ue.cpu().numpy())  # complete on every rank
        np.testing.assert_array_equal(z["ie"], ie.cpu().numpy())
    tr.close()


T_Q, S_ITEMS, K = 700, 9000, 20


def _score_inputs():
    rng = np.random.RandomState(5)
    U, I, w, wu = make_model(9, T_Q, S_ITEMS, scale=12.0)
    cnt = rng.randint(1, 40, T_Q)
    rp = np.zeros(T_Q + 1, np.int32)
    rp[1:] = np.cumsum(cnt)
    col = np.concatenate([np.sort(rng.choice(S_ITEMS, c, replace=False)) for c in cnt]).astype(np.int32)
    return U, I, w, wu, rp, col


def _score_worker(rank, world, port, out_dir):
    from macr_b200 import ops
    from macr_b200.host import dist as mdist

    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    torch.cuda.set_device(rank)
    dev = torch.device("cuda", rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=dev)
    try:
        U, I, w, wu, rp, col = _score_inputs()
        to = lambda a: torch.from_numpy(a).to(dev)
        dU, dI, dw, dwu, drp, dcol = to(U), to(I), to(w), to(wu), to(rp), to(col)
        su = ops.score_gates(dU, dwu)
        for name, make in (("items", lambda: mdist.ShardedScorer(dI, dw, rank=rank, world=world)),
                           ("items_nccl", lambda: mdist.ShardedScorer(dI, dw, rank=rank, world=world, exchange="nccl")),
                           ("users", lambda: mdist.UserShardedScorer(dI, dw, rank=rank, world=world))):
            scorer = make()
            for rep in range(3):  # the peer buffers are reused call after call
                ids, sc = scorer.topk(dU, su, 40.0, drp, dcol, K)
            if name == "items":
                assert scorer.exchange == "p2p", scorer.exchange  # fused exchange + merge over peer memory
                scorer.check_peers()
                scorer.close()
            np.savez(os.path.join(out_dir, f"score_{name}_rank{rank}.npz"), ids=ids.cpu().numpy(),
                     sc=sc.cpu().numpy())
    finally:
        dist.destroy_process_group()


def test_sharded_scoring_over_nccl_equals_unsharded(tmp_path):
    if _n_gpus() < 2:
        pytest.skip("needs >= 2 GPUs")
    from macr_b200 import ops

    world = min(_n_gpus(), 4)
    _spawn(_score_worker, (world, _free_port(), str(tmp_path)), world, str(tmp_path))
    U, I, w, wu, rp, col = _score_inputs()
    to = lambda a: torch.from_numpy(a).cuda()
    dU, dI = to(U), to(I)
    ids, sc = ops.score_topk(dU, dI, ops.score_gates(dI, to(w)), ops.score_gates(dU, to(wu)), 40.0, to(rp),
                             to(col), K)
    for name in ("items", "items_nccl", "users"):
        for r in range(world):
            z = np.load(tmp_path / f"score_{name}_rank{r}.npz")
            np.testing.assert_array_equal(z["ids"], ids.cpu().numpy(), err_msg=f"{name} rank {r}")
            np.testing.assert_array_equal(z["sc"], sc.cpu().numpy(), err_msg=f"{name} rank {r}")
