"""Row-sharded MF training at config-5 table size (10 M users x 1 M items, d=64), one rank per GPU:

    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 \
        --master-port 29511 dev/bench_sharded_train.py [--users 10000000 --items 1000000 --batch 8192]

Each rank holds 1/N of both tables and of their Adam slots (RowShardedMFTrainer); a step is one
all-reduce of the 3B gathered rows + the single-GPU step graph on the local slice.  Prints one
JSON line per run (rank 0): interactions/s, ms/step, HBM bytes per rank and step."""
import argparse
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import torch.distributed as dist

from macr_b200 import ops
from macr_b200.host.dist import RowShardedMFTrainer, item_shard_bounds, user_shard_bounds


class _LazyRows:
    """[rows, 64] Xavier-scale table generated slice by slice (only the owned slice ever exists)."""

    def __init__(self, rows, seed):
        self.shape, self.seed = (rows, 64), seed

    def __getitem__(self, sl):
        lo, hi = sl.start or 0, sl.stop
        lim = np.float32(np.sqrt(6.0 / (self.shape[0] + 64)))
        x = np.random.default_rng(self.seed + lo).random((hi - lo, 64), dtype=np.float32)
        x *= 2 * lim
        x -= lim
        return x


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--users", type=int, default=10_000_000)
    ap.add_argument("--items", type=int, default=1_000_000)
    ap.add_argument("--batch", type=int, default=8192)
    ap.add_argument("--steps", type=int, default=30)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--cold", action="store_true",
                    help="keep m = v = 0: the dense sweep then skips the rows never touched (reads m, v only)")
    a = ap.parse_args()
    rank, world = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1))
    torch.cuda.set_device(int(os.environ.get("LOCAL_RANK", 0)))
    dev = torch.device("cuda", torch.cuda.current_device())
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    rng = np.random.RandomState(3)
    w, wu = (rng.uniform(-0.3, 0.3, 64).astype(np.float32) for _ in range(2))
    hp = ops.HParams.make(lr=1e-3, alpha=1e-3, beta=1e-3, decay=1e-5, batch_size=a.batch)
    tr = RowShardedMFTrainer(_LazyRows(a.users, 11), _LazyRows(a.items, 13), w, wu, hp, a.batch, rank=rank,
                             world=world, device=dev)
    if not a.cold:  # steady state: every row carries a momentum, the sweep moves 24 B per element
        t = tr.trainer.tab
        for x in (t.mU, t.vU, t.mI, t.vI):
            x.fill_(1e-9)
    g = torch.Generator(device=dev).manual_seed(5)  # same seed on every rank: identical batches
    mk = lambda hi: torch.randint(0, hi, (a.steps + a.warmup, a.batch), generator=g, device=dev, dtype=torch.int32)
    U_ids, P_ids, N_ids = mk(a.users), mk(a.items), mk(a.items)
    for s in range(a.warmup):
        tr.step_device(U_ids[s], P_ids[s], N_ids[s])
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for s in range(a.warmup, a.warmup + a.steps):
        loss = tr.step_device(U_ids[s], P_ids[s], N_ids[s])
    e1.record()
    torch.cuda.synchronize()
    ms = torch.tensor([e0.elapsed_time(e1) / a.steps], device=dev)
    if world > 1:
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    plain_ms = None
    if world == 1:  # the same tables without the exchange glue: ids are all owned, local id == global id
        for s in range(a.warmup):
            tr.trainer.step_device(U_ids[s], P_ids[s], N_ids[s])
        e0.record()
        for s in range(a.warmup, a.warmup + a.steps):
            tr.trainer.step_device(U_ids[s], P_ids[s], N_ids[s])
        e1.record()
        torch.cuda.synchronize()
        plain_ms = e0.elapsed_time(e1) / a.steps
    if rank == 0:
        ub, ib = user_shard_bounds(a.users, world), item_shard_bounds(a.items, world)
        rows = int(ub[1] - ub[0] + ib[1] - ib[0]) + 3 * a.batch
        print(json.dumps({"config": f"row-sharded MF step U={a.users} I={a.items} B={a.batch}", "n_gpus": world,
                          "ms_per_step": float(ms.item()), "interactions_per_sec": a.batch / (float(ms.item()) * 1e-3),
                          "hbm_bytes_per_rank_step": 24 * 64 * rows, "exchange_bytes_per_step": 3 * a.batch * 64 * 4,
                          "scaling": "strong", "adam_state": "cold" if a.cold else "warm", "loss": float(loss[0].item()),
                          "ms_per_step_without_exchange_glue": plain_ms}))
    tr.close()
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
