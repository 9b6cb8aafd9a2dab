"""torchrun --nproc-per-node N dev/shard_probe.py : where does the sharded step's time go?"""
import os, sys, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch, torch.distributed as dist
import bench
from macr_b200 import ops
from macr_b200.host.dist import RowShardedMFTrainer

rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
torch.cuda.set_device(int(os.environ["LOCAL_RANK"]))
dev = torch.device("cuda", torch.cuda.current_device())
dist.init_process_group("nccl", device_id=dev)
w, wu = bench.synth_model(12345, 8, 8)[2:]
hp = ops.HParams.make(**bench.C5_HP)
B, nb, K = bench.C5_BATCH, 24, 20
bh = bench.synth_batches(12345, nb, bench.C5_USERS, bench.C5_ITEMS, B)
ids = torch.from_numpy(bh).to(dev)
losses = torch.zeros((nb, 4), device=dev)
out = {}
for mode in ("push", "allreduce"):
    sh = RowShardedMFTrainer(bench.DeviceRows(bench.C5_USERS, 11, dev), bench.DeviceRows(bench.C5_ITEMS, 13, dev),
                             w, wu, hp, B, rank=rank, world=world, device=dev, exchange=mode)
    t = sh.trainer.tab
    for x in (t.mU[: sh.n_lu], t.mI[: sh.n_li]): x.fill_(1e-9)
    for x in (t.vU[: sh.n_lu], t.vI[: sh.n_li]): x.fill_(1e-12)
    def timed(fn, n=K):
        for s in range(3): fn(s)
        torch.cuda.synchronize(); dist.barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for s in range(n): fn(3 + s)
        e1.record(); torch.cuda.synchronize()
        return e0.elapsed_time(e1) / n
    out[mode + "_step_ms"] = timed(lambda s: sh.step_ids3(ids[s % nb].view(-1), B, losses[s % nb:s % nb + 1]))
    out[mode + "_exchange_ms"] = timed(lambda s: sh.exchange_rows(ids[s % nb].view(-1), B))
    # graph only: local ids of one batch, no exchange, ranks free-running
    sh.exchange_rows(ids[0].view(-1), B)
    l3 = sh._local3[: 3 * B].clone().view(1, 3, B)
    out[mode + "_graph_only_ms"] = timed(lambda s: sh.trainer.run(l3, losses[:1]))
    # exchange with a host sync between steps (no overlap at all): pure latency view
    sh.check_peers()
    sh.close(); del sh; torch.cuda.empty_cache()
res = [None] * world
dist.all_gather_object(res, out)
if rank == 0:
    print(json.dumps(res))
dist.destroy_process_group()
