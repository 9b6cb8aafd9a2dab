"""developer probe: where the time of UserShardedScorer.topk goes at world > 1"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch, torch.distributed as dist
import bench
from macr_b200 import ops
from macr_b200.host.dist import UserShardedScorer, all_gather_rows

rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
torch.cuda.set_device(int(os.environ["LOCAL_RANK"]))
dev = torch.device("cuda", int(os.environ["LOCAL_RANK"]))
dist.init_process_group("nccl", device_id=dev)
Us, Is, ws_, wus = bench.synth_model(777)
dU, dI = torch.from_numpy(Us * 10).to(dev), torch.from_numpy(Is * 10).to(dev)
T_q = bench.N_TEST_USERS
q = torch.from_numpy(np.random.RandomState(5).permutation(bench.N_USERS)[:T_q].astype(np.int32)).to(dev)
mrp, mcol = bench.synth_mask(9, T_q, 27)
dmrp, dmcol = torch.from_numpy(mrp).to(dev), torch.from_numpy(mcol).to(dev)
sc = UserShardedScorer(dI, torch.from_numpy(ws_).to(dev), rank=rank, world=world)
Uq = ops.gather_rows(dU, q)
su = ops.score_gates(Uq, torch.from_numpy(wus).to(dev))


def timeit(fn, n=20):
    for _ in range(3):
        fn()
    dist.barrier(); torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(n):
        fn()
    b.record(); torch.cuda.synchronize()
    return a.elapsed_time(b) / n

t_local = timeit(lambda: sc.topk_local(Uq, su, 40.0, dmrp, dmcol, 20))
ids, s = sc.topk_local(Uq, su, 40.0, dmrp, dmcol, 20)
both = torch.cat([ids, s.view(torch.int32)], dim=1)
t_gather = timeit(lambda: all_gather_rows(both, T_q, world, rank))
t_full = timeit(lambda: sc.topk(Uq, su, 40.0, dmrp, dmcol, 20))
if rank == 0:
    print(f"topk_local {t_local:.3f} ms  all_gather_rows {t_gather:.3f} ms  topk {t_full:.3f} ms")
dist.destroy_process_group()
