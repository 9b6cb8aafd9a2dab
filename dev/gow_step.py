"""Quick probe: gowalla-shape MF step (ring of 3 models / back to back), no other bench blocks.
  MACR_GRAPH_UNROLL=8 python dev/gow_step.py      (MACR_B200_LIB=/path/to/other/libmacr_b200.so for an A/B)"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import bench
cx = bench.Ctx()
torch = cx.torch
from macr_b200 import ops
K = int(os.environ.get("K", "240"))
hp = ops.HParams.make(**bench.HP)
nb = min(bench.N_BATCHES, K)
bh = bench.synth_batches(12345, nb)
batches = torch.from_numpy(bh).to(cx.dev)
losses = torch.zeros((nb, 4), dtype=torch.float32, device=cx.dev)
ring = [ops.MFTrainer(*bench.synth_model(12345 + 31 * k), hp, max_batch=bench.BATCH, device=cx.dev) for k in range(3)]
tr = ring[0]
def ring_time(chunk):
    for s in range(9):
        ring[s % 3].run(batches[:chunk], losses[:chunk])
    r0, r1 = cx.events()
    torch.cuda.synchronize(); r0.record()
    n = 0
    while n < K:
        ring[(n // chunk) % 3].run(batches[:chunk], losses[:chunk]); n += chunk
    r1.record(); torch.cuda.synchronize()
    return r0.elapsed_time(r1) * 1e3 / n
def b2b():
    tr.run(batches[:nb], losses[:nb])
    e0, e1 = cx.events()
    torch.cuda.synchronize(); e0.record()
    tr.run(batches[:nb], losses[:nb])
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) * 1e3 / nb
print("unroll=%s ring(1 step/call) %.2f us  ring(8 steps/call) %.2f us  back-to-back %.2f us  final loss %.7f" % (
    os.environ.get("MACR_GRAPH_UNROLL", "1"), ring_time(1), ring_time(8), b2b(), float(losses[nb - 1, 0].item())))
