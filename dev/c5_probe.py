"""Config-5 step probe: where do the 3.55 ms go?  (developer script, one GPU)"""
import os, sys, json, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import bench
from macr_b200 import ops

dev = torch.device("cuda", 0)
torch.cuda.set_device(dev)
w, wu = bench.synth_model(12345, 8, 8)[2:]
hp = ops.HParams.make(**bench.C5_HP)
B = bench.C5_BATCH
DIV = int(os.environ.get("PROBE_DIV", "1"))  # tables of 1/DIV the size (what one of DIV ranks holds)
NU, NI = bench.C5_USERS // DIV, bench.C5_ITEMS // DIV
U = bench.DeviceRows(NU, 11, dev)[0:NU]
I = bench.DeviceRows(NI, 13, dev)[0:NI]
tr = ops.MFTrainer(U, I, w, wu, hp, max_batch=B, device=dev)
t = tr.tab
mode = os.environ.get("PROBE_STATE", "fill")
if mode == "fill":
    for x in (t.mU, t.mI): x.fill_(1e-9)
    for x in (t.vU, t.vI): x.fill_(1e-12)
else:
    g = torch.Generator(device=dev).manual_seed(1)
    for m_, v_ in ((t.mU, t.vU), (t.mI, t.vI)):
        m_.normal_(0.0, 1e-4, generator=g); v_.uniform_(1e-9, 1e-7, generator=g)
nb = 24
bz = bench.synth_batches(12345, nb, NU, NI, B)
bu = bz.copy(); bu[:, 1] = np.random.RandomState(3).randint(0, NI, (nb, B))
out = {"div": DIV, "sweep_ctas": os.environ.get("MACR_SWEEP_CTAS_PER_SM", "auto")}
for name, b in (("zipf", bz), ("uniform", bu)):
    d_b = torch.from_numpy(b).to(dev)
    losses = torch.zeros((nb, 4), device=dev)
    tr.run(d_b[:4], losses[:4]); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); tr.run(d_b[4:], losses[4:]); e1.record(); torch.cuda.synchronize()
    out[name + "_epoch_ms"] = e0.elapsed_time(e1) / (nb - 4)
    e0.record()
    for s in range(4, nb): tr.run(d_b[s:s+1], losses[s:s+1])
    e1.record(); torch.cuda.synchronize()
    out[name + "_per_call_ms"] = e0.elapsed_time(e1) / (nb - 4)
e0.record()
for _ in range(5):
    ops.adam_sweep_untouched(t.U, t.mU, t.vU, None, 1e-6)
    ops.adam_sweep_untouched(t.I, t.mI, t.vI, None, 1e-6)
e1.record(); torch.cuda.synchronize()
out["sweep_alone_ms"] = e0.elapsed_time(e1) / 5
print(json.dumps(out))
