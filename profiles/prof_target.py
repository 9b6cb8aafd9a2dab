#!/usr/bin/env python
"""Short ncu target: a few launches of every hot kernel at the bench shapes (gowalla, B=4096).

  ncu --set full --clock-control none --import-source on -k regex:macr -c 60 \
      -o gpurun_out/prof_rN python profiles/prof_target.py

Nothing printed by a run under ncu is a bench value."""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
from macr_b200 import ops  # noqa: E402


def main():
    dev = torch.device("cuda", 0)
    steps = int(os.environ.get("PROF_STEPS", "3"))
    U, I, w, wu = bench.synth_model(12345)
    hp = ops.HParams.make(**bench.HP)
    tr = ops.MFTrainer(U, I, w, wu, hp, max_batch=bench.BATCH, device=dev)
    batches = torch.from_numpy(bench.synth_batches(12345, steps)).to(dev)
    losses = torch.zeros((steps, 4), dtype=torch.float32, device=dev)
    for s in range(steps):
        tr.run(batches[s:s + 1], losses[s:s + 1])
    torch.cuda.synchronize()
    # steady-state sweep (every row has non-zero moments)
    rows = bench.N_USERS + bench.N_ITEMS
    a = torch.randn((rows, 64), device=dev) * 0.05
    b = torch.randn((rows, 64), device=dev) * 1e-4
    c = torch.rand((rows, 64), device=dev) * 1e-7 + 1e-9
    for _ in range(2):
        ops.adam_sweep_untouched(a, b, c, None, 1e-4)
    torch.cuda.synchronize()
    # scoring: 15424 x 40981, top-20, masked
    if os.environ.get("PROF_SCORING", "1") == "1":
        Us, Is, ws_, wus = bench.synth_model(777)
        dU, dI = torch.from_numpy(Us * 10).to(dev), torch.from_numpy(Is * 10).to(dev)
        q = torch.from_numpy(np.random.RandomState(5).permutation(bench.N_USERS)[:bench.N_TEST_USERS]
                             .astype(np.int32)).to(dev)
        mrp, mcol = bench.synth_mask(9, bench.N_TEST_USERS, 27)
        Uq = ops.gather_rows(dU, q)
        si = ops.score_gates(dI, torch.from_numpy(ws_).to(dev))
        su = ops.score_gates(Uq, torch.from_numpy(wus).to(dev))
        ops.score_topk(Uq, dI, si, su, 40.0, torch.from_numpy(mrp).to(dev),
                       torch.from_numpy(mcol).to(dev), bench.TOPK)
        torch.cuda.synchronize()
    tr.close()
    # LightGCN: 2 steps on the real yelp2018 adjacency when it is staged (data/ is git-ignored)
    path = os.path.join(ROOT, "data", "yelp2018")
    if os.environ.get("PROF_LGCN", "1") == "1" and os.path.exists(os.path.join(path, "train.txt")):
        import contextlib
        import io

        from macr_b200.host.data_lgcn import Data

        with contextlib.redirect_stdout(io.StringIO()):
            data = Data(path, 4096)
            rowptr, col, val = data.adj_csr("pre")
        rng = np.random.RandomState(2)
        lim = lambda r: np.sqrt(6.0 / (r + 64))
        Ue = rng.uniform(-lim(data.n_users), lim(data.n_users), (data.n_users, 64)).astype(np.float32)
        Ie = rng.uniform(-lim(data.n_items), lim(data.n_items), (data.n_items, 64)).astype(np.float32)
        wv = rng.uniform(-0.3, 0.3, 64).astype(np.float32)
        lt = ops.LGCNTrainer(rowptr, col, val, Ue, Ie, wv, wv.copy(), 2,
                             ops.HParams.make(lr=1e-3, alpha=1e-2, beta=1e-3, decay=1e-4, batch_size=4096),
                             max_batch=4096, device=dev)
        b = np.stack([np.stack([rng.permutation(data.n_users)[:4096], rng.randint(0, data.n_items, 4096),
                                rng.randint(0, data.n_items, 4096)]) for _ in range(2)]).astype(np.int32)
        lt.run(torch.from_numpy(b).to(dev), True)
        torch.cuda.synchronize()
        lt.close()


if __name__ == "__main__":
    main()
