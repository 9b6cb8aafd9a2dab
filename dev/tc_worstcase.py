"""worst case for the candidate lists: every user's train items are exactly its top-scoring items
(what a well-trained model does), ml_10m shapes and train-list lengths"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from macr_b200 import ops

dev = torch.device("cuda")
rng = np.random.RandomState(0)
T_users, n_items, K, c = 13878, 8790, 20, 40.0
lim = lambda r: np.sqrt(6.0 / (r + 64))
U = (rng.uniform(-lim(T_users), lim(T_users), (T_users, 64)) * 30).astype(np.float32)
I = (rng.uniform(-lim(n_items), lim(n_items), (n_items, 64)) * 30).astype(np.float32)
w = rng.uniform(-0.3, 0.3, 64).astype(np.float32)
wu = rng.uniform(-0.3, 0.3, 64).astype(np.float32)
dU, dI = torch.from_numpy(U).to(dev), torch.from_numpy(I).to(dev)
si, su = ops.score_gates(dI, torch.from_numpy(w).to(dev)), ops.score_gates(dU, torch.from_numpy(wu).to(dev))
lens = np.minimum(2000, np.maximum(5, rng.lognormal(np.log(109), 0.9, T_users).astype(np.int64)))  # median 109
S = ops.score_matrix(dU, dI, si, su, c)                     # dense scores, to pick the top items as train items
order = torch.argsort(S, dim=1, descending=True)[:, :2000].cpu().numpy()
del S
rowptr = np.zeros(T_users + 1, np.int32); rowptr[1:] = np.cumsum(lens)
col = np.concatenate([np.sort(order[t, :lens[t]]) for t in range(T_users)]).astype(np.int32)
mrp, mcol = torch.from_numpy(rowptr).to(dev), torch.from_numpy(col).to(dev)
stats = torch.zeros(2, dtype=torch.int64, device=dev)
ti, ts = ops.score_topk_tc(dU, dI, si, su, c, mrp, mcol, K, stats=stats)
ei, es = ops.score_topk_exact(dU, dI, si, su, c, mrp, mcol, K)
print("exact match:", bool((ti == ei).all().item() and (ts == es).all().item()), " rows redone by the exact kernel:",
      int(stats[0].item()), "of", T_users, " train-list median", int(np.median(lens)), "mean %.0f" % lens.mean())
for fn, name in ((lambda: ops.score_topk_tc(dU, dI, si, su, c, mrp, mcol, K), "tcgen05 path (with fallback rows)"),
                 (lambda: ops.score_topk_exact(dU, dI, si, su, c, mrp, mcol, K), "exact fp32 kernel")):
    for _ in range(2):
        fn()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(5):
        fn()
    b.record(); torch.cuda.synchronize()
    print("  %-36s %.3f ms" % (name, a.elapsed_time(b) / 5))
