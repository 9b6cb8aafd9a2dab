import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import numpy as np, torch
from macr_b200 import ops
from macr_b200._lib import lib
from helpers import make_model, make_interactions, lists_to_csr
T_users, n_items, K, c = 15424, 40981, 20, 40.0
splits = tuple(int(x) for x in (sys.argv[1:3] if len(sys.argv) > 2 else (1, 3)))
U, I, w, wu = make_model(0, T_users, n_items, scale=10.0)
rng = np.random.RandomState(1)
cnt = np.maximum(1, rng.poisson(27, T_users))
rowptr = np.zeros(T_users + 1, np.int32); rowptr[1:] = np.cumsum(cnt)
col = np.concatenate([np.sort(rng.choice(n_items, size=k, replace=False)) for k in cnt]).astype(np.int32)
dev = torch.device("cuda")
dU, dI = torch.from_numpy(U).to(dev), torch.from_numpy(I).to(dev)
si, su = ops.score_gates(dI, torch.from_numpy(w).to(dev)), ops.score_gates(dU, torch.from_numpy(wu).to(dev))
mrp, mcol = torch.from_numpy(rowptr).to(dev), torch.from_numpy(col).to(dev)
import ctypes
dbg = int(os.environ.get("TC_DBG", "0"))
h = ctypes.CDLL(lib()._name)
prof = torch.zeros(64, dtype=torch.int64, device=dev)
h.macr_score_tc_debug.argtypes = [ctypes.c_int, ctypes.c_void_p]
h.macr_score_tc_debug(dbg, ctypes.c_void_p(prof.data_ptr()) if os.environ.get("TC_PROF") else None)
for _ in range(3):
    ops.score_topk_tc(dU, dI, si, su, c, mrp, mcol, K)
torch.cuda.synchronize()
if os.environ.get("TC_PROF"):
    pr = prof.cpu().tolist()
    for mode, name in ((0, "MAX"), (1, "FILTER")):
        b = 32 * mode
        print(f"{name}: producer  wait A_EMPTY {pr[b+0]}  wait B_EMPTY {pr[b+1]}  total {pr[b+2]}")
        print(f"{name}: mma       wait A_FULL {pr[b+4]}  wait B_FULL {pr[b+5]}  wait TM_EMPTY wg0 {pr[b+6]} wg1 {pr[b+7]}  total {pr[b+8]}  tiles {pr[b+9]}")
        for g in (0, 1):
            print(f"{name}: epilogue wg{g}  bar.sync {pr[b+12+4*g]}  wait TM_FULL {pr[b+13+4*g]}  total {pr[b+14+4*g]}")
