"""ncu / timing target: the stateless B x B grid call at B = 4096 and 8192."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from macr_b200 import ops
dev = torch.device("cuda", 0)
for B in (4096, 8192):
    g = [torch.randn(B, device=dev) * sd for sd in (0.05, 0.05, 0.1, 0.1, 0.1)]
    nbytes = ops.lib().macr_grid_bce_workspace_bytes(B)
    bufs = (torch.empty(nbytes, dtype=torch.uint8, device=dev), torch.empty(3, device=dev), torch.empty((5, B), device=dev))
    for _ in range(3):
        ops.grid_bce(*g, 1e-2, 1e-3, bufs=bufs)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize(); e0.record()
    for _ in range(50):
        ops.grid_bce(*g, 1e-2, 1e-3, bufs=bufs)
    e1.record(); torch.cuda.synchronize()
    print(B, "us per call (memset + gates + grid):", e0.elapsed_time(e1) / 50 * 1e3)
