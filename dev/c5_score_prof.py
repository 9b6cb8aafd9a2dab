"""ncu target: the config-5 scoring call (262 144 users x 1 M items, masked top-20), one GPU."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
cx = bench.Ctx()
r = bench.sharded_scoring(cx, reps=1)
print(r["ms_per_eval"], r["candidates_per_row_rank0"], r["checksum"])
