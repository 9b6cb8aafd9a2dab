"""ncu target: the gowalla-shape scoring call (15 424 x 40 981, masked top-20), a few repetitions."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
cx = bench.Ctx()
r = bench.gowalla_scoring(cx, reps=int(os.environ.get("REPS", "2")))
print(r["ms_per_eval"], r["candidates_per_row"], r["checksum"])
