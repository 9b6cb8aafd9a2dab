"""The CPU-runnable parts of bench.py's contract: the `--impl reference` arm prints ONE JSON line
with the keys the driver reads, and the CPU scoring baseline (the reference's own C++ evaluator
when oracle/_ref is built) returns a positive throughput on a small sample."""
import importlib.util
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_reference_arm_prints_one_contract_line():
    # the full 10M x 1M tables need ~11 GB of host memory and ~2 s per step: the contract is checked
    # on a 1 % scale model of the same workload (MACR_BENCH_SCALE is a dry-run knob of bench.py)
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "2",
                          "--warmup", "1"], capture_output=True, text=True, timeout=600, cwd=ROOT,
                         env=dict(os.environ, MACR_BENCH_SCALE="0.01"))
    assert out.returncode == 0, out.stderr[-2000:]
    lines = [l for l in out.stdout.splitlines() if l.startswith("{")]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["metric"] == "train_interactions_per_sec"
    assert d["unit"] == "interactions/s" and d["higher_is_better"] is True and d["value"] > 0
    assert d["config"]["workload"] and d["steps"] == 2 and d["warmup"] == 1 and d["scaling"] == "strong"
    cb = d["cpu_baseline"]
    assert cb["kind"] in ("port", "reference") and cb["cores"] >= 1 and cb["sample"] and cb["value"] == d["value"]
    e = d["e2e"]
    assert e["value"] == d["value"] and e["unit"] == d["unit"]
    assert e["h2d_bytes_per_step"] == 0 and e["d2h_bytes_per_step"] == 0


def test_cpu_scoring_baseline_runs_on_a_small_sample():
    spec = importlib.util.spec_from_file_location("bench_under_test", os.path.join(ROOT, "bench.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)  # importing must not start a benchmark
    r = mod.cpu_scoring_baseline(sample_users=64, reps=1)
    assert r["unit"] == "scores/s" and r["value"] > 0 and r["kind"] in ("reference", "port") and r["cores"] >= 1
    json.dumps(r)
