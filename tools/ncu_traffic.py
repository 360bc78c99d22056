#!/usr/bin/env python
"""Per-kernel DRAM traffic of this round's kernels from their `ncu --set full` captures (read here, no GPU needed):

    python tools/ncu_traffic.py <tag>      # reads gpurun_out/<tag>_{k_count,k_locate,k_extract}.ncu-rep

writes profiles/kernel_traffic.json (what bench.py reports as roofline.traffic) and profiles/<tag>_<kernel>_ncu_summary.txt.
Traffic = dram__bytes_read.sum + dram__bytes_write.sum of ONE launch; `requested` = 32-byte sectors the kernel's global loads
asked L1 for (l1tex__t_sectors_pipe_lsu_mem_global_op_ld) — dram / requested > 1 means DRAM moved bytes nobody asked for."""
import csv
import io
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SCALE = {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1.0, "Tbyte": 1e12}
TIME = {"ns": 1e-9, "us": 1e-6, "ms": 1e-3, "s": 1.0, "nsecond": 1e-9, "usecond": 1e-6, "msecond": 1e-3, "second": 1.0}


def read_rep(rep):
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], stdout=subprocess.PIPE, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr, units = rows[0], rows[1]
    r = rows[2]

    def val(name, scale=None):
        i = hdr.index(name)
        v = float(r[i].replace(",", ""))
        if scale:
            v *= scale.get(units[i], 1.0)
        return v
    out = {"kernel_name": r[hdr.index("Kernel Name")][:120],
           "dram_bytes_per_launch": val("dram__bytes_read.sum", SCALE) + val("dram__bytes_write.sum", SCALE),
           "dram_bytes_read": val("dram__bytes_read.sum", SCALE),
           "capture_kernel_ms": val("gpu__time_duration.sum", TIME) * 1e3,
           "l2_hit_pct": val("lts__t_sector_hit_rate.pct"),
           "requested_bytes": 32.0 * val("l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum"),
           "lanes_per_instruction": val("smsp__thread_inst_executed_per_inst_executed.ratio"),
           "warp_instructions": val("smsp__inst_executed.sum"),
           "registers": val("launch__registers_per_thread"),
           "issue_active_pct": val("smsp__issue_active.avg.pct_of_peak_sustained_active"),
           "warps_active_pct": val("sm__warps_active.avg.pct_of_peak_sustained_active")}
    out["dram_over_requested"] = out["dram_bytes_per_launch"] / max(out["requested_bytes"], 1.0)
    out["capture_dram_gb_per_s"] = out["dram_bytes_per_launch"] / (out["capture_kernel_ms"] / 1e3) / 1e9
    return out


def main():
    tag = sys.argv[1]
    commit = subprocess.run(["git", "-C", ROOT, "rev-parse", "--short", "HEAD"], stdout=subprocess.PIPE, text=True).stdout.strip()
    res = {"captured_at_commit": commit, "capture": "ncu --set full --clock-control none --import-source on, one launch per kernel (gpurun_out/%s_*.ncu-rep)" % tag,
           "workload": "bench.py default (BASELINE.json configs[1]): 1 M patterns len 4-64 over the 1 GiB-text index; locate max 1000 hits; extractUntilBoundary of 1 M located hits",
           "note": "cold-cache, serialised launches under the profiler: the kernel's SHARE of traffic matters, its time is in capture_kernel_ms", "kernels": {}}
    for k in ("k_count", "k_locate", "k_extract"):
        rep = os.path.join(ROOT, "gpurun_out", "%s_%s.ncu-rep" % (tag, k))
        if not os.path.exists(rep):
            continue
        res["kernels"][k] = read_rep(rep)
        summ = subprocess.run([sys.executable, os.path.join(ROOT, "tools", "ncu_summary.py"), rep], stdout=subprocess.PIPE, text=True).stdout
        with open(os.path.join(ROOT, "profiles", "%s_%s_ncu_summary.txt" % (tag, k)), "w") as fh:
            fh.write(summ)
    with open(os.path.join(ROOT, "profiles", "kernel_traffic.json"), "w") as fh:
        json.dump(res, fh, indent=1)
    print(json.dumps(res, indent=1))


if __name__ == "__main__":
    main()
