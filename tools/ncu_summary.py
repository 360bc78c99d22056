#!/usr/bin/env python
"""Summarise an .ncu-rep (read here, no GPU needed): python tools/ncu_summary.py <rep> [kernel-substring]"""
import csv
import io
import subprocess
import sys

WANT = [
    "gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread", "launch__occupancy_limit_registers",
    "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__issue_active.avg.pct_of_peak_sustained_active",
    "smsp__warps_eligible.avg.per_cycle_active", "smsp__inst_executed.sum", "smsp__thread_inst_executed_per_inst_executed.ratio",
    "sm__throughput.avg.pct_of_peak_sustained_elapsed", "dram__bytes_read.sum", "dram__bytes_write.sum",
    "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram__throughput.avg.pct_of_peak_sustained_elapsed",
    "lts__t_sector_hit_rate.pct", "lts__throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_sectors_srcunit_tex_op_read.sum",
    "l1tex__t_sector_hit_rate.pct", "l1tex__t_requests_pipe_lsu_mem_global_op_ld.sum", "l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum",
    "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio", "smsp__average_warp_latency_issue_stalled_long_scoreboard.ratio",
    "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio", "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio", "smsp__average_warps_issue_stalled_branch_resolving_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio", "smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio", "smsp__average_warps_issue_stalled_no_instruction_per_issue_active.ratio",
]


def main():
    rep = sys.argv[1]
    sub = sys.argv[2] if len(sys.argv) > 2 else ""
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], stdout=subprocess.PIPE, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr, units = rows[0], rows[1]
    kcol = hdr.index("Kernel Name")
    for r in rows[2:]:
        if sub and sub not in r[kcol]:
            continue
        print("kernel:", r[kcol][:100])
        for w in WANT:
            if w in hdr:
                i = hdr.index(w)
                print("  %-82s %s %s" % (w, r[i], units[i]))
        rd = float(r[hdr.index("dram__bytes_read.sum")])
        wr = float(r[hdr.index("dram__bytes_write.sum")])
        scale = {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1.0}
        tot = rd * scale[units[hdr.index("dram__bytes_read.sum")]] + wr * scale[units[hdr.index("dram__bytes_write.sum")]]
        print("  dram_bytes_per_launch %.0f" % tot)
        # derived: achieved HBM GB/s, 32-byte sectors per global-load request, DRAM sectors per second
        tu = {"ns": 1e-9, "us": 1e-6, "ms": 1e-3, "s": 1.0, "nsecond": 1e-9, "usecond": 1e-6, "msecond": 1e-3, "second": 1.0}
        ti = hdr.index("gpu__time_duration.sum")
        secs = float(r[ti]) * tu.get(units[ti], 1e-9)
        print("  derived: achieved_hbm_gb_per_s %.1f   dram_sectors_per_s %.1f G" % (tot / secs / 1e9, tot / 32.0 / secs / 1e9))
        try:
            req = float(r[hdr.index("l1tex__t_requests_pipe_lsu_mem_global_op_ld.sum")])
            sec = float(r[hdr.index("l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum")])
            print("  derived: sectors_per_global_load_request %.2f (32 lanes x 1 record = 32 when all lanes load distinct records)" % (sec / req))
        except (ValueError, ZeroDivisionError):
            pass


if __name__ == "__main__":
    main()
