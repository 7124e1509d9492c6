#!/usr/bin/env python
"""Summarise .ncu-rep captures (read here, no GPU) into profiles/<tag>_<name>.txt:
the handful of counters the roofline argument rests on, per profiled launch."""
import csv
import io
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
WANT = [
    "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
    "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
    "dram__cycles_active.avg.pct_of_peak_sustained_elapsed",
    "lts__throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_sector_hit_rate.pct",
    "l1tex__throughput.avg.pct_of_peak_sustained_elapsed", "l1tex__t_sector_hit_rate.pct",
    "sm__throughput.avg.pct_of_peak_sustained_elapsed",
    "sm__warps_active.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_fmaheavy.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active",
    "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__issue_active.avg.pct_of_peak_sustained_active",
    "smsp__inst_executed.sum", "smsp__sass_thread_inst_executed_op_ffma_pred_on.sum",
    "smsp__sass_thread_inst_executed_op_fmul_pred_on.sum",
    "smsp__sass_thread_inst_executed_op_fadd_pred_on.sum",
    "launch__registers_per_thread", "launch__grid_size", "launch__block_size",
    "launch__waves_per_multiprocessor", "launch__occupancy_limit_shared_mem",
    "launch__occupancy_limit_registers", "launch__occupancy_limit_warps",
    "sm__maximum_warps_per_active_cycle_pct", "smsp__cycles_active.avg",
    "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_dispatch_stall_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio",
]


def summarise(rep, out, note=""):
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True,
                         text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr, units = rows[0], rows[1]
    lines = [f"# {os.path.basename(rep)}  (ncu --set full --clock-control none, B200){note}"]
    traffic = []
    for r in rows[2:]:
        lines.append(f"\nkernel: {r[hdr.index('Kernel Name')]}")
        rd = wr = None
        for w in WANT:
            if w in hdr:
                i = hdr.index(w)
                lines.append(f"  {w:70s} {r[i]:>16s} {units[i]}")
                if w == "dram__bytes_read.sum":
                    rd = (float(r[i]), units[i])
                if w == "dram__bytes_write.sum":
                    wr = (float(r[i]), units[i])
        if rd and wr:
            scale = {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1.0}
            t = rd[0] * scale[rd[1]] + wr[0] * scale[wr[1]]
            traffic.append(t)
            lines.append(f"  {'=> DRAM traffic (read+write)':70s} {t / 1e9:16.4f} GB")
    open(out, "w").write("\n".join(lines) + "\n")
    return sum(traffic) / len(traffic) if traffic else None


if __name__ == "__main__":
    tag = sys.argv[1]
    traffic_path = os.path.join(ROOT, "profiles", "ncu_traffic.json")
    tr = json.load(open(traffic_path)) if os.path.exists(traffic_path) else {}
    for spec in sys.argv[2:]:
        name, rep, workload = spec.split(":")
        t = summarise(rep, os.path.join(ROOT, "profiles", f"{tag}_{name}.txt"))
        if workload and t:
            tr[workload] = t
        print(name, t)
    json.dump(tr, open(traffic_path, "w"), indent=1)
