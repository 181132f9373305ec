#!/usr/bin/env python
"""Summarise ncu output brought back from the GPU box into text files under profiles/.

  python tools/ncu_summary.py launches gpurun_out/<tag>/launches.csv            -> per-kernel launch shares
  python tools/ncu_summary.py full gpurun_out/<tag>/prof_x.ncu-rep              -> key metrics of each captured launch
(reads .ncu-rep with the local `ncu -i`, which needs no GPU)."""
import collections
import csv
import io
import subprocess
import sys

KEYS = [
    "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
    "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram__cycles_active.avg.pct_of_peak_sustained_elapsed",
    "lts__t_bytes.sum", "l1tex__t_bytes.sum",
    "sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm__cycles_elapsed.avg", "sm__cycles_active.avg",
    "smsp__inst_executed.sum", "smsp__issue_active.avg.pct_of_peak_sustained_active",
    "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active", "sm__pipe_fmaheavy_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active", "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_lsu.sum", "sm__inst_executed_pipe_uniform.sum",
    "sm__warps_active.avg.pct_of_peak_sustained_active", "sm__maximum_warps_per_active_cycle_pct",
    "launch__registers_per_thread", "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem",
    "launch__grid_size", "launch__block_size", "launch__shared_mem_per_block_dynamic", "launch__waves_per_multiprocessor",
    "smsp__average_warp_latency_per_inst_issued.ratio",
    "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_dispatch_stall_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio",
    "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
]


def launches(path):
    rows = list(csv.reader(open(path)))
    hi = [i for i, r in enumerate(rows) if r and r[0] == "ID"][0]
    hdr, data = rows[hi], rows[hi + 1:]
    kn, mv = hdr.index("Kernel Name"), hdr.index("Metric Value")
    agg = collections.OrderedDict()
    for r in data:
        if len(r) > mv:
            agg.setdefault(r[kn].split("(")[0][:90], []).append(float(r[mv].replace(",", "")))
    tot = sum(sum(v) for v in agg.values())
    print(f"# ncu --metrics gpu__time_duration.sum --clock-control none : {path}")
    print(f"# {sum(len(v) for v in agg.values())} launches, {tot / 1e3:.1f} us total (cold-cache, serialised: compare SHARES)")
    print(f"{'kernel':92s} {'n':>5s} {'sum_us':>10s} {'avg_us':>9s} {'share':>6s}")
    for k, v in sorted(agg.items(), key=lambda kv: -sum(kv[1])):
        print(f"{k:92s} {len(v):5d} {sum(v) / 1e3:10.1f} {sum(v) / len(v) / 1e3:9.2f} {sum(v) / tot:6.3f}")


def full(path):
    out = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    hdr, units = rows[0], rows[1]
    print(f"# ncu --set full --clock-control none --import-source on : {path}")
    for r in rows[2:]:
        print(f"## {r[hdr.index('Kernel Name')]}  grid={r[hdr.index('Grid Size')]} block={r[hdr.index('Block Size')]}")
        for k in KEYS:
            if k in hdr:
                i = hdr.index(k)
                print(f"  {k:82s} {r[i]:>18s} {units[i]}")
        try:
            rd = float(r[hdr.index("dram__bytes_read.sum")].replace(",", ""))
            wr = float(r[hdr.index("dram__bytes_write.sum")].replace(",", ""))
            mult = {"Mbyte": 1e6, "Gbyte": 1e9, "Kbyte": 1e3, "byte": 1.0}
            rd *= mult[units[hdr.index("dram__bytes_read.sum")]]
            wr *= mult[units[hdr.index("dram__bytes_write.sum")]]
            print(f"  {'traffic (dram read+write) bytes per launch':82s} {rd + wr:18.0f}")
        except Exception:
            pass


if __name__ == "__main__":
    {"launches": launches, "full": full}[sys.argv[1]](sys.argv[2])
