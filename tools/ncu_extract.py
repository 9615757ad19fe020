#!/usr/bin/env python
"""Condense ncu output into the small, diffable summaries kept under profiles/.

  ncu_extract.py launches <launch-list.csv>        per-kernel totals of a `--metrics gpu__time_duration.sum` launch list
  ncu_extract.py full <report.ncu-rep> [...]       key metrics of `--set full` captures (needs `ncu` on PATH)
"""
import collections
import csv
import subprocess
import sys

KEYS = [
    "Kernel Name", "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
    "dram__throughput.avg.pct_of_peak_sustained_elapsed", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
    "lts__t_sector_hit_rate.pct", "lts__t_sectors_srcunit_tex_op_read.sum",
    "TPC.TriageCompute.sm__pipe_tensor_cycles_active_realtime.avg.pct_of_peak_sustained_elapsed",
    "sm__mem_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
    "sm__inst_executed.sum", "smsp__inst_executed.sum", "sm__inst_issued.avg.pct_of_peak_sustained_active",
    "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__warps_active.avg.pct_of_peak_sustained_active",
    "launch__registers_per_thread", "launch__grid_size", "launch__block_size", "launch__occupancy_limit_registers",
    "launch__waves_per_multiprocessor", "sm__cycles_elapsed.max", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
    "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "smsp__average_warp_latency_issue_stalled_long_scoreboard.ratio",
    "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio",
    "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed", "l1tex__data_pipe_tc_wavefronts_mem_shared.sum",
    "l1tex__data_pipe_tc_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed", "l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed",
    "l1tex__m_xbar2l1tex_read_bytes_mem_global_op_tma_ld.sum", "sm__cycles_elapsed.max.per_second",
    "smsp__sass_inst_executed_op_tmem_ldt.sum", "smsp__inst_executed_op_shared_atom.sum", "sm__inst_executed_pipe_fp64.sum",
    "smsp__sass_thread_inst_executed_op_fadd_pred_on.sum", "smsp__sass_thread_inst_executed_op_ffma_pred_on.sum",
]


def launches(path):
    with open(path) as f:
        lines = [l for l in f if not l.startswith("==")]
    agg = collections.OrderedDict()
    n = 0
    for row in csv.DictReader(lines):
        if row.get("Metric Name") != "gpu__time_duration.sum":
            continue
        v = float(row["Metric Value"].replace(",", ""))
        v *= {"ns": 1e-3, "us": 1.0, "ms": 1e3, "s": 1e6}[row["Metric Unit"]]
        a = agg.setdefault(row["Kernel Name"], [0, 0.0])
        a[0] += 1
        a[1] += v
        n += 1
    tot = sum(a[1] for a in agg.values())
    print(f"# {path}: {n} launches, {tot / 1e3:.3f} ms total (ncu per-launch times: cold caches, serialised)")
    print("launches,total_ms,share_pct,avg_us,kernel")
    for k, a in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        print(f"{a[0]},{a[1] / 1e3:.3f},{100 * a[1] / tot:.2f},{a[1] / a[0]:.1f},\"{k[:140]}\"")


def full(paths):
    for path in paths:
        out = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
        rows = list(csv.reader(out.splitlines()))
        hdr, units = rows[0], rows[1]
        for vals in rows[2:]:
            d = dict(zip(hdr, zip(units, vals)))
            print(f"# {path}")
            for k in KEYS:
                if k in d:
                    print(f"{k},{d[k][1][:150]},{d[k][0]}")


if __name__ == "__main__":
    if sys.argv[1] == "launches":
        launches(sys.argv[2])
    else:
        full(sys.argv[2:])
