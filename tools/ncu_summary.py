#!/usr/bin/env python
"""Summarise ncu artefacts brought back in gpurun_out/ into small text files
under profiles/ (the tracked evidence).

  ncu_summary.py launches <launches.csv> <out.md>     per-kernel totals/shares
  ncu_summary.py full <report.ncu-rep> <out.md>       key metrics per launch
"""
import csv
import io
import subprocess
import sys
from collections import OrderedDict

KEYS = [
    "gpu__time_duration.sum",
    "dram__bytes_read.sum", "dram__bytes_write.sum",
    "dram__throughput.avg.pct_of_peak_sustained_elapsed",
    "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
    "lts__t_bytes.sum",
    "sm__throughput.avg.pct_of_peak_sustained_elapsed",
    "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active",
    "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_tensor.sum",
    "sm__pipe_tensor_cycles_active_realtime.avg.pct_of_peak_sustained_elapsed",
    "smsp__pipe_tensor_subpipe_dmma_cycles_active.avg",
    "sm__cycles_elapsed.avg",
    "sm__warps_active.avg.pct_of_peak_sustained_active",
    "sm__maximum_warps_per_active_cycle_pct",
    "launch__registers_per_thread", "launch__grid_size", "launch__block_size",
    "launch__shared_mem_per_block_dynamic", "launch__shared_mem_per_block_static",
    "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem",
    "launch__waves_per_multiprocessor",
    "smsp__cycles_active.avg", "smsp__inst_executed.sum",
    "smsp__average_warp_latency_issue_stalled_long_scoreboard.pct",
    "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_no_instruction_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_membar_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_branch_resolving_per_issue_active.ratio",
    "smsp__pipe_tensor_subpipe_dmma_cycles_active.avg.pct_of_peak_sustained_active",
    "launch__cluster_size", "launch__cluster_max_active",
    "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
    "local_load_bytes", "smsp__inst_executed_op_local_ld.sum",
    "smsp__inst_executed_op_local_st.sum",
]


def short(name):
    name = name.replace("void ", "").replace("<unnamed>::", "")
    i = name.find("(")
    return name[:i] if i > 0 else name


def launches(path, out):
    rows = []
    with open(path, newline="") as fh:
        lines = [l for l in fh if l.startswith('"')]
    rd = csv.DictReader(io.StringIO("".join(lines)))
    for r in rd:
        if r.get("Metric Name") == "gpu__time_duration.sum":
            rows.append((r["Kernel Name"], float(r["Metric Value"]),
                         r.get("Metric Unit", "ns")))
    agg = OrderedDict()
    for k, v, u in rows:
        scale = {"ns": 1e-6, "us": 1e-3, "ms": 1.0}.get(u, 1e-6)
        a = agg.setdefault(short(k), [0, 0.0])
        a[0] += 1
        a[1] += v * scale
    tot = sum(a[1] for a in agg.values())
    with open(out, "w") as fh:
        fh.write(f"# launch list summary of {path}\n\n")
        fh.write(f"{len(rows)} launches, {tot:.3f} ms of kernel time "
                 "(ncu-serialised, cold cache: compare SHARES)\n\n")
        fh.write("| kernel | launches | total ms | share |\n|---|---|---|---|\n")
        for k, (c, t) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
            fh.write(f"| `{k}` | {c} | {t:.3f} | {100 * t / tot:.1f}% |\n")


def full(path, out):
    txt = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"],
                         capture_output=True, text=True).stdout
    rd = list(csv.reader(io.StringIO(txt)))
    hdr, units, body = rd[0], rd[1], rd[2:]
    with open(out, "w") as fh:
        fh.write(f"# ncu --set full summary of {path}\n\n")
        for r in body:
            fh.write(f"## {short(r[hdr.index('Kernel Name')])}  "
                     f"(id {r[0]})\n\n| metric | value | unit |\n|---|---|---|\n")
            for k in KEYS:
                # some metrics carry a section prefix ("TPC.TriageCompute.<name>")
                hits = [i for i, h in enumerate(hdr) if h == k or h.endswith("." + k)]
                if hits:
                    i = hits[0]
                    fh.write(f"| {k} | {r[i]} | {units[i]} |\n")
            fh.write("\n")


if __name__ == "__main__":
    {"launches": launches, "full": full}[sys.argv[1]](sys.argv[2], sys.argv[3])
