"""Turns ncu outputs brought back in gpurun_out/ into the text summaries committed under profiles/.

    python tools/ncu_summary.py launches gpurun_out/launches.csv  > profiles/rN_launches.txt
    python tools/ncu_summary.py kernel   gpurun_out/prof.ncu-rep  > profiles/rN_kernel.txt
"""
import csv
import io
import subprocess
import sys
from collections import OrderedDict

RAW_KEYS = [
    "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
    "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_bytes.sum", "lts__t_sector_hit_rate.pct",
    "l1tex__t_sector_hit_rate.pct", "l1tex__t_bytes.sum", "launch__registers_per_thread", "launch__grid_size", "launch__block_size",
    "launch__shared_mem_per_block_static", "launch__shared_mem_per_block_dynamic", "launch__occupancy_limit_registers",
    "launch__waves_per_multiprocessor", "sm__warps_active.avg.pct_of_peak_sustained_active",
    "smsp__thread_inst_executed_per_inst_executed.ratio", "smsp__inst_executed.sum",
    "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
    "smsp__cycles_active.avg", "sm__cycles_elapsed.max", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_tensor.sum",
    "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_no_instruction_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_branch_resolving_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_membar_per_issue_active.ratio",
]


def launches(path):
    rows = [r for r in csv.reader(open(path)) if len(r) > 5]
    hdr = rows[0]
    ki, vi, ui = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
    agg = OrderedDict()
    for r in rows[1:]:
        v = float(r[vi].replace(",", ""))
        v = {"ns": v / 1000, "ms": v * 1000}.get(r[ui], v)
        a = agg.setdefault(r[ki][:110], [0, 0.0])
        a[0] += 1
        a[1] += v
    tot = sum(a[1] for a in agg.values())
    print(f"# {path}: {len(rows) - 1} launches, {tot / 1000:.2f} ms of device time (ncu: cold cache, serialised -- compare shares)")
    print(f"{'launches':>8} {'total_us':>12} {'share':>7}  kernel")
    for k, a in sorted(agg.items(), key=lambda x: -x[1][1]):
        print(f"{a[0]:8d} {a[1]:12.1f} {100 * a[1] / tot:6.1f}%  {k}")


def kernel(path, top=25):
    raw = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr, units = rows[0], rows[1]
    print(f"# {path}: ncu --set full, {len(rows) - 2} profiled launch(es)")
    for n, r in enumerate(rows[2:]):
        print(f"\n## launch {n}: {r[hdr.index('Kernel Name')]}")
        for k in RAW_KEYS:
            if k in hdr:
                i = hdr.index(k)
                print(f"{k:85s} {r[i]:>16s} {units[i]}")
    src = subprocess.run(["ncu", "-i", path, "--page", "source", "--csv", "--print-source", "cuda,sass"],
                         capture_output=True, text=True).stdout
    cur, agg = None, {}
    for r in csv.reader(io.StringIO(src)):
        if len(r) >= 2 and r[0] == "File Path":
            cur = r[1].split("/")[-1]
            continue
        if len(r) < 8 or r[0] in ("Line No", "Function Name", ""):
            continue
        try:
            a = agg.setdefault((cur, int(r[0]), r[1].strip()), [0, 0])
            a[0] += int(r[4])
            a[1] += int(r[7])
        except ValueError:
            pass
    ts, ti = max(1, sum(a[0] for a in agg.values())), max(1, sum(a[1] for a in agg.values()))
    print(f"\n## source lines by warp-stall samples (all profiled launches; {ts} samples, {ti} warp instructions)")
    for k, a in sorted(agg.items(), key=lambda x: -x[1][0])[:top]:
        print(f"{k[0]}:{k[1]:<5d} samples {100 * a[0] / ts:5.1f}%  inst {100 * a[1] / ti:5.1f}%  {k[2][:110]}")
    print("\n## source lines by executed warp instructions")
    for k, a in sorted(agg.items(), key=lambda x: -x[1][1])[:top]:
        print(f"{k[0]}:{k[1]:<5d} inst {100 * a[1] / ti:5.1f}%  samples {100 * a[0] / ts:5.1f}%  {k[2][:110]}")


if __name__ == "__main__":
    {"launches": launches, "kernel": kernel}[sys.argv[1]](sys.argv[2])
