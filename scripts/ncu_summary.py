"""Condense ncu output into the small text summaries kept under profiles/.

    python scripts/ncu_summary.py launches gpurun_out/launches.csv > profiles/launches_rNN.txt
    python scripts/ncu_summary.py full gpurun_out/prof.ncu-rep > profiles/fused_full_rNN.txt
"""
import collections
import csv
import subprocess
import sys

KEEP = [
    "gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
    "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem", "launch__shared_mem_per_block_dynamic",
    "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__issue_active.avg.pct_of_peak_sustained_active",
    "smsp__inst_executed.sum", "smsp__thread_inst_executed_per_inst_executed.ratio",
    "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
    "sm__throughput.avg.pct_of_peak_sustained_elapsed", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
    "dram__bytes_read.sum", "dram__bytes_write.sum", "lts__t_sector_hit_rate.pct", "l1tex__t_sector_hit_rate.pct",
    "sass__inst_executed_register_spilling", "smsp__warps_eligible.avg.per_cycle_active",
    "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
]


def launches(path, exclude=()):
    """`exclude`: substrings of kernel names left out of the total (diagnostics that run after the timed regions of the
    profiled command, e.g. the FMA peak measurement and the generation of the synthetic cloud)."""
    lines = [l for l in open(path) if not l.startswith("==")]
    agg = collections.OrderedDict()
    skipped = collections.OrderedDict()
    for row in csv.DictReader(lines):
        try:
            v = float(row["Metric Value"].replace(",", ""))
        except (KeyError, ValueError):
            continue
        unit = row["Metric Unit"]
        v = v / 1e3 if unit == "ns" else v * 1e3 if unit == "ms" else v * 1e6 if unit == "s" else v
        name = row["Kernel Name"][:110]
        a = (skipped if any(x in name for x in exclude) else agg).setdefault(name, [0, 0.0])
        a[0] += 1
        a[1] += v
    tot = sum(a[1] for a in agg.values())
    print(f"# per-kernel device time (us), cold-cache serialised launches under ncu; total {tot:.1f} us")
    for k, (c, t) in skipped.items():
        print(f"# not in the total (runs outside the timed regions): {t:10.1f} us {c:4d}x  {k[:80]}")
    for k, (c, t) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        print(f"{t:12.1f} us {c:5d}x {100 * t / tot:6.2f}%  {k}")


def full(path):
    out = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr, units, data = rows[0], rows[1], rows[2:]
    name_col = hdr.index("Kernel Name")
    for r in data:
        print(f"## {r[name_col][:120]}")
        for i, h in enumerate(hdr):
            if h in KEEP or ("issue_stalled" in h and h.endswith("per_issue_active.ratio")):
                print(f"{h:90s} {units[i]:16s} {r[i]}")
        print()


if __name__ == "__main__":
    if sys.argv[1] == "launches":
        launches(sys.argv[2], tuple(sys.argv[3:]))
    else:
        full(sys.argv[2])
