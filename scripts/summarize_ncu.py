#!/usr/bin/env python3
"""Summarise ncu CSV exports (gpurun_out/*_raw.csv, launches.csv) into profiles/<round>/.

    python scripts/summarize_ncu.py gpurun_out profiles/r01 [tag]
"""
import collections
import csv
import os
import shutil
import sys

KEYS = [
    "gpu__time_duration.sum", "launch__registers_per_thread", "launch__waves_per_multiprocessor",
    "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem",
    "sm__warps_active.avg.pct_of_peak_sustained_active",
    "TPC.TriageCompute.sm__pipe_tensor_cycles_active_realtime.avg.pct_of_peak_sustained_elapsed",
    "TPC.TriageCompute.sm__pipe_fp64_cycles_active_realtime.avg.pct_of_peak_sustained_elapsed",
    "sm__throughput.avg.pct_of_peak_sustained_elapsed", "smsp__issue_active.avg.pct_of_peak_sustained_active",
    "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
    "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed",
    "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "smsp__inst_executed.sum",
    "smsp__average_warps_issue_stalled_no_instruction_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio",
]


def launches_table(path):
    rows = list(csv.reader(open(path)))
    hdr, agg = None, collections.OrderedDict()
    for r in rows:
        if len(r) > 10 and r[0] == "ID":
            hdr = r
            continue
        if hdr and len(r) == len(hdr):
            d = dict(zip(hdr, r))
            name = d["Kernel Name"].split("(")[0]
            v = float(d["Metric Value"].replace(",", ""))
            v = v / 1e6 if d["Metric Unit"] == "ns" else (v / 1e3 if d["Metric Unit"] == "us" else v)
            a = agg.setdefault(name, [0, 0.0])
            a[0] += 1
            a[1] += v
    tot = sum(a[1] for a in agg.values())
    out = ["| kernel | launches | total ms | ms/launch | share |", "|---|---|---|---|---|"]
    for k, a in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        if a[1] / tot < 0.001:
            continue
        out.append(f"| `{k[:70]}` | {a[0]} | {a[1]:.2f} | {a[1] / a[0]:.4f} | {a[1] / tot * 100:.1f}% |")
    return "\n".join(out)


def main():
    src, dst = sys.argv[1], sys.argv[2]
    tag = sys.argv[3] if len(sys.argv) > 3 else "ncu"
    os.makedirs(dst, exist_ok=True)
    md = [f"# ncu summary ({tag})", "",
          "Command: `python bench.py --steps 2 --warmup 3 --rounds-per-step 2 --chains 65536 --no-e2e --no-cpu-baseline`"
          " (German-shaped, 65536 chains), `--clock-control none`.", "",
          "## Launch list (gpu__time_duration, cold-cache, serialised: compare shares)", ""]
    lp = os.path.join(src, "launches.csv")
    if os.path.isfile(lp):
        md.append(launches_table(lp))
        shutil.copy(lp, os.path.join(dst, f"{tag}_launches.csv"))
    md += ["", "## Full-set captures (one launch each)", ""]
    for fn in sorted(os.listdir(src)):
        if not fn.endswith("_raw.csv") or os.path.getsize(os.path.join(src, fn)) < 1000:
            continue
        rows = list(csv.reader(open(os.path.join(src, fn))))
        hdr, units, vals = rows[0], rows[1], rows[2]
        name = fn[:-8]
        kn = vals[hdr.index("Kernel Name")] if "Kernel Name" in hdr else name
        md += [f"### {name}: `{kn[:90]}`", "", "| metric | value | unit |", "|---|---|---|"]
        for k in KEYS:
            if k in hdr:
                i = hdr.index(k)
                md.append(f"| {k} | {vals[i]} | {units[i]} |")
        md.append("")
        shutil.copy(os.path.join(src, fn), os.path.join(dst, f"{tag}_{fn}"))
        det = os.path.join(src, name + "_details.csv")
        if os.path.isfile(det):
            shutil.copy(det, os.path.join(dst, f"{tag}_{name}_details.csv"))
    open(os.path.join(dst, f"{tag}_summary.md"), "w").write("\n".join(md) + "\n")
    print("\n".join(md))


if __name__ == "__main__":
    main()
