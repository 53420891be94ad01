"""Summarise ncu outputs (run here, no GPU): a launch list CSV -> per-kernel share table, and one or
more .ncu-rep captures -> the handful of metrics that matter (duration, tensor/XU/ALU/FMA pipe %,
DRAM bytes, registers, issue activity, top stall reasons).  Usage:
  python scripts/summarize_ncu.py --launches gpurun_out/launches_r01.csv --reps gpurun_out/attn_r01.ncu-rep ... > profiles/r01_summary.md
"""
import argparse
import collections
import csv
import io
import re
import subprocess


def launches(path):
    lines = [l for l in open(path) if not l.startswith("==")]
    tot, cnt = collections.Counter(), collections.Counter()
    for row in csv.DictReader(lines):
        if row.get("Metric Name") != "gpu__time_duration.sum":
            continue
        name = re.sub(r"\(.*", "", row["Kernel Name"]).replace("wvn::<unnamed>::", "").replace("void ", "")
        v = float(row["Metric Value"].replace(",", ""))
        v *= {"ns": 1, "us": 1e3, "ms": 1e6}.get(row["Metric Unit"], 1)
        tot[name] += v
        cnt[name] += 1
    T = sum(tot.values())
    print(f"### Launch list `{path}` — {sum(cnt.values())} launches, {T / 1e6:.3f} ms total (cold-cache, serialised: compare shares)\n")
    print("| kernel | launches | total ms | share | avg µs |\n|---|---:|---:|---:|---:|")
    for k, v in tot.most_common(24):
        print(f"| `{k[:70]}` | {cnt[k]} | {v / 1e6:.3f} | {100 * v / T:.1f}% | {v / cnt[k] / 1e3:.1f} |")
    print()


KEYS = [
    "gpu__time_duration.sum", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active", "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active", "smsp__issue_active.avg.pct_of_peak_sustained_active",
    "sm__warps_active.avg.pct_of_peak_sustained_active", "dram__bytes_read.sum", "dram__bytes_write.sum",
    "lts__throughput.avg.pct_of_peak_sustained_elapsed", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
    "launch__registers_per_thread", "launch__grid_size", "launch__block_size", "launch__waves_per_multiprocessor",
    "smsp__inst_executed.sum",
]


def rep(path):
    out = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    if len(rows) < 3:
        print(f"(could not read {path})")
        return
    hdr, units = rows[0], rows[1]
    print(f"### `{path}`\n")
    for r in rows[2:]:
        name = r[hdr.index("Kernel Name")]
        print(f"**{name[:100]}**\n")
        print("| metric | value | unit |\n|---|---:|---|")
        for k in KEYS:
            if k in hdr:
                i = hdr.index(k)
                print(f"| {k} | {r[i]} | {units[i]} |")
        stalls = [(float(r[i] or 0), h.replace("smsp__pcsamp_warps_issue_stalled_", "")) for i, h in enumerate(hdr)
                  if h.startswith("smsp__pcsamp_warps_issue_stalled_") and not h.endswith("_not_issued")]
        stalls.sort(reverse=True)
        print("\ntop stall samples: " + ", ".join(f"{n} {int(v)}" for v, n in stalls[:6]) + "\n")


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--launches")
    ap.add_argument("--reps", nargs="*", default=[])
    a = ap.parse_args()
    if a.launches:
        launches(a.launches)
    for p in a.reps:
        rep(p)
