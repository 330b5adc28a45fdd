#!/usr/bin/env python
"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list: one training step =
the launches between the last two `k_mark` kernels (first kernel of geomae_voxel_scatter).

    python tools/summarize_launches.py gpurun_out/launches.csv [--md out.md] [--seq out_seq.txt]
"""
import argparse
import csv
import sys
from collections import OrderedDict


def read(path):
    rows = []
    with open(path, newline="") as f:
        lines = [ln for ln in f if not ln.startswith("==")]
    rd = csv.DictReader(lines)
    for r in rd:
        if r.get("Metric Name") != "gpu__time_duration.sum":
            continue
        val = float(r["Metric Value"].replace(",", ""))
        unit = r.get("Metric Unit", "ns")
        us = val / 1e3 if unit in ("ns", "nsecond") else val if unit in ("us", "usecond") else val * 1e3
        rows.append((r["Kernel Name"], us, r.get("Grid Size", ""), r.get("Block Size", ""), r.get("Stream", "")))
    return rows


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("csv")
    ap.add_argument("--md")
    ap.add_argument("--seq")
    ap.add_argument("--marker", default="k_mark")
    a = ap.parse_args()
    rows = read(a.csv)
    marks = [i for i, r in enumerate(rows) if a.marker in r[0].split("(")[0]]
    if len(marks) >= 2:
        lo, hi = marks[-2], marks[-1]
    else:
        lo, hi = 0, len(rows)
    step = rows[lo:hi]
    total = sum(r[1] for r in step)
    agg = OrderedDict()
    for name, us, *_ in step:
        short = name.split("(")[0][:100]
        t = agg.setdefault(short, [0.0, 0])
        t[0] += us
        t[1] += 1
    out = [f"one step = launches {lo}..{hi} of {len(rows)}: {len(step)} launches, {total / 1e3:.3f} ms summed device time", "",
           "| device time (us) | launches | avg (us) | share | kernel |", "|---:|---:|---:|---:|---|"]
    for name, (us, n) in sorted(agg.items(), key=lambda kv: -kv[1][0]):
        out.append(f"| {us:.1f} | {n} | {us / n:.1f} | {100 * us / total:.1f}% | `{name}` |")
    own = sum(us for name, (us, n) in agg.items() if name.startswith("k_") or "::k_" in name or "k_tc" in name)
    out.append("")
    out.append(f"hand-written kernels (`k_*`): {own:.0f} us = {100 * own / total:.1f}% of the step")
    text = "\n".join(out)
    print(text)
    if a.md:
        open(a.md, "w").write(text + "\n")
    if a.seq:
        with open(a.seq, "w") as f:
            for i, (name, us, grid, block, stream) in enumerate(step):
                f.write(f"{i:5d} {us:9.1f} us  grid {grid:>18s} block {block:>14s} stream {stream:>4s}  {name.split('(')[0][:90]}\n")


if __name__ == "__main__":
    sys.exit(main())
