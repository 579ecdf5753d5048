"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list by kernel: launches, average / total time, share.

    python tools/launch_summary.py gpurun_out/x/launches.csv [--skip N] > summary.csv
"""
import csv
import re
import sys
from collections import OrderedDict


def main():
    path = sys.argv[1]
    skip = int(sys.argv[sys.argv.index("--skip") + 1]) if "--skip" in sys.argv else 0
    rows = []
    with open(path) as f:
        lines = [l for l in f if not l.startswith("==")]
    for r in csv.DictReader(lines):
        if r.get("Metric Name") != "gpu__time_duration.sum":
            continue
        v = float(r["Metric Value"].replace(",", ""))
        unit = r.get("Metric Unit", "ns")
        us = v / 1e3 if unit in ("ns", "nsecond") else v if unit in ("us", "usecond") else v * 1e3
        name = re.sub(r"\(.*$", "", r["Kernel Name"])
        rows.append((name, r.get("Grid Size", ""), r.get("Block Size", ""), us))
    rows = rows[skip:]
    agg = OrderedDict()
    for name, g, b, us in rows:
        k = (name, g, b)
        n, t = agg.get(k, (0, 0.0))
        agg[k] = (n + 1, t + us)
    tot = sum(t for _, t in agg.values())
    print("kernel,grid,block,launches,avg_us,total_us,share_pct")
    for (name, g, b), (n, t) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        print(f'"{name}","{g}","{b}",{n},{t / n:.2f},{t:.1f},{100 * t / tot:.2f}')
    print(f'"TOTAL","","",{sum(n for n, _ in agg.values())},,{tot:.1f},100.00')


if __name__ == "__main__":
    main()
