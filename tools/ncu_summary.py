#!/usr/bin/env python
"""Summarise an `ncu --csv --metrics ...` launch list: per kernel name, launches and the median of every metric.
Usage: python tools/ncu_summary.py gpurun_out/launches.csv [> profiles/rN_xxx_summary.txt]"""
import csv
import statistics
import sys
from collections import OrderedDict, defaultdict


def main(path):
    rows = []
    with open(path, newline="") as f:
        lines = [ln for ln in f if ln.startswith('"')]
    rd = csv.DictReader(lines)
    per = OrderedDict()
    for r in rd:
        name = r["Kernel Name"].split("(")[0].replace("void ", "")
        key = (name, r["Grid Size"], r["Block Size"])
        per.setdefault(key, defaultdict(list))[r["Metric Name"] + " [" + r["Metric Unit"] + "]"].append(
            float(r["Metric Value"].replace(",", "")))
    for (name, grid, block), m in per.items():
        n = max(len(v) for v in m.values())
        print(f"{name}  grid={grid} block={block}  launches={n}")
        for k, v in m.items():
            print(f"    {k:70s} median {statistics.median(v):16.1f}   min {min(v):16.1f}")


if __name__ == "__main__":
    main(sys.argv[1])
