#!/usr/bin/env python
"""Hot SASS instructions of a `--set full --import-source on` ncu report: the instructions that collect the most
stall samples, with the dominant stall reason.  Usage: python tools/ncu_hot.py report.ncu-rep [N]"""
import csv
import subprocess
import sys


def main():
    rep = sys.argv[1]
    top = int(sys.argv[2]) if len(sys.argv) > 2 else 40
    out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source=sass"],
                         capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hi = next(i for i, r in enumerate(rows) if r and r[0] == "Address")
    hdr = rows[hi]
    ix = {h: i for i, h in enumerate(hdr)}
    stall_cols = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
    body = []
    for r in rows[hi + 1:]:                    # first kernel of the report only
        if r and r[0] in ("Kernel Name", "Address"):
            break
        if len(r) == len(hdr):
            body.append(r)
    tot = sum(int(r[ix["# Samples"]] or 0) for r in body)
    agg = {}
    for r in body:
        for c in stall_cols:
            agg[c] = agg.get(c, 0) + int(r[ix[c]] or 0)
    print("total samples", tot, {k: v for k, v in sorted(agg.items(), key=lambda kv: -kv[1]) if v})
    ranked = sorted(range(len(body)), key=lambda i: -int(body[i][ix["# Samples"]] or 0))[:top]
    for i in sorted(ranked):
        r = body[i]
        st = sorted(((int(r[ix[c]] or 0), c) for c in stall_cols), reverse=True)[:2]
        print(f"{i:5d} {int(r[ix['# Samples']]):6d}  {r[ix['Source']][:90]:90s} {st}")


if __name__ == "__main__":
    main()
