#!/usr/bin/env python
"""Split the stall samples and executed instructions of a `--set full --import-source on` ncu report at the kernel's
CTA barriers (BAR.SYNC): one line per phase with its share of samples, instructions and top stall reasons.
Usage: python tools/ncu_phases.py report.ncu-rep"""
import collections
import csv
import subprocess
import sys


def main():
    rep = sys.argv[1]
    out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source=sass"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hi = next(i for i, r in enumerate(rows) if r and r[0] == "Address")
    hdr = rows[hi]
    ix = {h: i for i, h in enumerate(hdr)}
    body = [r for r in rows[hi + 1:] if len(r) == len(hdr)]
    bars = [i for i, r in enumerate(body) if "BAR.SYNC" in r[ix["Source"]]]
    S = lambda a, b, col="# Samples": sum(int(r[ix[col]] or 0) for r in body[a:b])
    stall_cols = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
    edges = [0] + bars + [len(body)]
    total = S(0, len(body))
    for a, b in zip(edges[:-1], edges[1:]):
        tot = S(a, b)
        d = {c[6:]: S(a, b, c) for c in stall_cols}
        d = {k: round(v / max(tot, 1), 3) for k, v in sorted(d.items(), key=lambda kv: -kv[1]) if v > tot * 0.04}
        ops = collections.Counter()
        for r in body[a:b]:
            src = r[ix["Source"]].split()
            op = (src[1] if src[0].startswith("@") else src[0]).split(".")[0]
            ops[op] += int(r[ix["Instructions Executed"]] or 0)
        ninst = sum(ops.values())
        print(f"sass lines {a}-{b}: {100 * tot / max(total, 1):.1f} % of samples, {ninst} warp-instructions; stalls {d}")
        print("    opcodes:", {k: v for k, v in ops.most_common(12)})


if __name__ == "__main__":
    main()
