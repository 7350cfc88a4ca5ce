#!/usr/bin/env python
"""Instruction mix per kernel from `cuobjdump -sass` (static counts; loops count once).
Usage: python tools/sass_mix.py <obj|so> [kernel-name-regex]"""
import collections
import re
import subprocess
import sys


def main():
    path = sys.argv[1]
    pat = re.compile(sys.argv[2]) if len(sys.argv) > 2 else None
    txt = subprocess.run(["cuobjdump", "-sass", path], capture_output=True, text=True).stdout
    demangle = {}
    cur, mix = None, {}
    for line in txt.splitlines():
        m = re.search(r"Function : (\S+)", line)
        if m:
            cur = m.group(1)
            mix[cur] = collections.Counter()
            continue
        m = re.match(r"\s+/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)", line)
        if m and cur:
            mix[cur][m.group(1).split(".")[0]] += 1
    names = list(mix)
    dm = subprocess.run(["cu++filt"] + names, capture_output=True, text=True).stdout.splitlines()
    for n, d in zip(names, dm):
        if pat and not pat.search(d):
            continue
        c = mix[n]
        tot = sum(v for k, v in c.items() if k not in ("NOP",))
        fp = {k: c[k] for k in ("FFMA", "FADD", "FMUL", "FFMA2", "FADD2", "FMUL2") if c[k]}
        rest = {k: v for k, v in c.most_common(14) if k not in fp and k != "NOP"}
        print(f"{d[:110]}\n   total {tot}  fp {fp}\n   other {rest}")


if __name__ == "__main__":
    main()
