#!/usr/bin/env python
"""Write the SASS of one kernel (first function whose mangled name contains every given substring) to a listing
with an opcode histogram header.  Usage: python tools/sass_dump.py <obj|so> <out.sass> <substr> [substr...]"""
import collections
import re
import subprocess
import sys


def main():
    path, out, subs = sys.argv[1], sys.argv[2], sys.argv[3:]
    txt = subprocess.run(["cuobjdump", "-sass", path], capture_output=True, text=True).stdout
    cur, keep, name = None, [], None
    for line in txt.splitlines():
        m = re.search(r"Function : (\S+)", line)
        if m:
            if keep:
                break
            cur = m.group(1)
            if all(s in cur for s in subs):
                name = cur
            continue
        if name and cur == name and re.match(r"\s+/\*[0-9a-f]{4,}\*/", line):
            keep.append(re.sub(r"\s*/\* 0x[0-9a-f]+ \*/\s*$", "", line))
    if not keep:
        sys.exit(f"no function matching {subs}")
    hist = collections.Counter()
    for ln in keep:
        m = re.match(r"\s+/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_]+)", ln)
        if m:
            hist[m.group(1)] += 1
    with open(out, "w") as f:
        f.write(f"// {name}\n// SASS of sm_100a build (cuobjdump -sass), Opcode histogram: "
                + ", ".join(f"{k}:{v}" for k, v in hist.most_common(16)) + "\n")
        f.write("\n".join(keep) + "\n")
    print(out, len(keep), "instructions")


if __name__ == "__main__":
    main()
