#!/usr/bin/env python
"""profiles/chain_traffic.json from an `ncu --set full` report of bench.py's chain kernel: DRAM bytes read + written by ONE
launch, tied to the sha256 of the kernel's sources so that bench.py reports `roofline.traffic` only while the capture still
describes the kernel it times.  Usage: python tools/chain_traffic.py gpurun_out/<name>.ncu-rep [capture-label]"""
import csv
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402


def main():
    rep = sys.argv[1]
    label = sys.argv[2] if len(sys.argv) > 2 else os.path.basename(rep)
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr, units = rows[0], rows[1]
    for vals in rows[2:]:
        name = vals[hdr.index("Kernel Name")]
        if "chain_kernel" not in name:
            continue
        def get(metric):
            i = hdr.index(metric)
            v, u = float(vals[i]), units[i].lower()
            return v * {"byte": 1, "kbyte": 1e3, "mbyte": 1e6, "gbyte": 1e9}[u]
        rd, wr = get("dram__bytes_read.sum"), get("dram__bytes_write.sum")
        js = {"dram_bytes_per_launch": rd + wr, "dram_bytes_read": rd, "dram_bytes_write": wr, "kernel": name[:60],
              "capture": label, "kernel_source_sha256": bench.chain_source_sha(),
              "how": "ncu --set full --clock-control none -k regex:chain_kernel -c 1 python bench.py --steps 2 --warmup 3 --no-cpu --no-extra"}
        with open(os.path.join(ROOT, "profiles", "chain_traffic.json"), "w") as f:
            json.dump(js, f, indent=1)
        print(json.dumps(js))
        return
    raise SystemExit("no chain_kernel launch in the report")


if __name__ == "__main__":
    main()
