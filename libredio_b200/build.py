"""Build libredio_cuda.so in-tree with nvcc for sm_100a (B200).  No torch, no JIT cache:

    python -m libredio_b200.build            # incremental
    python -m libredio_b200.build --force

Objects go to libredio_b200/build/, the library to libredio_b200/libredio_cuda.so (git-ignored, but it
travels with the repo snapshot to the GPU box).  ptxas -v output is kept next to each object.
"""
from __future__ import annotations

import concurrent.futures as cf
import json
import os
import shutil
import subprocess
import sys
import time

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
BUILD = os.path.join(HERE, "build")
LIB = os.path.join(HERE, "libredio_cuda.so")
SHIMS = {
    "libkissfft.so": ["shim_kissfft.cpp"],
    "libsamplerate.so": ["shim_samplerate.cpp"],
}

NVCC = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
ARCH = ["-gencode", "arch=compute_100a,code=sm_100a"]
CFLAGS = ["-O3", "-std=c++17", "-lineinfo", "-Xcompiler", "-fPIC,-Wall", "-Xptxas", "-v",
          "--expt-relaxed-constexpr", "-I", os.path.join(HERE, "..", "include")]


_TIMES: dict[str, float] = {}
# longest translation units first, so that a fresh build's critical path is one long file and not a long file started last
# (seconds measured with 8 workers on the build container; unknown files go in front)
_COST_FILE = os.path.join(HERE, "build_order.json")


def _newer(target: str, deps: list[str]) -> bool:
    if not os.path.exists(target):
        return True
    t = os.path.getmtime(target)
    return any(os.path.getmtime(d) > t for d in deps)


def _compile(src: str, headers: list[str], force: bool) -> str:
    obj = os.path.join(BUILD, os.path.basename(src).rsplit(".", 1)[0] + ".o")
    if force or _newer(obj, [src] + headers):
        cmd = [NVCC, *ARCH, *CFLAGS, "-c", src, "-o", obj]
        t0 = time.perf_counter()
        r = subprocess.run(cmd, capture_output=True, text=True)
        _TIMES[os.path.basename(src)] = round(time.perf_counter() - t0, 1)
        with open(obj + ".ptxas.log", "w") as f:
            f.write(" ".join(cmd) + "\n" + r.stdout + r.stderr)
        if r.returncode != 0:
            raise RuntimeError(f"nvcc failed for {src}:\n{r.stdout}\n{r.stderr}")
    return obj


def build(force: bool = False, verbose: bool = False) -> str:
    os.makedirs(BUILD, exist_ok=True)
    headers = [os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith((".cuh", ".h", ".hpp"))]
    headers.append(os.path.join(HERE, "..", "include", "libredio_cuda.h"))
    srcs = sorted(os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith(".cu"))
    try:
        with open(_COST_FILE) as fh:
            cost = json.load(fh)
    except Exception:
        cost = {}
    srcs.sort(key=lambda p: -cost.get(os.path.basename(p), 1e9))
    with cf.ThreadPoolExecutor(max_workers=min(8, len(srcs))) as ex:
        objs = list(ex.map(lambda s: _compile(s, headers, force), srcs))
    if force or _newer(LIB, objs):
        cmd = [NVCC, *ARCH, "-shared", "-o", LIB, *objs, "-cudart", "static"]
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError(f"link failed:\n{r.stdout}\n{r.stderr}")
    # ABI shims that sit behind the reference's existing FFI seams (kissfft.rs:11, samplerate.rs:32)
    for name, files in SHIMS.items():
        paths = [os.path.join(CSRC, f) for f in files]
        if not all(os.path.exists(p) for p in paths):
            continue
        out = os.path.join(HERE, name)
        if force or _newer(out, paths + [LIB]):
            cmd = ["g++", "-O2", "-std=c++17", "-fPIC", "-shared", "-Wall", "-I", os.path.join(HERE, "..", "include"),
                   "-o", out, *paths, "-L", HERE, "-lredio_cuda", "-Wl,-rpath,$ORIGIN"]
            r = subprocess.run(cmd, capture_output=True, text=True)
            if r.returncode != 0:
                raise RuntimeError(f"shim {name} failed:\n{r.stdout}\n{r.stderr}")
    if _TIMES and os.environ.get("LRC_BUILD_RECORD_TIMES"):
        cost.update(_TIMES)
        with open(_COST_FILE, "w") as fh:
            json.dump(dict(sorted(cost.items(), key=lambda kv: -kv[1])), fh, indent=1)
    if verbose:
        for o in objs:
            with open(o + ".ptxas.log") as f:
                sys.stdout.write(f.read())
    return LIB


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
