#!/usr/bin/env python
"""Sustained behaviour of the headline chain kernel under the power cap: per-window mean launch time and SM clock over a few
seconds of back-to-back launches (bench.py times 20 launches on a GPU that has just warmed up).  Usage: python tools/sustained_chain.py [seconds]"""
import json
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from libredio_b200 import blocks, synth  # noqa: E402


def main():
    seconds = float(sys.argv[1]) if len(sys.argv) > 1 else 6.0
    ctx = blocks.Context(0)
    dev = ctx.tdev
    frames, k = 65536, 64
    n = frames * 10240 + 54
    g = torch.Generator(device=dev).manual_seed(0)
    x = torch.view_as_complex(torch.randn(n, 2, device=dev, generator=g))
    ch = blocks.Chain(ctx, synth.lpf_taps(64, 0.04), 10, 1024)
    out = torch.empty((frames // k, 1024), dtype=torch.float32, device=dev)
    try:
        import pynvml
        pynvml.nvmlInit()
        h = pynvml.nvmlDeviceGetHandleByIndex(0)
        clk = lambda: (pynvml.nvmlDeviceGetClockInfo(h, pynvml.NVML_CLOCK_SM), pynvml.nvmlDeviceGetClockInfo(h, pynvml.NVML_CLOCK_MEM),
                       pynvml.nvmlDeviceGetPowerUsage(h) / 1000.0, pynvml.nvmlDeviceGetTemperature(h, 0))
    except Exception:
        clk = lambda: (0, 0, 0.0, 0)
    for _ in range(3):
        ch.run(x, k, out)
    torch.cuda.synchronize()
    t_end = time.perf_counter() + seconds
    w = 0
    while time.perf_counter() < t_end:
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(100):
            ch.run(x, k, out)
        e1.record()
        torch.cuda.synchronize()
        sm, mem, pw, tc = clk()
        print(json.dumps({"window": w, "ms_per_launch": e0.elapsed_time(e1) / 100, "sm_mhz": sm, "mem_mhz": mem, "power_w": pw, "temp_c": tc}), flush=True)
        w += 1
    ch.close(); ctx.close()


if __name__ == "__main__":
    main()
