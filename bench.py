#!/usr/bin/env python
"""bench.py -- BASELINE.json's metric on BASELINE.json's config.

metric : Msamples/s of cf32 through the FIR-decimate + FFT chain (64-tap low-pass /10 -> 1024-pt Hann FFT ->
         |X|^2 averaged over K=64 frames), absolute and as a fraction of the HBM roofline.
step   : one pass of the fused chain kernel over one batch of synthetic cf32 samples resident in HBM:
         65536 frames = 2^26 decimated samples into the FFT (configs[1]) = 671 088 694 input samples (5.4 GB),
         much larger than the 126 MB L2, so no L2 flush is needed between steps.
N > 1  : one process per GPU (torchrun); each rank owns an independent stream of the same size (weak
         scaling, channels sharded across GPUs); the only exchange is the gather of the 4 MB of output rows
         per rank over NVLink, inside the timed region: lrc_gather (copy-engine P2P pushes, no SMs) by
         default, NCCL all-gather with --gather nccl or when CUDA IPC is unavailable.
e2e    : the same step through the host-buffer entry point lrc_chain_run_host (pinned host input, chunked
         H2D overlapped with the kernel through the double-buffered device ring, rows copied back).
--impl reference : the reference's own CPU path (oracle/: strict-f32 restatement of dsputils::convolve +
         the vendored kissfft built with the reference's flags), all host cores, bounded sample.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

NTAPS, DECIM, NFFT, K_AVG = 64, 10, 1024, 64
FRAMES = 65536                      # 2^26 decimated samples / 1024
BYTES_PER_SAMPLE = 8.0              # compulsory traffic = first read of cf32 (+ 4/(10*K) B of rows)
METRIC = "Msamples/s cf32 through FIR-decimate+FFT chain"
UNIT = "Msamples/s"


def n_input(frames: int) -> int:
    return frames * NFFT * DECIM + NTAPS - DECIM


def workload_name(frames: int) -> str:
    return (f"cf32 {n_input(frames)} samples -> FIR{NTAPS}/{DECIM} -> {frames} frames x {NFFT}-pt Hann FFT "
            f"(2^{int(np.log2(frames * NFFT))} samples into the FFT) -> |X|^2 avg K={K_AVG}")


# ------------------------------------------------------------------------------------------------------
# clocks sampler (nvidia-smi during the timed region)
# ------------------------------------------------------------------------------------------------------
class ClockSampler:
    """SM clock and clock-event reasons sampled through NVML every ~2 ms from a thread (nvidia-smi -lms cannot
    deliver a sample inside a region that lasts tens of milliseconds); nvidia-smi is the fallback."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
         "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index: int, uuid: str | None = None):
        self.index, self.uuid = index, uuid
        self.sm, self.mx, self.pw, self.reasons = [], [], [], set()
        self.stop_flag = threading.Event()
        self.th = None
        self.how = None
        self.nv = None
        try:
            import pynvml
            pynvml.nvmlInit()
            h = None
            if uuid:
                for cand in (uuid, "GPU-" + uuid):
                    try:
                        h = pynvml.nvmlDeviceGetHandleByUUID(cand if isinstance(cand, bytes) else cand.encode())
                        break
                    except Exception:
                        h = None
            if h is None:
                h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.nv, self.h = pynvml, h
            self.max_sm = float(pynvml.nvmlDeviceGetMaxClockInfo(h, pynvml.NVML_CLOCK_SM))
            self.how = "nvml"
        except Exception:
            self.nv = None

    def _poll_nvml(self):
        nv, h = self.nv, self.h
        names = {nv.nvmlClocksEventReasonHwSlowdown: "hw_slowdown",
                 nv.nvmlClocksEventReasonHwThermalSlowdown: "hw_thermal_slowdown",
                 nv.nvmlClocksEventReasonSwThermalSlowdown: "sw_thermal_slowdown",
                 nv.nvmlClocksEventReasonSwPowerCap: "sw_power_cap",
                 nv.nvmlClocksEventReasonHwPowerBrakeSlowdown: "hw_power_brake_slowdown"}
        while not self.stop_flag.is_set():
            try:
                self.sm.append(float(nv.nvmlDeviceGetClockInfo(h, nv.NVML_CLOCK_SM)))
                self.mx.append(self.max_sm)
                r = int(nv.nvmlDeviceGetCurrentClocksEventReasons(h))
                for bit, nm in names.items():
                    if r & bit:
                        self.reasons.add(nm)
                try:
                    self.pw.append(nv.nvmlDeviceGetPowerUsage(h) / 1000.0)
                except Exception:
                    pass
            except Exception:
                break
            time.sleep(0.002)

    def _poll_smi(self):
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for ln in self.proc.stdout:
            p = [s.strip() for s in ln.split(",")]
            if len(p) < 9:
                continue
            try:
                self.sm.append(float(p[1])); self.mx.append(float(p[2])); self.pw.append(float(p[3]))
            except ValueError:
                continue
            for nm, v in zip(names, p[5:9]):
                if v.lower().startswith("active"):
                    self.reasons.add(nm)

    def start(self):
        if self.nv is not None:
            self.th = threading.Thread(target=self._poll_nvml, daemon=True)
            self.th.start()
            return
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "20", "-i", str(self.index)],
                stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.how = "nvidia-smi -lms 20"
            self.th = threading.Thread(target=self._poll_smi, daemon=True)
            self.th.start()
        except Exception:
            self.how = None

    def n(self) -> int:
        return len(self.sm)

    def stop(self) -> dict:
        if self.how is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["no NVML and no nvidia-smi"], "samples": 0}
        self.stop_flag.set()
        if self.nv is None:
            time.sleep(0.05)
            self.proc.terminate()
            try:
                self.proc.wait(timeout=2)
            except Exception:
                self.proc.kill()
        if self.th:
            self.th.join(timeout=2)
        return {"sm_mhz": float(np.median(self.sm)) if self.sm else None,
                "sm_max_mhz": max(self.mx) if self.mx else None,
                "power_w_max": max(self.pw) if self.pw else None,
                "reasons": sorted(self.reasons), "samples": len(self.sm), "how": self.how}


# ------------------------------------------------------------------------------------------------------
# CPU reference arm / baseline
# ------------------------------------------------------------------------------------------------------
def cpu_model() -> str:
    """host CPU model string (the CPU baseline states core count and model of the box it ran on)"""
    try:
        with open("/proc/cpuinfo") as f:
            for ln in f:
                if ln.lower().startswith("model name"):
                    return ln.split(":", 1)[1].strip()
    except Exception:
        pass
    return "unknown"


def cpu_chain_rate(seconds_target: float, threads: int, frames_per_call: int = 64, full: bool = False):
    """Time the CPU chain (oracle FIR restatement + vendored kissfft at the reference's build flags) on a
    bounded sample with `threads` host threads.  Returns (Msamples/s, description, kind, frames, secs)."""
    import oracle
    from libredio_b200 import synth
    from concurrent.futures import ThreadPoolExecutor
    taps = synth.lpf_taps(NTAPS, 0.04)
    win = synth.hann_periodic(NFFT)
    x = synth.cf32_noise_tones(n_input(frames_per_call), seed=2)
    use_ref = oracle.have_ref()

    def one(_):
        psd, nfr = oracle.chain_psd_cpu(x, taps, DECIM, NFFT, win, use_ref=use_ref, opt=False, full=full)
        return nfr

    t0 = time.perf_counter()
    one(0)
    per_call = time.perf_counter() - t0
    calls = max(threads, int(seconds_target / max(per_call, 1e-6)) * threads)
    calls = (calls + threads - 1) // threads * threads
    t0 = time.perf_counter()
    with ThreadPoolExecutor(max_workers=threads) as ex:
        frames = sum(ex.map(one, range(calls)))
    dt = time.perf_counter() - t0
    msps = frames * NFFT * DECIM / dt / 1e6
    kind = "port"
    desc = (f"{frames} frames ({frames * NFFT * DECIM} input samples) of the same workload, {threads} threads: "
            f"strict-f32 C restatement of dsputils::convolve "
            + ("(every lag computed as in dsputils.rs:30-32, then every 10th kept) + " if full
               else "(decimating: only kept outputs are computed) + ")
            + ("vendored kiss_fft.c built with the reference's flags (libkissfft/Makefile:4, no -O)" if use_ref
               else "restated kissfft (oracle/_ref not present)") + " + Hann + |X|^2 accumulation")
    return msps, desc, kind, frames, dt


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    cores = os.cpu_count() or 1
    vals = []
    for _ in range(args.warmup):
        cpu_chain_rate(0.5, cores)
    per_step = max(2.0, min(20.0, 90.0 / max(args.steps, 1)))
    t_all = 0.0
    frames_all = 0
    desc = kind = ""
    for _ in range(args.steps):
        msps, desc, kind, frames, dt = cpu_chain_rate(per_step, cores)
        vals.append(msps); t_all += dt; frames_all += frames
    value = frames_all * NFFT * DECIM / t_all / 1e6
    f_msps, f_desc, _, _, _ = cpu_chain_rate(1.0, cores, full=True)
    out = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * t_all / max(args.steps, 1),
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": workload_name(FRAMES), "sample_per_step": desc},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "cpu_model": cpu_model(), "kind": kind, "sample": desc,
                         "as_written_all_lags": {"value": f_msps, "unit": UNIT, "sample": f_desc,
                                                 "note": "informational: the reference has no decimating FIR; `value` is the "
                                                         "stronger CPU baseline that skips the 9 of 10 dropped outputs"}},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(out), flush=True)


# ------------------------------------------------------------------------------------------------------
# our arm
# ------------------------------------------------------------------------------------------------------
def run_ours(args):
    import torch
    import torch.distributed as dist
    from libredio_b200 import blocks, synth, capi

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; the product has no CPU fallback (use --impl reference for the CPU arm)")
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    ctx = blocks.Context(local)
    dev = ctx.tdev
    frames = args.frames
    n_in = n_input(frames)
    taps = synth.lpf_taps(NTAPS, 0.04)
    chain = blocks.Chain(ctx, taps, DECIM, NFFT, capi.WINDOW_HANN)

    # synthetic input, generated on the device for the HBM-resident leg: noise + 3 tones, seeded per rank
    g = torch.Generator(device=dev).manual_seed(2 + rank)
    x = torch.view_as_complex(torch.randn(n_in, 2, device=dev, generator=g))
    CH = 1 << 24
    for f, a in ((0.011, 2.0), (-0.0273, 1.0), (0.0402, 0.5)):
        for s in range(0, n_in, CH):
            e = min(n_in, s + CH)
            ph = (2 * np.pi * f) * torch.arange(s, e, device=dev, dtype=torch.float64)
            x[s:e] += torch.polar(torch.full((e - s,), a, device=dev, dtype=torch.float32), (ph % (2 * np.pi)).float())
    rows = frames // K_AVG
    # two output slots: the NVLink gather of step i (NCCL's stream) overlaps the kernel of step i+1
    outs = [torch.empty((rows, NFFT), dtype=torch.float32, device=dev) for _ in range(2)]
    pending = [None, None]
    # the only exchange: every rank's output rows go to every peer.  Default: lrc_gather (copy engines over NVLink,
    # zero SMs, overlaps the next step's persistent kernel); fallback / --gather nccl: NCCL all-gather, whose kernel
    # cannot start while the chain kernel fills every SM and therefore serialises with it.
    gather, gather_kind, gath = None, "none (1 GPU)", None
    if world > 1:
        ok = torch.zeros(1, device=dev)
        if args.gather == "ce":
            try:
                gather = blocks.Gather(ctx, rank, world, rows * NFFT * 4, slots=2).connect_distributed()
                ok += 1
            except Exception as e:                         # e.g. CUDA IPC not permitted in this container
                print(f"bench.py rank {rank}: lrc_gather unavailable ({e}); using the NCCL all-gather", file=sys.stderr)
                gather = None
        # bounded-time probe before the timed loop depends on it: one push + arrival wait per slot on a side stream,
        # polled from the host.  A peer whose flag write never arrives would otherwise hang the bench inside a
        # device-side wait.  The probing stream can never drain in that case, so the process cannot fall back: it
        # fails fast with a message instead (rerun with --gather nccl).
        if gather is not None:
            side = torch.cuda.Stream(device=dev)
            with torch.cuda.stream(side):
                for b in range(2):
                    gather.push(b, outs[b])
                    gather.wait(b)
                pev = torch.cuda.Event()
                pev.record()
            t0, probe_stuck = time.perf_counter(), False
            while not pev.query():
                if time.perf_counter() - t0 > args.gather_probe_s:
                    probe_stuck = True
                    break
                time.sleep(0.005)
            if probe_stuck:
                print(f"bench.py rank {rank}: lrc_gather probe did not complete in {args.gather_probe_s:.0f} s "
                      "(a peer's arrival flag never came); rerun with --gather nccl", file=sys.stderr, flush=True)
                os._exit(3)
        dist.all_reduce(ok)                                # all ranks must agree on the mechanism
        if int(ok.item()) != world:
            if gather is not None:
                gather.close()
            gather = None
        if gather is None:
            gath = [torch.empty((world * rows, NFFT), dtype=torch.float32, device=dev) for _ in range(2)]
            gather_kind = "NCCL all_gather_into_tensor of output rows (async, 2 slots)"
        else:
            gather_kind = "lrc_gather: copy-engine P2P push of output rows into every peer's slot (CUDA IPC, 2 slots)"

    def step(i):
        b = i & 1
        if gather is not None:
            gather.wait_sent(b)                            # slot reuse: the previous push must have read outs[b]
        elif pending[b] is not None:
            pending[b].wait()                              # slot reuse: its previous gather must be done
            pending[b] = None
        chain.run(x, K_AVG, outs[b])
        if gather is not None:
            gather.push(b, outs[b])
        elif world > 1:
            pending[b] = dist.all_gather_into_tensor(gath[b], outs[b], async_op=True)

    def drain():
        for b in range(2):
            if gather is not None:
                gather.wait(b)                             # every peer's rows for this slot have ARRIVED here
            elif pending[b] is not None:
                pending[b].wait()
                pending[b] = None

    for i in range(max(args.warmup, 3)):
        step(i)
    drain()
    torch.cuda.synchronize()

    try:
        uuid = str(torch.cuda.get_device_properties(local).uuid)
    except Exception:
        uuid = None
    sampler = ClockSampler(local, uuid) if rank == 0 else None
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    if sampler:
        sampler.start()
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(args.steps + 1)]
    kev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.steps)]
    ev[0].record()
    for i in range(args.steps):
        b = i & 1
        if gather is not None:
            gather.wait_sent(b)
        elif pending[b] is not None:
            pending[b].wait()
            pending[b] = None
        kev[i][0].record()
        chain.run(x, K_AVG, outs[b])
        kev[i][1].record()
        if gather is not None:
            gather.push(b, outs[b])
        elif world > 1:
            pending[b] = dist.all_gather_into_tensor(gath[b], outs[b], async_op=True)
        if i == args.steps - 1:
            drain()                                        # the last gathers are inside the timed region
        ev[i + 1].record()
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    if gather is not None:
        # the gathered rows must be what an NCCL all-gather of the same outputs gives (every step computes the same
        # rows from the same resident input, so both slots hold the final result of every rank)
        torch.cuda.synchronize()
        for b in range(2):
            want = torch.empty((world * rows, NFFT), dtype=torch.float32, device=dev)
            dist.all_gather_into_tensor(want, outs[b])
            got = gather.buffer(b).reshape(world * rows, NFFT)
            if not torch.equal(got, want):
                raise SystemExit(f"bench.py rank {rank}: lrc_gather slot {b} differs from the NCCL all-gather")
    total_ms = ev[0].elapsed_time(ev[-1])
    kern_ms = float(np.mean([a.elapsed_time(b) for a, b in kev]))
    clocks = None
    if sampler:
        # NVML refreshes its clock reading only every few tens of ms: when the timed region was too short
        # to be seen, keep the same steps running (untimed) under the sampler until it has a few readings.
        window = "timed region"
        if sampler.n() < 8 or total_ms < 300.0:
            t_end = time.perf_counter() + 0.6
            i = 0
            while time.perf_counter() < t_end:
                chain.run(x, K_AVG, outs[i & 1]); i += 1
                if i % 16 == 0:
                    torch.cuda.synchronize()
            torch.cuda.synchronize()
            window = f"timed region ({total_ms:.0f} ms) + 0.6 s of the same step repeated untimed right after it"
        clocks = sampler.stop()
        clocks["window"] = window
    t = torch.tensor([total_ms], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    total_ms = float(t.item())
    ms_per_step = total_ms / args.steps
    value = world * n_in / (ms_per_step * 1e-3) / 1e6

    # ---- e2e through the host-buffer entry point (rank-local pinned input) ---------------------------
    e2e_steps = max(1, min(args.steps, 3))
    e2e_frames = min(frames, args.e2e_frames)
    n_e = n_input(e2e_frames)
    xh = torch.empty(n_e, dtype=torch.complex64, pin_memory=True)
    xh.copy_(x[:n_e])
    rows_h = torch.empty((e2e_frames // K_AVG, NFFT), dtype=torch.float32, pin_memory=True)
    chain.run_host(xh, K_AVG, rows_h)                      # warm-up (allocates the ring)
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(e2e_steps):
        chain.run_host(xh, K_AVG, rows_h)                  # synchronous: returns when rows are on the host
    torch.cuda.synchronize()
    e2e_s = time.perf_counter() - t0
    te = torch.tensor([e2e_s], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(te, op=dist.ReduceOp.MAX)
    e2e_value = world * n_e * e2e_steps / float(te.item()) / 1e6
    # the same chain fed in the reference's wire format (rtlsdr u8 I,Q: 2 bytes per sample over PCIe, unpacked on
    # the device): informational, NOT the headline e2e (which keeps cf32 host buffers like the metric says)
    iqh = torch.empty(2 * n_e, dtype=torch.uint8, pin_memory=True)
    iqh.copy_((x[:n_e].view(torch.float32).reshape(-1).clamp(-1, 1) * 127 + 127).round().to(torch.uint8))
    chain.run_host_u8(iqh, K_AVG, rows_h)
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    t0 = time.perf_counter()
    for _ in range(e2e_steps):
        chain.run_host_u8(iqh, K_AVG, rows_h)
    torch.cuda.synchronize()
    tu = torch.tensor([time.perf_counter() - t0], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(tu, op=dist.ReduceOp.MAX)
    e2e_u8_value = world * n_e * e2e_steps / float(tu.item()) / 1e6
    del iqh
    # what bounds e2e: a bare pinned-host -> device copy of the same buffer on this box's PCIe link
    xd = torch.empty(n_e, dtype=torch.complex64, device=dev)
    xd.copy_(xh, non_blocking=True)
    torch.cuda.synchronize()
    c0, c1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    c0.record()
    for _ in range(3):
        xd.copy_(xh, non_blocking=True)
    c1.record()
    torch.cuda.synchronize()
    h2d_gbs = 3 * n_e * 8 / (c0.elapsed_time(c1) * 1e-3) / 1e9
    del xd

    if rank == 0:
        peaks_path = os.path.join(ROOT, "MEASURED_PEAKS.json")
        if os.path.exists(peaks_path):
            peak = float(json.load(open(peaks_path))["hbm_gbs"]); peak_src = "measured (MEASURED_PEAKS.json hbm_gbs)"
        else:
            peak = 6650.0; peak_src = "fallback (B200_PROFILING.md 6.65 TB/s)"
        alg_bytes = n_in * BYTES_PER_SAMPLE + rows * NFFT * 4
        achieved = alg_bytes / (kern_ms * 1e-3) / 1e9
        traffic = None
        tp = os.path.join(ROOT, "profiles", "chain_traffic.json")
        if os.path.exists(tp):
            try:
                traffic = json.load(open(tp)).get("dram_bytes_per_launch")
            except Exception:
                traffic = None
        cpu = None
        if world == 1 and not args.no_cpu:
            msps, desc, kind, _, _ = cpu_chain_rate(args.cpu_seconds, os.cpu_count() or 1)
            cpu = {"value": msps, "unit": UNIT, "cores": os.cpu_count() or 1, "cpu_model": cpu_model(), "kind": kind,
                   "sample": desc}
            f_msps, f_desc, _, _, _ = cpu_chain_rate(0.4, os.cpu_count() or 1, full=True)
            cpu["as_written_all_lags"] = {"value": f_msps, "unit": UNIT, "sample": f_desc,
                                          "note": "informational: the reference has no decimating FIR; `value` is the "
                                                  "stronger CPU baseline that skips the 9 of 10 dropped outputs"}
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": max(args.warmup, 3), "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": workload_name(frames), "ntaps": NTAPS, "decim": DECIM, "nfft": NFFT, "k_avg": K_AVG,
                       "window": "hann", "l2": "input 5.4 GB per step >> 126 MB L2, no flush needed",
                       "sharding": f"{world} independent streams, one per GPU; gather of output rows only",
                       "gather": gather_kind,
                       "e2e_workload": workload_name(e2e_frames) + f", {e2e_steps} steps through lrc_chain_run_host"},
            "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                         "traffic": traffic, "peak_source": peak_src, "kernel": "chain_kernel<64,10,10,7> (+ psd_reduce)",
                         "kernel_ms": kern_ms, "algorithmic_bytes_per_launch": alg_bytes},
            "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": n_e * 8,
                    "d2h_bytes_per_step": (e2e_frames // K_AVG) * NFFT * 4,
                    "bound": "pcie h2d", "h2d_copy_gbs_measured": h2d_gbs,
                    "frac_of_h2d_copy": (e2e_value / world) * 8e6 / (h2d_gbs * 1e9)},
            "e2e_u8_wire_format": {"value": e2e_u8_value, "unit": UNIT, "h2d_bytes_per_step": n_e * 2,
                                   "note": "same chain through lrc_chain_run_host_u8: host buffers hold rtlsdr u8 I,Q, "
                                           "data_to_samples runs on the device; informational"},
            "gpu_launches": 2 * args.steps,
            "clocks": clocks,
        }
        if cpu:
            line["cpu_baseline"] = cpu
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.barrier()
    if gather is not None:
        gather.close()
    chain.close()
    ctx.close()
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--frames", type=int, default=FRAMES, help="frames per step (default: BASELINE config)")
    ap.add_argument("--e2e-frames", type=int, default=16384, dest="e2e_frames",
                    help="frames per e2e step through host buffers (1.3 GB pinned by default)")
    ap.add_argument("--cpu-seconds", type=float, default=1.5, dest="cpu_seconds",
                    help="per-thread seconds of CPU work for the cpu_baseline sample")
    ap.add_argument("--no-cpu", action="store_true", dest="no_cpu")
    ap.add_argument("--gather-probe-s", type=float, default=30.0, dest="gather_probe_s",
                    help="N > 1: seconds the lrc_gather connectivity probe may take before falling back to NCCL")
    ap.add_argument("--gather", default="ce", choices=["ce", "nccl"],
                    help="N > 1: output gather by lrc_gather (copy engines over NVLink, default) or NCCL all-gather")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
