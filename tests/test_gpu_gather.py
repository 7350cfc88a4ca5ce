"""GPU suite: the copy-engine output gather (include/libredio_cuda.h "(e)", libredio_b200/csrc/k_gather.cu).

A single-GPU box can still exercise both connection modes: two gather objects of one process on the same device
(lrc_gather_connect_local), and two PROCESSES on cuda:0 mapping each other's receive buffers through CUDA IPC
(lrc_gather_connect) -- the mode bench.py uses with one process per GPU.  Bit-exact: the gather only moves bytes."""
import os
import subprocess
import sys

import numpy as np
import pytest
import torch

from libredio_b200 import blocks

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_gather_world_1_is_a_local_copy(ctx):
    n = 1000
    g = blocks.Gather(ctx, 0, 1, n * 4, slots=2)
    src = torch.arange(n, dtype=torch.float32, device=ctx.tdev)
    for it in range(3):
        slot = it & 1
        g.wait_sent(slot)
        src.add_(1.0)
        g.push(slot, src)
        g.wait(slot)
        assert torch.equal(g.buffer(slot)[0], src)
    ctx.sync()
    g.close()


@pytest.mark.parametrize("nbytes", [4, 4096, 4 * 1024 * 1024 + 12])
def test_gather_two_local_ranks_same_device(ctx, nbytes):
    world, slots = 2, 2
    gs = [blocks.Gather(ctx, r, world, nbytes, slots) for r in range(world)]
    blocks.Gather.connect_local(gs)
    n = nbytes // 4
    srcs = [[torch.empty(n, dtype=torch.float32, device=ctx.tdev) for _ in range(slots)] for _ in range(world)]
    streams = [torch.cuda.Stream() for _ in range(world)]
    for it in range(5):
        slot = it % slots
        for r in range(world):
            with torch.cuda.stream(streams[r]):
                gs[r].wait_sent(slot)
                srcs[r][slot].fill_(float(100 * it + r))
                gs[r].push(slot, srcs[r][slot])
        for r in range(world):
            with torch.cuda.stream(streams[r]):
                gs[r].wait(slot)
                buf = gs[r].buffer(slot)
                got = buf.clone()
            streams[r].synchronize()
            for p in range(world):
                assert torch.all(got[p] == float(100 * it + p)), (it, r, p)
    ctx.sync()
    for g in gs:
        g.close()


def test_gather_push_before_connect_is_an_error(ctx):
    g = blocks.Gather(ctx, 0, 2, 64, 1)
    with pytest.raises(blocks.capi.LrcError):
        g.push(0, torch.zeros(16, dtype=torch.float32, device=ctx.tdev))
    with pytest.raises(blocks.capi.LrcError):
        blocks.Gather(ctx, 2, 2, 64, 1)
    g.close()


_WORKER = r'''
import os, sys
sys.path.insert(0, {root!r})
import numpy as np, torch, torch.distributed as dist
from libredio_b200 import blocks
rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
dist.init_process_group("gloo", rank=rank, world_size=world)
dev = rank % torch.cuda.device_count()
ctx = blocks.Context(dev)
rows, nfft, slots = 64, 1024, 2
g = blocks.Gather(ctx, rank, world, rows * nfft * 4, slots).connect_distributed()
outs = [torch.empty((rows, nfft), dtype=torch.float32, device=ctx.tdev) for _ in range(slots)]
gen = torch.Generator(device=ctx.tdev)
for it in range(6):
    slot = it % slots
    g.wait_sent(slot)
    gen.manual_seed(1000 * it + rank)
    outs[slot].copy_(torch.randn((rows, nfft), device=ctx.tdev, generator=gen))
    g.push(slot, outs[slot])
    g.wait(slot)
    got = g.buffer(slot).clone().reshape(world, rows, nfft)
    torch.cuda.synchronize()
    for p in range(world):
        pdev = p % torch.cuda.device_count()
        gp = torch.Generator(device=torch.device("cuda", pdev)); gp.manual_seed(1000 * it + p)
        want = torch.randn((rows, nfft), device=torch.device("cuda", pdev), generator=gp).to(ctx.tdev)
        assert torch.equal(got[p], want), (it, rank, p)
    dist.barrier()                       # flow control of this test: nobody re-pushes a slot a peer still reads
torch.cuda.synchronize()
dist.barrier()
g.close(); ctx.close()
dist.destroy_process_group()
print("rank", rank, "ok")
'''


def test_gather_two_processes_cuda_ipc(tmp_path):
    script = tmp_path / "gather_worker.py"
    script.write_text(_WORKER.format(root=ROOT))
    env = dict(os.environ, MASTER_ADDR="127.0.0.1", MASTER_PORT="29641", WORLD_SIZE="2", OMP_NUM_THREADS="1")
    procs = [subprocess.Popen([sys.executable, str(script)], env=dict(env, RANK=str(r)), stdout=subprocess.PIPE,
                              stderr=subprocess.STDOUT, text=True) for r in range(2)]
    outs = [p.communicate(timeout=300)[0] for p in procs]
    for r, (p, o) in enumerate(zip(procs, outs)):
        assert p.returncode == 0, f"rank {r} failed:\n{o}"
        assert f"rank {r} ok" in o


def test_gather_to_one_root_only_that_rank_receives(ctx):
    """lrc_gather_set_root: a gather instead of an all-gather -- every rank pushes ONE block (to the root), the root's
    buffer holds every rank's block, nothing lands on the other ranks and their wait() has nothing to wait for."""
    world, slots, n = 3, 2, 2048
    gs = [blocks.Gather(ctx, r, world, n * 4, slots) for r in range(world)]
    blocks.Gather.connect_local(gs)
    for g in gs:
        g.set_root(1)
    srcs = [torch.empty(n, dtype=torch.float32, device=ctx.tdev) for _ in range(world)]
    for it in range(4):
        slot = it % slots
        for r in range(world):
            gs[r].wait_sent(slot)
            srcs[r].fill_(float(10 * it + r + 1))
            gs[r].push(slot, srcs[r])
        for r in range(world):
            gs[r].wait(slot)
        ctx.sync()
        torch.cuda.synchronize()
        root_buf = gs[1].buffer(slot)
        for p in range(world):
            assert torch.all(root_buf[p] == float(10 * it + p + 1)), (it, p)
        for r in (0, 2):
            assert torch.all(gs[r].buffer(slot) == 0), (it, r)           # never written: still the zeros of create
    with pytest.raises(blocks.capi.LrcError):
        gs[0].set_root(0)                                                # pushes were already issued
    for g in gs:
        g.close()


def _host_gather_worker(rank, world, name, nbytes, q):
    """one process per rank, all on cuda:0: the host gather needs no peer access, only the shared-memory segment"""
    import numpy as np
    import torch
    from libredio_b200 import blocks
    ctx = blocks.Context(0)
    g = blocks.Gather(ctx, rank, world, nbytes, slots=2, host_shm=name, root=0)
    n = nbytes // 4
    for step in range(5):
        b = step & 1
        src = torch.full((n,), float(100 * step + rank), dtype=torch.float32, device=ctx.tdev)
        g.wait_sent(b)
        g.push(b, src)
    if rank == 0:
        g.wait(0)                                        # slot 0: order the current stream behind every rank's latest push ...
        torch.cuda.synchronize()
        g.wait_host(1, 20000)                            # ... slot 1: a CPU consumer polls the flag words in the shared segment
        last = {0: 4, 1: 3}                              # last step that pushed each slot
        ok = all(bool((g.buffer(b)[r] == float(100 * last[b] + r)).all()) for b in (0, 1) for r in range(world))
        q.put(ok)
    else:
        torch.cuda.synchronize()
        q.put(True)
    import time
    time.sleep(0.5 if rank == 0 else 0.0)                # the root owns the segment: let the others finish first
    g.close(); ctx.close()


@pytest.mark.gpu
def test_gather_to_host_shared_memory_two_processes():
    """lrc_gather_create_host: every rank copies its block D2H into one page-locked POSIX shared-memory segment and raises its
    flag there; the root orders a stream behind the flags and reads all blocks from host memory"""
    import multiprocessing as mp
    import os
    mpc = mp.get_context("spawn")
    q = mpc.Queue()
    name = f"/lrc_test_gather_{os.getpid()}"
    world, nbytes = 3, 1 << 20
    ps = [mpc.Process(target=_host_gather_worker, args=(r, world, name, nbytes, q)) for r in range(world)]
    for p in ps:
        p.start()
    res = [q.get(timeout=120) for _ in ps]
    for p in ps:
        p.join(timeout=60)
    assert all(res) and all(p.exitcode == 0 for p in ps)
