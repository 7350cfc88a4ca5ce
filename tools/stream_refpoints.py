#!/usr/bin/env python
"""Reference points for the HBM-bound kernels: what PyTorch's own elementwise kernels reach on this GPU with the
same read:write mix (so that a fraction of the 1:1 copy peak can be read against the mix's own practical ceiling).
Not a product path; prints one JSON line per mix."""
import json
import torch

def timeit(fn, iters=10, warm=3):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(iters + 1)]
    ev[0].record()
    for i in range(iters):
        fn(); ev[i + 1].record()
    torch.cuda.synchronize()
    return sorted(ev[i].elapsed_time(ev[i + 1]) for i in range(iters))[iters // 2]

dev = torch.device("cuda", 0)
n = 1 << 30
out = []
x8 = torch.randint(0, 256, (n,), dtype=torch.uint8, device=dev)
yf = torch.empty(n, dtype=torch.float32, device=dev)
ms = timeit(lambda: yf.copy_(x8)); out.append(("u8 -> f32 convert (1 B read : 4 B written, the unpack mix)", n * 5, ms))
ms = timeit(lambda: yf.zero_()); out.append(("memset f32 (write only)", n * 4, ms))
xf = torch.empty(n, dtype=torch.float32, device=dev).normal_()
ms = timeit(lambda: yf.copy_(xf)); out.append(("f32 copy (1:1, the mix MEASURED_PEAKS.json uses)", n * 8, ms))
ms = timeit(lambda: torch.sum(xf)); out.append(("f32 sum (read only)", n * 4, ms))
del x8
xc = torch.view_as_complex(torch.randn(n // 2, 2, device=dev))
ya = torch.empty(n // 2, dtype=torch.float32, device=dev)
ms = timeit(lambda: torch.abs(xc, out=ya)); out.append(("complex64 -> f32 abs (8 B read : 4 B written, the discriminator mix)", n // 2 * 12, ms))
for name, b, ms in out:
    print(json.dumps({"mix": name, "GB/s": b / (ms * 1e-3) / 1e9, "ms": ms, "bytes": b}), flush=True)
