"""Host <-> device copy bandwidth of the box with pinned buffers (what bounds the upload / download legs of bench.py's e2e).
    python tools/pcie_bw.py"""
import time

import torch

for mb in (8, 32, 128):
    n = mb << 20
    h = torch.empty(n, dtype=torch.uint8, pin_memory=True)
    d = torch.empty(n, dtype=torch.uint8, device="cuda")
    for name, fn in (("h2d", lambda: d.copy_(h, non_blocking=True)), ("d2h", lambda: h.copy_(d, non_blocking=True))):
        fn(); torch.cuda.synchronize()
        t0 = time.perf_counter()
        for _ in range(10):
            fn()
        torch.cuda.synchronize()
        dt = (time.perf_counter() - t0) / 10
        print(f"{name} {mb:4d} MiB: {n / dt / 1e9:6.1f} GB/s ({dt * 1e3:.2f} ms)")
