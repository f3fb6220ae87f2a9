"""Raw pinned-memory copy rates of the box (context for bench.py's e2e numbers)."""
import time

import torch

dev = torch.device("cuda", 0)
for mb in (1, 4, 16, 64):
    n = mb << 20
    h = torch.empty(n, dtype=torch.uint8).pin_memory()
    d = torch.empty(n, dtype=torch.uint8, device=dev)
    for name, (src, dst) in {"h2d": (h, d), "d2h": (d, h)}.items():
        for _ in range(3):
            dst.copy_(src, non_blocking=True)
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        reps = 50
        for _ in range(reps):
            dst.copy_(src, non_blocking=True)
        torch.cuda.synchronize()
        dt = time.perf_counter() - t0
        print(f"{name} {mb} MiB: {reps * n / dt / 1e9:.1f} GB/s, {1e6 * dt / reps:.1f} us/copy")
