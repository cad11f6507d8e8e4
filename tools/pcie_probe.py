#!/usr/bin/env python3
"""Host -> device bandwidth from page-locked memory for a few copy sizes, on one stream and
alternating two streams (what bounds bench.py's e2e)."""
import time
import torch

x = torch.empty(1 << 28, dtype=torch.float32).pin_memory()   # 1 GiB
d = torch.empty_like(x, device="cuda")
streams = [torch.cuda.Stream(), torch.cuda.Stream()]
for ns in (1, 2):
    for n in (8 << 20, 64 << 20, 1 << 30):
        k = n // 4
        reps = max(2, (4 << 30) // n)
        for timed in (False, True):
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            for i in range(reps):
                off = (i * k) % (x.numel() - k + 1) if n < (1 << 30) else 0
                with torch.cuda.stream(streams[i % ns]):
                    d[off:off + k].copy_(x[off:off + k], non_blocking=True)
            torch.cuda.synchronize()
            el = time.perf_counter() - t0
        print("%d stream(s), %4d MiB copies: %.1f GB/s" % (ns, n >> 20, reps * n / el / 1e9))
