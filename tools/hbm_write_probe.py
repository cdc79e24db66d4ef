#!/usr/bin/env python
"""Calibrate the write-only HBM ceiling on this GPU: torch fill_ and copy_ over 2 GiB (CUDA events, best of 10)."""
import torch
n = 2 * 1024 ** 3
a = torch.empty(n, dtype=torch.uint8, device="cuda")
b = torch.empty(n, dtype=torch.uint8, device="cuda")
def best(fn, reps=10):
    fn(); torch.cuda.synchronize()
    t = []
    for _ in range(reps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record(); torch.cuda.synchronize()
        t.append(e0.elapsed_time(e1))
    return min(t)
ms = best(lambda: a.fill_(1))
print(f"fill_ 2 GiB: {ms:.3f} ms -> {n / ms / 1e6:.0f} GB/s write-only")
ms = best(lambda: a.view(torch.int64).fill_(1))
print(f"fill_ int64 2 GiB: {ms:.3f} ms -> {n / ms / 1e6:.0f} GB/s write-only")
ms = best(lambda: b.copy_(a))
print(f"copy_ 2 GiB: {ms:.3f} ms -> {2 * n / ms / 1e6:.0f} GB/s read+write")
ms = best(lambda: torch.cuda.memset if False else a.zero_())
print(f"zero_ 2 GiB: {ms:.3f} ms -> {n / ms / 1e6:.0f} GB/s write-only")
