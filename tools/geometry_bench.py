#!/usr/bin/env python
"""Random-play step throughput for every board size / player count with its own kernels: thread-per-env kernels
for N <= 7 (csrc/blk_small.cu), warp-per-env specialisations for 14x14 and 20x20."""
import sys
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
import torch
from blokus_rl_b200 import BlokusEngine
for (N, P, n) in ((5, 2, 1048576), (6, 2, 1048576), (7, 2, 1048576), (7, 4, 1048576), (14, 2, 131072), (14, 4, 131072),
                  (20, 2, 65536), (20, 4, 65536)):
    eng = BlokusEngine(N, P)
    for fmt in ("bytes", "bits"):
        s = eng.new_states(n)
        buf = eng.make_buffers(n, fmt, sample=True)
        eng.step(s, None, buffers=buf, mask=fmt, sample=True, seed=1)
        for _ in range(5):
            eng.step(s, buf.next_action, buffers=buf, mask=fmt, sample=True, seed=1, auto_reset=True)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(50):
            eng.step(s, buf.next_action, buffers=buf, mask=fmt, sample=True, seed=1, auto_reset=True)
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / 50
        bytes_per = (eng.mask_bytes if fmt == "bytes" else eng.mask_words * 4) + 2 * eng.state_words * 4 + 17
        print(f"{N}x{N} {P}p {fmt:5s} A={eng.num_actions:6d} n={n}: {ms:.3f} ms  {n / ms * 1e3:.3e} steps/s  {bytes_per * n / ms / 1e6:.0f} GB/s")
    eng.close()
