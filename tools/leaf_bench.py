#!/usr/bin/env python
"""Leaf expansion micro-benchmark: step(mask only, bytes) + observe for small batches (launch-latency regime)."""
import sys
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
import torch
from blokus_rl_b200 import BlokusEngine
eng = BlokusEngine(20, 4)
for B in (64, 256, 1024, 4096, 16384):
    leaves = eng.new_states(B)
    o = eng.step(leaves, None, mask=None, sample=True, seed=3)
    for _ in range(20):
        o = eng.step(leaves, o.next_action, mask=None, sample=True, seed=3)
    buf = eng.make_buffers(B, "bytes")
    obs = torch.empty((B, 8, 20, 20), dtype=torch.float32, device="cuda")
    for name, fn in (("step", lambda: eng.step(leaves, None, buffers=buf, mask="bytes")),
                     ("observe", lambda: eng.observe(leaves, out=obs))):
        for _ in range(5):
            fn()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(200):
            fn()
        e1.record()
        torch.cuda.synchronize()
        print(f"B={B:6d} {name:8s} {e0.elapsed_time(e1) / 200 * 1e3:8.1f} us per call")
