#!/usr/bin/env python
"""Step throughput into a CONTIGUOUS bool [n, 30433] mask (rows not 16 B aligned) vs the padded layout."""
import sys
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
import torch
from blokus_rl_b200 import BlokusEngine
eng = BlokusEngine(20, 4)
n = 65536
s = eng.new_states(n)
buf = eng.make_buffers(n, None, sample=True)
for name, mask in (("padded [n, 30464] view", eng.alloc_mask(n, "bytes")),
                   ("contiguous [n, 30433]", torch.empty((n, eng.num_actions), dtype=torch.uint8, device="cuda"))):
    eng.reset(s)
    eng.step(s, None, buffers=buf, mask=mask, sample=True, seed=1)
    for _ in range(10):
        eng.step(s, buf.next_action, buffers=buf, mask=mask, sample=True, seed=1, auto_reset=True)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(200):
        eng.step(s, buf.next_action, buffers=buf, mask=mask, sample=True, seed=1, auto_reset=True)
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 200
    print(f"{name}: {ms:.4f} ms per step, {n / ms * 1e3:.3e} steps/s")
