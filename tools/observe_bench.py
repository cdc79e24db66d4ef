#!/usr/bin/env python
"""Streaming kernels: observe / board_contents / reset throughput vs the HBM write roofline."""
import sys
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
import torch
from blokus_rl_b200 import BlokusEngine
eng = BlokusEngine(20, 4)
n = 65536
s = eng.new_states(n)
o = eng.step(s, None, mask=None, sample=True, seed=3)
for _ in range(20):
    o = eng.step(s, o.next_action, mask=None, sample=True, seed=3)
obs = torch.empty((n, 8, 20, 20), dtype=torch.float32, device="cuda")
def timed(fn, reps=20):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps
ms = timed(lambda: eng.observe(s, out=obs)); print(f"observe {n}: {ms:.3f} ms  {n * 12800 / ms / 1e6:.0f} GB/s written  {n / ms * 1e3:.3e} obs/s")
ms = timed(lambda: eng.board_contents(s)); print(f"board_contents {n}: {ms:.3f} ms  {n * 400 / ms / 1e6:.0f} GB/s")
ms = timed(lambda: eng.reset(s)); print(f"reset {n}: {ms:.3f} ms  {n * 352 / ms / 1e6:.0f} GB/s")
ms = timed(lambda: eng.game_ended(s)); print(f"game_ended {n}: {ms:.3f} ms")
