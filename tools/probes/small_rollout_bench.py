#!/usr/bin/env python
"""7x7 two-player playouts: thread-per-playout kernel vs the warp-per-playout kernel."""
import sys
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parents[2]))
import torch
from blokus_rl_b200 import BlokusEngine
eng = BlokusEngine(7, 2)
roots = eng.new_states(1024)
o = eng.step(roots, None, mask=None, sample=True, seed=2)
for _ in range(2):
    o = eng.step(roots, o.next_action, mask=None, sample=True, seed=2)
for warp in (False, True):
    for _ in range(2):
        r = eng.rollout(roots, 1024, seed=7, warp_kernels=warp)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(5):
        r = eng.rollout(roots, 1024, seed=7, warp_kernels=warp)
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 5
    print(f"{'warp-per-playout' if warp else 'thread-per-playout'}: {1024 * 1024 / ms * 1e3:.3e} playouts/s, "
          f"{1024 * 1024 * float(r.plies.float().mean()) / ms * 1e3:.3e} plies/s ({float(r.plies.float().mean()):.1f} plies per playout)")
