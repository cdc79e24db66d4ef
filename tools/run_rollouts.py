#!/usr/bin/env python
"""Small driver for profiling the rollout kernel (BASELINE configs[2] shape, reduced): python tools/run_rollouts.py [roots] [per_root]"""
import sys
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
import torch
from blokus_rl_b200 import BlokusEngine

roots_n = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
per = int(sys.argv[2]) if len(sys.argv) > 2 else 256
eng = BlokusEngine(20, 4)
roots = eng.new_states(roots_n)
out = eng.step(roots, None, mask=None, sample=True, seed=24)
for _ in range(24):
    out = eng.step(roots, out.next_action, mask=None, sample=True, seed=24)
for rep in range(3):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    r = eng.rollout(roots, per, seed=7)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1)
    plies = r.plies.float().mean().item()
    print(f"rollouts {roots_n * per} in {ms:.2f} ms -> {roots_n * per / ms * 1e3:.3e} rollouts/s, {roots_n * per * plies / ms * 1e3:.3e} plies/s (mean {plies:.1f} plies)")
