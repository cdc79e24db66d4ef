#!/usr/bin/env python
"""MCTS simulations/s of BatchedMCTS on the GPU engine (the reference: ~5.2 sims/s)."""
import sys, time
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
import torch
from blokus_rl_b200 import BlokusEngine
from blokus_rl_b200.mcts import BatchedMCTS, UniformEvaluator, RolloutEvaluator

eng = BlokusEngine(20, 4)
for B, sims, ev, name in ((1, 25, UniformEvaluator(), "uniform"), (256, 25, UniformEvaluator(), "uniform"),
                          (2048, 25, UniformEvaluator(), "uniform"), (256, 25, RolloutEvaluator(16), "rollout16")):
    s = eng.new_states(B)
    out = eng.step(s, None, mask=None, sample=True, seed=1)
    for _ in range(16):
        out = eng.step(s, out.next_action, mask=None, sample=True, seed=1)
    search = BatchedMCTS(eng, ev)
    roots = search.add_roots(s)
    search.simulate(roots)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(sims):
        search.simulate(roots)
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    print(f"B={B:5d} sims={sims} eval={name:9s}: {B * sims / dt:10.0f} sims/s  ({dt / sims * 1e3:.1f} ms per lockstep simulation, {search.launches} launches)")

# ---- device-resident forest (csrc/blk_puct.cu): no host work per simulation ----
from blokus_rl_b200.gpu_puct import GpuPuct
for B, sims, ev, name in ((256, 50, UniformEvaluator(), "uniform"), (4096, 50, UniformEvaluator(), "uniform"),
                          (16384, 25, UniformEvaluator(), "uniform"), (1024, 25, RolloutEvaluator(16), "rollout16")):
    s = eng.new_states(B)
    out = eng.step(s, None, mask=None, sample=True, seed=1)
    for _ in range(16):
        out = eng.step(s, out.next_action, mask=None, sample=True, seed=1)
    search = GpuPuct(eng, ev, num_trees=B, max_simulations=sims + 8, mean_edges_per_node=420)
    search.set_roots(s)
    for _ in range(4):                 # two eager runs, the CUDA-graph capture, one replay
        search.simulate()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(sims):
        search.simulate()
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    search.check()
    print(f"GPU forest B={B:5d} sims={sims} eval={name:9s}: {B * sims / dt:10.0f} sims/s  ({dt / sims * 1e3:.2f} ms per lockstep simulation)")
