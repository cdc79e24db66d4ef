#!/usr/bin/env python
"""How often does keying search nodes by path (the lockstep GpuPuct kernels) instead of by board cells (the reference,
alphazero/mcts.py:37; the host BatchedMCTS and the fused device search keep that) change a search result?  Plays whole games
with tree reuse on the device forest and on the host BatchedMCTS and counts the moves whose root visit counts differ.
   python tools/transposition_check.py [N] [P] [sims] [games] [--lockstep]
Default = the fused search (board-keyed: expect 0 differences); --lockstep = the path-keyed kernels (round 1: 1 of 133 at 7x7 / 200)."""
import sys
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
import numpy as np
import torch
from blokus_rl_b200 import BlokusEngine
from blokus_rl_b200.gpu_puct import GpuPuct
from blokus_rl_b200.mcts import BatchedMCTS, UniformEvaluator

lockstep = "--lockstep" in sys.argv
argv = [a for a in sys.argv[1:] if a != "--lockstep"]
N, P, sims, B = (int(x) for x in (argv[:4] + ["7", "2", "200", "16"][len(argv):]))
eng = BlokusEngine(N, P)
roots_t = eng.new_states(B)
out = eng.step(roots_t, None, mask=None, sample=True, seed=8)
eng.step(roots_t, out.next_action, mask=None)
gpu = GpuPuct(eng, UniformEvaluator(), num_trees=B, max_simulations=sims * 90, mean_edges_per_node=120 if N <= 7 else 400,
              fused=not lockstep)
gpu.set_roots(roots_t)
host = BatchedMCTS(eng, UniformEvaluator())
roots = host.add_roots(roots_t)
moves = differ = 0
for move in range(4 * 21):
    live = [t for t in range(B) if roots[t].terminal is None]
    if not live:
        break
    gpu.run(sims, 1.0)
    for _ in range(sims):
        host.simulate([roots[t] for t in live], 1.0, tree_ids=live)
    stats = gpu.root_stats()
    acts = np.full(B, -1, np.int32)
    for t in live:
        ids, n, q, p = host.stats(t, roots[t])
        gids, gn, gq, gp = stats[t]
        moves += 1
        if list(gn) != list(n):
            differ += 1
        acts[t] = int(ids[int(np.argmax(n))])
    # both follow the HOST's (reference-keyed) choice so the games stay comparable
    gpu.advance(torch.as_tensor(acts).cuda())
    nxt = host.child_states([roots[t] for t in live], [int(acts[t]) for t in live])
    for t, s in zip(live, nxt):
        roots[t] = s
gpu.check()
print(f"{N}x{N} {P}p, {sims} simulations per move, {B} games with tree reuse: {differ} of {moves} searches have root visit "
      f"counts that differ between the {'path-keyed lockstep' if lockstep else 'board-keyed fused'} device search and the board-keyed host search")
