#!/usr/bin/env python
"""One-off soak: long lockstep parity runs (GPU engine vs CPU oracle, every ply, every env) over several geometries.
   python tools/soak_parity.py [scale]     # ~1-2 minutes at scale 1 on a B200 box (the oracle is the slow side)"""
import sys
import time
from pathlib import Path
ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT)); sys.path.insert(0, str(ROOT / "tests"))
from blokus_rl_b200 import BlokusEngine
from oracle.oracle import Oracle
from helpers import lockstep

scale = float(sys.argv[1]) if len(sys.argv) > 1 else 1.0
total_steps = total_games = 0
t0 = time.time()
for (N, P, rule, n, plies, seed) in ((7, 2, 1, 800, 100, 21), (6, 4, 0, 400, 80, 22), (5, 4, 1, 400, 60, 23), (6, 2, 1, 400, 80, 24), (20, 4, 0, 1500, 200, 11), (20, 4, 1, 600, 150, 12), (20, 2, 0, 400, 120, 13),
                                     (14, 2, 1, 800, 150, 14), (14, 4, 0, 500, 150, 15), (7, 2, 0, 1500, 120, 16),
                                     (7, 4, 0, 400, 100, 17), (10, 2, 1, 400, 120, 18), (5, 2, 0, 400, 60, 19),
                                     (12, 4, 0, 300, 120, 25), (13, 2, 1, 300, 120, 26), (15, 4, 0, 300, 140, 27), (16, 2, 0, 300, 140, 28)):
    eng, orc = BlokusEngine(N, P, score_rule=rule), Oracle(N, P, rule)
    steps, games = lockstep(eng, orc, n=int(n * scale), plies=plies, seed=seed, env_id_base=seed * 100000, check_naive_every=1009)
    print(f"{N}x{N} {P}p rule {rule}: {steps} steps, {games} games finished -- bit-exact (states, masks bytes+bits, counts, flags, "
          f"terminal vectors, scores, sampled actions)", flush=True)
    total_steps += steps; total_games += games
    eng.close()
print(f"SOAK OK: {total_steps} env steps, {total_games} games, {time.time() - t0:.0f} s")

# ---- playouts: every logged action replayed through the oracle (same Philox stream), final scores and winners compared ----
import numpy as np
import torch
for (N, P, rule, n_roots, per_root, root_plies) in ((20, 4, 0, 48, 32, 24), (20, 4, 1, 24, 16, 40), (14, 2, 0, 48, 32, 10), (7, 2, 0, 64, 32, 2)):
    eng, orc = BlokusEngine(N, P, score_rule=rule), Oracle(N, P, rule)
    roots = []
    for g in range(n_roots):
        o = orc.new_state()
        for _ in range(root_plies):
            if orc.field(o, "done"):
                break
            orc.step(o, orc.sample_action(o, 5, g), fast=True)
        roots.append(o)
    words = torch.tensor(np.stack([orc.pack(o) for o in roots]).view(np.int32)).cuda()
    seed = 0xC0FFEE + N
    out = eng.rollout(words, per_root, seed=seed, log_actions=True)
    torch.cuda.synchronize()
    log = out.action_log.cpu().numpy().view(np.uint16)
    fs, win, plies = out.final_scores.cpu().numpy(), out.winners.cpu().numpy(), out.plies.cpu().numpy()
    nply = 0
    for r, root in enumerate(roots):
        for j in range(per_root):
            o = orc.copy(root)
            k = 0
            while not orc.field(o, "done"):
                a = int(log[r, j, k])
                assert a == orc.sample_action(o, seed, r * per_root + j, stream=1)
                assert orc.step(o, a, fast=True) == 0
                k += 1
            assert plies[r, j] == k and (fs[r, j] == orc.final_scores(o)[:P]).all() and win[r, j] == orc.winners(o)
            nply += k
    print(f"playouts {N}x{N} {P}p rule {rule}: {n_roots * per_root} playouts, {nply} plies replayed through the oracle -- identical", flush=True)
    eng.close()
print("SOAK PLAYOUTS OK")
