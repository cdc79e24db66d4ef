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
for (N, P, rule, n, plies, seed) in ((20, 4, 0, 1500, 200, 11), (20, 4, 1, 600, 150, 12), (20, 2, 0, 400, 120, 13),
                                     (14, 2, 1, 800, 150, 14), (14, 4, 0, 500, 150, 15), (7, 2, 0, 1500, 120, 16),
                                     (7, 4, 0, 400, 100, 17), (10, 2, 1, 400, 120, 18), (5, 2, 0, 400, 60, 19)):
    eng, orc = BlokusEngine(N, P, score_rule=rule), Oracle(N, P, rule)
    steps, games = lockstep(eng, orc, n=int(n * scale), plies=plies, seed=seed, env_id_base=seed * 100000, check_naive_every=1009)
    print(f"{N}x{N} {P}p rule {rule}: {steps} steps, {games} games finished -- bit-exact (states, masks bytes+bits, counts, flags, "
          f"terminal vectors, scores, sampled actions)", flush=True)
    total_steps += steps; total_games += games
    eng.close()
print(f"SOAK OK: {total_steps} env steps, {total_games} games, {time.time() - t0:.0f} s")
