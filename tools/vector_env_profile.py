#!/usr/bin/env python
"""Where a vector-env step (boundary B2, the PPO surface) spends its host time at the reference's default of 4 envs."""
import cProfile
import pstats
import sys
import time
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
import numpy as np
from blokus_rl_b200 import BlokusEngine
from blokus_rl_b200.vector_env import BlokusVectorEnv

E = int(sys.argv[1]) if len(sys.argv) > 1 else 4
env = BlokusVectorEnv(E, engine=BlokusEngine(7, 2), seed=1)
env.reset()
rng = np.random.default_rng(0)


def agent_step():
    ids, cnt = env.legal_ids_padded()
    acts = ids[np.arange(E), (rng.random(E) * cnt).astype(np.int64)].astype(np.int64)
    env.step(acts)


def run(reps):
    t0 = time.perf_counter()
    for _ in range(reps):
        agent_step()
    return time.perf_counter() - t0


run(20)
reps = 2000 if E <= 64 else 100
dt = run(reps)
print(f"{E} envs: {E * reps / dt:.0f} agent steps/s, {dt / reps * 1e6:.1f} us per vector step")
pr = cProfile.Profile()
pr.enable()
run(reps)
pr.disable()
pstats.Stats(pr).sort_stats("tottime").print_stats(28)
