#!/usr/bin/env python
"""Throughput of the gym-style vector env (boundary B2, config/ppo_blokus_7x7.yml shape: 7x7, 2 players, random-bot
opponent inside step) with a random masked policy on the device."""
import sys
import time
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
import torch
from blokus_rl_b200.vector_env import BlokusVectorEnv

for num_envs in (4, 1024, 16384):
    env = BlokusVectorEnv(num_envs, board_size=7, num_players=2, seed=1)
    env.reset()
    def policy():
        return torch.multinomial(env.action_mask.float(), 1).squeeze(1).to(torch.int32)
    for _ in range(5):
        env.step_device(policy(), check=False)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    steps, episodes = 200 if num_envs <= 1024 else 50, 0
    for _ in range(steps):
        obs, reward, done, _ = env.step_device(policy(), check=False)
        episodes += int(done.sum())
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    print(f"num_envs={num_envs:6d}: {num_envs * steps / dt:.3e} agent steps/s ({dt / steps * 1e3:.2f} ms per vector step, {episodes} episodes)")

# ---- the NumPy surface the reference's PPO loop calls (ppo/trainer.py:155, 385): np actions in, np obs / reward /
# terminated out, legal index lists per env -- every call synchronises with the host ----
import numpy as np
rng = np.random.default_rng(0)
for num_envs in (4, 256, 4096):          # config/ppo_blokus_7x7.yml uses num_envs = 4
    env = BlokusVectorEnv(num_envs, board_size=7, num_players=2, seed=1)
    env.reset()
    legal = env.get_attr("ai_possible_indexes")
    steps = 200 if num_envs <= 256 else 30
    t0 = time.perf_counter()
    for _ in range(steps):
        actions = np.array([ids[rng.integers(len(ids))] for ids in legal], dtype=np.int64)
        obs, reward, terminated, truncated, info = env.step(actions)
        legal = env.get_attr("ai_possible_indexes")
    dt = time.perf_counter() - t0
    print(f"host surface num_envs={num_envs:5d}: {num_envs * steps / dt:.3e} agent steps/s ({dt / steps * 1e3:.2f} ms per step + index lists)")
