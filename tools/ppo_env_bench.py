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
