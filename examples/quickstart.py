#!/usr/bin/env python
"""Quick tour of the engine on one B200:  python examples/quickstart.py

1. 4,096 envs of 20x20 four-player Blokus played with the on-device uniform-random policy (full legal masks every ply);
2. the tensors a policy/value net consumes: obs float32 [B,8,20,20] + mask bool [B,30433];
3. 100,000 random playouts from a mid-game position;
4. PUCT searches with the trees on the GPU: 256 at once, and ONE tree the way the reference's MCTS player searches.
"""
import sys
from pathlib import Path

sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
import torch

from blokus_rl_b200 import BlokusEngine
from blokus_rl_b200.gpu_puct import GpuPuct

eng = BlokusEngine(board_size=20, num_players=4)          # raises EngineError without a B200: there is no CPU fallback
n = 4096
states = eng.new_states(n)                                 # int32 [n, 88]: 4 x 20 row bitboards + inventories + meta + scores
out = eng.step(states, None, mask="bytes", sample=True, seed=1)      # first masks + first random legal actions
games = 0
for ply in range(120):
    out = eng.step(states, out.next_action, mask="bytes", sample=True, seed=1, auto_reset=True)
    games += int((out.flags & 1).sum())
print(f"{n} envs x 120 plies: {games} games finished, mean legal actions now {out.legal_count.float().mean():.1f}")
print("mask :", tuple(out.mask.shape), out.mask.dtype, "| winners' terminal vector example:", out.terminal[(out.flags & 1).bool()][:1].tolist())

obs = eng.observe(states)                                  # float32 [n, 8, 20, 20], contiguous, on the GPU
print("obs  :", tuple(obs.shape), obs.dtype, obs.device)

roots = states[:100].contiguous()
ro = eng.rollout(roots, per_root=1000, seed=7)             # 100,000 playouts to the end of the game
print("playouts: mean final scores of root 0 =", ro.final_scores[0].float().mean(0).tolist(),
      "| value estimate =", (ro.value_sum[0] / 1000).tolist())

search = GpuPuct(eng, num_trees=256, max_simulations=64)   # uniform prior (the reference's DumbNet): whole simulations in one kernel;
search.set_roots(states[:256].contiguous())                # pass TorchNetEvaluator(net) for a real net (lockstep kernels + the net)
search.run(50, cpuct=1.0)                                  # 50 simulations of every tree: ONE launch
print("PUCT: most visited root actions of the first 5 trees:", search.best_actions()[:5].tolist())
meta, cells = eng.action_to_cells(int(search.best_actions()[0]))
print("      tree 0 plays piece", meta[0], "covering cells", cells)

# one tree, 200 simulations per move, tree kept across moves -- the reference's "mcts" arena player (players/mcts_player.py)
from blokus_rl_b200.backend import EngineBackend
from blokus_rl_b200.game_wrapper import BlokusGameWrapper
from blokus_rl_b200.players import MCTSPlayer, RandomPlayer
game = BlokusGameWrapper(board_size=20, number_of_players=4, backend=EngineBackend(engine=eng))
players = [MCTSPlayer(game, simulations=200)] + [RandomPlayer(game) for _ in range(3)]     # reference_compat=False: 16 warps, virtual loss
s, cur = game.get_init_board()
while game.get_game_ended(s) is None:
    s, cur = players[cur].update_state(s, cur)
print("MCTS (seat 0) vs three random players, terminal vector:", game.get_game_ended(s).tolist())
