#!/usr/bin/env python
"""CPU rate of the REFERENCE's own search (blokus_rl/alphazero/mcts.py, unmodified) with the uniform DumbNet prior over
this repo's game wrapper on the CPU oracle -- BASELINE.md section 3 item (iii).  Needs /root/reference (build container only)."""
import sys
import time
from pathlib import Path
ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT)); sys.path.insert(0, str(ROOT / "tests"))
import numpy as np
import ref_stubs
from fake_nets import UniformNet
from oracle_backend import OracleBackend
from blokus_rl_b200.game_wrapper import BlokusGameWrapper

MCTS = ref_stubs.load_reference_mcts().MCTS
backend = OracleBackend(20, 4)
game = BlokusGameWrapper(board_size=20, number_of_players=4, backend=backend)
rng = np.random.default_rng(0)
total_sims, t0 = 0, time.perf_counter()
for root in range(8):
    s, p = game.get_init_board()
    for _ in range(24):
        s, p = game.get_next_state(s, p, int(rng.choice(backend.legal_ids(s))))
    tree = MCTS(game, UniformNet(4))
    for _ in range(100):                       # compare_arena's "mcts" player: MCTS with DumbNet, 100 simulations
        tree.simulate(s, p)
        total_sims += 1
dt = time.perf_counter() - t0
print(f"reference mcts.py over the CPU oracle: {total_sims} simulations in {dt:.1f} s = {total_sims / dt:.0f} sims/s (1 core; includes root setup)")
