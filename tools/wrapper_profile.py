#!/usr/bin/env python
"""Where a single-env ply through the game wrapper (boundary B1) spends its host time: cProfile over ~2 s of random play."""
import cProfile
import pstats
import sys
import time
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
from blokus_rl_b200.backend import EngineBackend
from blokus_rl_b200.game_wrapper import BlokusGameWrapper

N, P = (int(sys.argv[1]), int(sys.argv[2])) if len(sys.argv) > 2 else (20, 4)
game = BlokusGameWrapper(board_size=N, number_of_players=P, backend=EngineBackend(N, P))


def play(seconds):
    plies, t0 = 0, time.perf_counter()
    while time.perf_counter() - t0 < seconds:
        s, p = game.get_init_board()
        while True:
            s, p = game.get_next_state(s, p, game.get_sample_move(s))
            game.get_valid_moves(s, p)
            plies += 1
            if game.get_game_ended(s) is not None:
                break
    return plies, time.perf_counter() - t0


play(0.5)
n, dt = play(2.0)
print(f"{N}x{N} {P}p: {n / dt:.0f} plies/s without the profiler ({dt / n * 1e6:.1f} us per ply)")
pr = cProfile.Profile()
pr.enable()
n, dt = play(2.0)
pr.disable()
print(f"{n} plies under the profiler")
pstats.Stats(pr).sort_stats("tottime").print_stats(22)
