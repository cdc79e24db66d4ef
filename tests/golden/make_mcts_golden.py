#!/usr/bin/env python
"""Generate tests/golden/mcts_golden.json by running the REFERENCE's own search,
/root/reference/blokus_rl/alphazero/mcts.py (loaded unmodified by file path: it imports only math + numpy),
over this repo's game wrapper on the CPU oracle, with deterministic fake nets (tests/fake_nets.py).

Run in the build container (needs /root/reference):   python tests/golden/make_mcts_golden.py
The vectors pin rows a8/a9 of SURVEY.md section 8: per-simulation score vectors, root visit counts N, running
means Q, priors P, and get_distribution at T=1 and T=0.
"""
import json
import sys
from pathlib import Path

import numpy as np

HERE = Path(__file__).resolve().parent
sys.path.insert(0, str(HERE.parents[1]))
sys.path.insert(0, str(HERE.parent))

import ref_stubs                                      # noqa: E402
from fake_nets import HashNet, UniformNet             # noqa: E402
from oracle_backend import OracleBackend              # noqa: E402
from blokus_rl_b200.game_wrapper import BlokusGameWrapper   # noqa: E402


def main():
    MCTS = ref_stubs.load_reference_mcts().MCTS
    cases = []
    for (N, P, plies_list) in ((20, 4, (0, 6, 20, 48)), (7, 2, (0, 3))):
        backend = OracleBackend(N, P)
        game = BlokusGameWrapper(board_size=N, number_of_players=P, backend=backend)
        for r, plies in enumerate(plies_list):
            rng = np.random.default_rng(1000 + r)
            s, player = game.get_init_board()
            for _ in range(plies):
                s, player = game.get_next_state(s, player, int(rng.choice(backend.legal_ids(s))))
            for net_name, net in (("uniform", UniformNet(P)), ("hash", HashNet(P))):
                for cpuct, sims in ((1.0, 40), (2.5, 40)):
                    tree = MCTS(game, net)
                    per_sim = [np.asarray(tree.simulate(s, player, cpuct), dtype=np.float64).tolist() for _ in range(sims)]
                    stats = tree.tree[game.string_representation(s)]
                    d1 = tree.get_distribution(s, 1)
                    d0 = tree.get_distribution(s, 0)
                    cases.append({
                        "board_size": N, "players": P, "plies": plies, "net": net_name, "cpuct": cpuct, "sims": sims,
                        "root_words": backend.words(s).tolist(), "root_player": int(player),
                        "scores": per_sim,
                        "ids": [int(a[0]) for a in stats[:, 0]],
                        "N": [float(x) for x in stats[:, 1]], "Q": [float(x) for x in stats[:, 2]],
                        "P": [float(x) for x in stats[:, 3]],
                        "dist_T1": [float(x) for x in d1[:, 1]], "dist_T0": [float(x) for x in d0[:, 1]],
                        "tree_size": len(tree.tree),
                    })
                    print(N, P, plies, net_name, cpuct, "tree", len(tree.tree), "maxN", max(cases[-1]["N"]))
    out = HERE / "mcts_golden.json"
    out.write_text(json.dumps({"generator": "tests/golden/make_mcts_golden.py",
                               "reference": "blokus_rl/alphazero/mcts.py (unmodified)", "numpy": np.__version__,
                               "cases": cases}))
    print("wrote", out, out.stat().st_size, "bytes")


if __name__ == "__main__":
    main()
