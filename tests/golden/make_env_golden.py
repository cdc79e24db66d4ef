#!/usr/bin/env python
"""Generate tests/golden/env_golden.json: regression vectors for the env hot path (states, legal id lists, transitions,
terminal vectors), produced by the CPU oracle (naive cell-by-cell formulation).

These are NOT reference outputs -- the reference's env engine (colosseumrl) cannot be run -- they
pin the restatement, so that neither the oracle nor the CUDA engine can drift silently.   python tests/golden/make_env_golden.py
"""
import json
import sys
from pathlib import Path

import numpy as np

HERE = Path(__file__).resolve().parent
sys.path.insert(0, str(HERE.parents[1]))
from oracle.oracle import Oracle   # noqa: E402


def main():
    cases = []
    for (N, P, rule) in ((20, 4, 0), (7, 2, 0), (14, 2, 1)):
        orc = Oracle(N, P, rule)
        for seed in range(4):
            s = orc.new_state()
            ply = 0
            checkpoints = {0, 1, 3, 8, 15, 24, 33, 42, 50, 57}
            while True:
                done = bool(orc.field(s, "done"))
                if ply in checkpoints or done:
                    legal = np.flatnonzero(orc.legal_mask(s, fast=False)).tolist()
                    case = {"board_size": N, "players": P, "score_rule": rule, "seed": seed, "ply": ply,
                            "words": orc.pack(s).tolist(), "mover": int(orc.field(s, "mover")), "done": done,
                            "legal": legal, "scores": orc.final_scores(s)[:P].tolist(),
                            "terminal": orc.terminal_values(s).tolist(), "winners": int(orc.winners(s)),
                            "transitions": []}
                    for a in ([legal[0], legal[len(legal) // 2], legal[-1]] if legal else []):
                        n = orc.copy(s)
                        assert orc.step(n, a) == 0
                        case["transitions"].append({"action": int(a), "words": orc.pack(n).tolist(),
                                                    "done": bool(orc.field(n, "done")),
                                                    "terminal": orc.terminal_values(n).tolist()})
                    cases.append(case)
                if done:
                    break
                orc.step(s, orc.sample_action(s, 77, seed))
                ply += 1
    out = HERE / "env_golden.json"
    out.write_text(json.dumps({"generator": "tests/golden/make_env_golden.py", "source": "oracle/blokus_oracle.c (naive formulation)",
                               "cases": cases}))
    print("wrote", out, len(cases), "cases", out.stat().st_size, "bytes")


if __name__ == "__main__":
    main()
