#!/usr/bin/env python
"""Small end-to-end exercise of every kernel for compute-sanitizer (memcheck / racecheck), checked against the oracle."""
import sys
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
sys.path.insert(0, str(Path(__file__).resolve().parents[1] / "tests"))
import torch
from blokus_rl_b200 import BlokusEngine
from oracle.oracle import Oracle
from helpers import lockstep

for (n, p) in ((20, 4), (7, 2)):
    eng, orc = BlokusEngine(n, p), Oracle(n, p)
    print(n, p, lockstep(eng, orc, n=20, plies=12 if n == 20 else 30, seed=1))
    s = eng.new_states(40)
    out = eng.step(s, None, mask=None, sample=True, seed=2)
    for _ in range(10):
        out = eng.step(s, out.next_action, mask="bits", sample=True, seed=2, auto_reset=True)
    eng.observe(s); eng.board_contents(s); eng.game_ended(s)
    raw = torch.zeros((40, eng.num_actions), dtype=torch.uint8, device="cuda")
    eng.step(s, None, mask=raw)
    r = eng.rollout(s, 3, seed=5, log_actions=True)
    torch.cuda.synchronize()
    print("rollout plies", r.plies.float().mean().item())
print("sanitize_run ok")
