import sys
sys.path.insert(0, "/root/repo")
import torch
from blokus_rl_b200 import BlokusEngine
eng = BlokusEngine(7, 2)
roots = eng.new_states(1024)
o = eng.step(roots, None, mask=None, sample=True, seed=2)
for _ in range(2):
    o = eng.step(roots, o.next_action, mask=None, sample=True, seed=2)
for _ in range(4):
    eng.rollout(roots, 1024, seed=7)
torch.cuda.synchronize()
