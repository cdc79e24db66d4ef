import sys
sys.path.insert(0, "/root/repo")
import torch
from blokus_rl_b200 import BlokusEngine
from blokus_rl_b200.gpu_puct import GpuPuct
eng = BlokusEngine(20, 4)
B = 16384
roots = eng.new_states(B)
o = eng.step(roots, None, mask=None, sample=True, seed=5)
for _ in range(24):
    o = eng.step(roots, o.next_action, mask=None, sample=True, seed=5)
s = GpuPuct(eng, num_trees=B, max_simulations=64, mean_edges_per_node=420, use_cuda_graph=False)
s.set_roots(roots)
for _ in range(30):
    s.simulate()
torch.cuda.synchronize()
