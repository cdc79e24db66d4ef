import sys
sys.path.insert(0, "/root/repo")
import torch
from blokus_rl_b200 import BlokusEngine
from blokus_rl_b200.gpu_puct import GpuPuct
from torch.profiler import profile, ProfilerActivity
eng = BlokusEngine(20, 4)
for B in (4096, 16384):
    roots = eng.new_states(B)
    o = eng.step(roots, None, mask=None, sample=True, seed=5)
    for _ in range(24):
        o = eng.step(roots, o.next_action, mask=None, sample=True, seed=5)
    s = GpuPuct(eng, num_trees=B, max_simulations=64, mean_edges_per_node=420, use_cuda_graph=False)
    s.set_roots(roots)
    for _ in range(10):
        s.simulate()
    torch.cuda.synchronize()
    with profile(activities=[ProfilerActivity.CUDA]) as prof:
        for _ in range(20):
            s.simulate()
        torch.cuda.synchronize()
    print("B =", B)
    print(prof.key_averages().table(sort_by="cuda_time_total", row_limit=8, max_name_column_width=60))
    del s
