import sys, time
sys.path.insert(0, "/root/repo")
import torch
from blokus_rl_b200.vector_env import BlokusVectorEnv
env = BlokusVectorEnv(16384, board_size=7, num_players=2, seed=1)
env.reset()
def policy():
    return torch.multinomial(env.action_mask.float(), 1).squeeze(1).to(torch.int32)
for _ in range(5):
    env.step_device(policy(), check=False)
torch.cuda.synchronize()
from torch.profiler import profile, ProfilerActivity
with profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU]) as prof:
    for _ in range(20):
        a = policy()
        env.step_device(a, check=False)
    torch.cuda.synchronize()
print(prof.key_averages().table(sort_by="cuda_time_total", row_limit=18, max_name_column_width=70))
t0=time.perf_counter()
for _ in range(50):
    env.step_device(policy(), check=False)
torch.cuda.synchronize(); print("ms per vector step", (time.perf_counter()-t0)/50*1e3)
