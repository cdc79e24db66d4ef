mkdir -p gpurun_out
cat > /tmp/sb.py <<'PY'
import sys, time, torch, os
sys.path.insert(0,'.')
from blokus_rl_b200 import BlokusEngine
from blokus_rl_b200.gpu_puct import GpuPuct
eng = BlokusEngine(20,4)
out=[]
for B in (1, 4096):
    roots = eng.new_states(B)
    o = eng.step(roots, None, mask=None, sample=True, seed=5)
    for _ in range(24): o = eng.step(roots, o.next_action, mask=None, sample=True, seed=5)
    sims = 200 if B < 4096 else 50
    s = GpuPuct(eng, num_trees=B, max_simulations=sims+8, mean_edges_per_node=500)
    def move():
        s.set_roots(roots); s.run(sims); return s.best_actions_device()
    for _ in range(3): move()
    torch.cuda.synchronize(); t0=time.perf_counter()
    for _ in range(5): move()
    torch.cuda.synchronize(); dt=(time.perf_counter()-t0)/5
    out.append(f"B{B}: {B*sims/dt:.3e}")
print(os.environ.get("BLOKUS_B200_LIB","product"), " ".join(out))
PY
python /tmp/sb.py
for v in noquads noprefetch neither; do BLOKUS_B200_LIB=build_exp/lib_$v.so python /tmp/sb.py; done
