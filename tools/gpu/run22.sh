set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_puct.py tests/test_gpu_guard_bands.py tests/test_gpu_dropin_reference.py tests/test_selfplay.py tests/test_gpu_adapters.py -m gpu -x -q 2>&1 | tail -8
python - <<'PY'
import sys, time, torch
sys.path.insert(0,'.')
from blokus_rl_b200 import BlokusEngine
from blokus_rl_b200.gpu_puct import GpuPuct
eng = BlokusEngine(20,4)
for B in (1, 16, 4096):
    roots = eng.new_states(B)
    o = eng.step(roots, None, mask=None, sample=True, seed=5)
    for _ in range(24): o = eng.step(roots, o.next_action, mask=None, sample=True, seed=5)
    for wpt in ((1,8,16) if B == 1 else (1,)):
        sims = 200 if B < 4096 else 50
        s = GpuPuct(eng, num_trees=B, max_simulations=sims+8, mean_edges_per_node=500, warps_per_tree=wpt)
        def move():
            s.set_roots(roots); s.run(sims); return s.best_actions_device()
        for _ in range(3): move()
        torch.cuda.synchronize(); t0=time.perf_counter()
        for _ in range(5): move()
        torch.cuda.synchronize(); dt=(time.perf_counter()-t0)/5
        s.check()
        print(f"B {B} warps_per_tree {wpt:2d} sims {sims}: {B*sims/dt:.3e} sims/s ({dt*1e3:.2f} ms per move)")
        del s
PY
