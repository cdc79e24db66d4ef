set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_puct.py tests/test_gpu_dropin_reference.py tests/test_gpu_adapters.py tests/test_selfplay.py tests/test_abi.py -m gpu -x -q 2>&1 | tail -30 > gpurun_out/pytest_gpu_r2_v2_puct.txt; cat gpurun_out/pytest_gpu_r2_v2_puct.txt
timeout 600 python bench.py --steps 100 --warmup 5 --no-cpu > gpurun_out/bench_r2_v2_n1.json 2> gpurun_out/bench_r2_v2_n1.err; tail -c 1500 gpurun_out/bench_r2_v2_n1.err
python -c "
import json; d=json.load(open('gpurun_out/bench_r2_v2_n1.json')); print({k:v for k,v in d['extra'].items() if 'mcts' in k and 'workload' not in k})"
timeout 300 python bench.py --workload puct --steps 5 > gpurun_out/bench_r2_v2_puct_n1.json 2> gpurun_out/bench_r2_v2_puct.err; tail -c 500 gpurun_out/bench_r2_v2_puct.err; cut -c1-300 gpurun_out/bench_r2_v2_puct_n1.json
