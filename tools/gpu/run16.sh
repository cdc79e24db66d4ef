set -x
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -15 > gpurun_out/pytest_gpu_r2_v6.txt; cat gpurun_out/pytest_gpu_r2_v6.txt
timeout 600 python bench.py --steps 200 --warmup 10 --no-cpu --no-extra > gpurun_out/bench_r2_v6_n1.json 2> gpurun_out/bench_r2_v6_n1.err; tail -c 600 gpurun_out/bench_r2_v6_n1.err
python -c "
import json; d=json.load(open('gpurun_out/bench_r2_v6_n1.json')); print(d['value'], d['e2e']['value'], d['e2e_mask_to_host'])"
