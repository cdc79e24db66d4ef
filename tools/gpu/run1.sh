set -x
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.max.sm --format=csv
lscpu | grep -E "Model name|^CPU\(s\)|Flags" | cut -c1-300 > gpurun_out/host_cpu.txt
timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -25 > gpurun_out/pytest_gpu_r2_v1.txt; cat gpurun_out/pytest_gpu_r2_v1.txt
timeout 600 python bench.py --steps 200 --warmup 10 > gpurun_out/bench_r2_v1_n1.json 2> gpurun_out/bench_r2_v1_n1.err; tail -c 600 gpurun_out/bench_r2_v1_n1.err
timeout 300 python bench.py --impl reference --steps 20 --warmup 3 > gpurun_out/bench_r2_v1_ref.json 2>/dev/null
timeout 300 python bench.py --workload rollouts --steps 5 > gpurun_out/bench_r2_v1_rollouts_n1.json 2> gpurun_out/bench_r2_v1_rollouts.err; tail -c 400 gpurun_out/bench_r2_v1_rollouts.err
timeout 300 python bench.py --workload puct --steps 5 > gpurun_out/bench_r2_v1_puct_n1.json 2> gpurun_out/bench_r2_v1_puct.err; tail -c 400 gpurun_out/bench_r2_v1_puct.err
timeout 300 python bench.py --workload puct --roots 1 --sims 200 --chain 50 --steps 5 > gpurun_out/bench_r2_v1_puct_b1.json 2>> gpurun_out/bench_r2_v1_puct.err
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_r2_v1.csv python bench.py --steps 3 --warmup 3 --no-cpu --no-extra > gpurun_out/ncu_bench.log 2>&1
timeout 900 python tools/profile_all.py --capture --tag r2_v1 2>&1 | tail -30
ls -la gpurun_out
