set -x
mkdir -p gpurun_out
timeout 900 python tools/profile_all.py --capture --tag r2_v5 2>&1 | tail -8
timeout 600 python bench.py --steps 200 --warmup 10 > gpurun_out/bench_r2_v5_n1.json 2> gpurun_out/bench_r2_v5_n1.err; tail -c 600 gpurun_out/bench_r2_v5_n1.err
timeout 300 python bench.py --impl reference --steps 20 --warmup 3 > gpurun_out/bench_r2_v5_ref.json 2>/dev/null
timeout 300 python bench.py --workload rollouts --steps 5 > gpurun_out/bench_r2_v5_rollouts_n1.json 2> gpurun_out/bench_r2_v5_rollouts.err
timeout 300 python bench.py --workload puct --steps 5 > gpurun_out/bench_r2_v5_puct_n1.json 2>> gpurun_out/bench_r2_v5_rollouts.err
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_r2_v5.csv python bench.py --steps 3 --warmup 3 --no-cpu --no-extra > gpurun_out/ncu_bench.log 2>&1
timeout 600 python tools/soak_parity.py 0.5 > gpurun_out/soak_parity_r2_v5.txt 2>&1; tail -4 gpurun_out/soak_parity_r2_v5.txt
