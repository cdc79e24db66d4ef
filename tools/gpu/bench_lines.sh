set -x
mkdir -p gpurun_out
timeout 900 python bench.py --steps 200 --warmup 10 > gpurun_out/bench_r2_v13_n1.json 2> gpurun_out/bench_r2_v13_n1.err; tail -c 400 gpurun_out/bench_r2_v13_n1.err
timeout 300 python bench.py --impl reference --steps 20 --warmup 3 > gpurun_out/bench_r2_v13_ref.json 2>/dev/null
timeout 300 python bench.py --workload rollouts --steps 5 > gpurun_out/bench_r2_v13_rollouts_n1.json 2> gpurun_out/bench_r2_v13_w.err
timeout 300 python bench.py --workload puct --steps 5 > gpurun_out/bench_r2_v13_puct_n1.json 2>> gpurun_out/bench_r2_v13_w.err
timeout 300 python bench.py --mask bits --steps 200 --no-cpu --no-extra > gpurun_out/bench_r2_v13_bits_n1.json 2>> gpurun_out/bench_r2_v13_w.err
timeout 300 python bench.py --board 7 --players 2 --envs 1048576 --steps 100 --no-cpu --no-extra > gpurun_out/bench_r2_v13_7x7_n1.json 2>> gpurun_out/bench_r2_v13_w.err
tail -c 300 gpurun_out/bench_r2_v13_w.err
