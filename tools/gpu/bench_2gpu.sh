set -x
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511"
timeout 600 $TR tools/dist_check.py > gpurun_out/dist_check_r2_2gpu.txt 2> gpurun_out/dist_check_r2_2gpu.err; tail -12 gpurun_out/dist_check_r2_2gpu.txt; tail -5 gpurun_out/dist_check_r2_2gpu.err
timeout 600 $TR bench.py --gpus 2 --steps 200 --warmup 10 > gpurun_out/bench_r2_v1_n2.json 2> gpurun_out/bench_n2.err; tail -c 300 gpurun_out/bench_n2.err
timeout 600 $TR bench.py --gpus 2 --workload rollouts --steps 5 > gpurun_out/bench_r2_v1_rollouts_n2.json 2>> gpurun_out/bench_n2.err
timeout 600 $TR bench.py --gpus 2 --workload puct --steps 5 > gpurun_out/bench_r2_v1_puct_n2.json 2>> gpurun_out/bench_n2.err
cat gpurun_out/bench_r2_v1_n2.json | cut -c1-1500
