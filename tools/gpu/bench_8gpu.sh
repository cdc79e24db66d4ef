set -x
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29533"
timeout 400 $TR tools/dist_check.py > gpurun_out/dist_check_r2_v11_8gpu.txt 2> gpurun_out/dist_check_8.err; tail -12 gpurun_out/dist_check_r2_v11_8gpu.txt; tail -3 gpurun_out/dist_check_8.err
timeout 400 $TR bench.py --gpus 8 --steps 200 --warmup 10 > gpurun_out/bench_r2_v11_n8.json 2> gpurun_out/bench_n8.err; tail -c 300 gpurun_out/bench_n8.err
timeout 300 $TR bench.py --gpus 8 --workload rollouts --steps 5 > gpurun_out/bench_r2_v11_rollouts_n8.json 2>> gpurun_out/bench_n8.err
timeout 300 $TR bench.py --gpus 8 --workload puct --steps 5 > gpurun_out/bench_r2_v11_puct_n8.json 2>> gpurun_out/bench_n8.err
python - <<'PY'
import json
for f in ["bench_r2_v11_n8","bench_r2_v11_rollouts_n8","bench_r2_v11_puct_n8"]:
    d=json.load(open(f"gpurun_out/{f}.json")); print(f, d["value"], d["e2e"]["value"], d["e2e"].get("per_rank_ms"), d.get("e2e_mask_to_host",{}).get("value"))
PY
