set -x
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29533"
timeout 600 $TR bench.py --gpus 8 --steps 200 --warmup 10 > gpurun_out/bench_r2_v7_n8.json 2> gpurun_out/bench_n8.err; tail -c 300 gpurun_out/bench_n8.err
timeout 600 $TR bench.py --gpus 8 --workload rollouts --steps 5 > gpurun_out/bench_r2_v7_rollouts_n8.json 2>> gpurun_out/bench_n8.err
timeout 600 $TR bench.py --gpus 8 --workload puct --steps 5 > gpurun_out/bench_r2_v7_puct_n8.json 2>> gpurun_out/bench_n8.err
timeout 300 python bench.py --impl reference --gpus 8 --steps 20 --warmup 3 > gpurun_out/bench_r2_v7_ref_n8box.json 2>/dev/null
python - <<'PY'
import json
for f in ["bench_r2_v7_n8","bench_r2_v7_rollouts_n8","bench_r2_v7_puct_n8","bench_r2_v7_ref_n8box"]:
    d=json.load(open(f"gpurun_out/{f}.json")); print(f, d["value"], d["e2e"]["value"], d["e2e"].get("per_rank_ms"), d.get("e2e_mask_to_host",{}).get("value"), d.get("cpu_baseline",{}).get("cores"))
PY
