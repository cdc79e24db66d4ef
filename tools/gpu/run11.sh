set -x
mkdir -p gpurun_out
python tools/kernel_bench.py --only step > gpurun_out/kb_u6.json 2>gpurun_out/kb.err; cat gpurun_out/kb_u6.json
for v in unroll3 unroll10 unroll15; do BLOKUS_B200_LIB=build_exp/lib_$v.so python tools/kernel_bench.py --only step > gpurun_out/kb_$v.json 2>>gpurun_out/kb.err; cat gpurun_out/kb_$v.json; done
timeout 600 python bench.py --steps 100 --warmup 5 --no-cpu > gpurun_out/bench_r2_v4_n1.json 2> gpurun_out/bench_r2_v4_n1.err; tail -c 800 gpurun_out/bench_r2_v4_n1.err
python -c "
import json; d=json.load(open('gpurun_out/bench_r2_v4_n1.json')); print({k:v for k,v in d['extra'].items() if 'ppo' in k})"
