set -x
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -15 > gpurun_out/pytest_gpu_r2_v3.txt; cat gpurun_out/pytest_gpu_r2_v3.txt
python tools/kernel_bench.py > gpurun_out/kb_product.json 2>gpurun_out/kb.err; cat gpurun_out/kb_product.json
for v in roll10x3 roll15x2; do BLOKUS_B200_LIB=build_exp/lib_$v.so python tools/kernel_bench.py --only rollout > gpurun_out/kb_$v.json 2>>gpurun_out/kb.err; cat gpurun_out/kb_$v.json; done
for v in step13x2 step28x1; do BLOKUS_B200_LIB=build_exp/lib_$v.so python tools/kernel_bench.py --only step > gpurun_out/kb_$v.json 2>>gpurun_out/kb.err; cat gpurun_out/kb_$v.json; done
tail -5 gpurun_out/kb.err
