set -x
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -15 > gpurun_out/pytest_gpu_r2_v3.txt; cat gpurun_out/pytest_gpu_r2_v3.txt
python tools/kernel_bench.py > gpurun_out/kb_product.json 2>gpurun_out/kb.err; cat gpurun_out/kb_product.json
tail -5 gpurun_out/kb.err
