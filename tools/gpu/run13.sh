set -x
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -15 > gpurun_out/pytest_gpu_r2_v5.txt; cat gpurun_out/pytest_gpu_r2_v5.txt
python tools/kernel_bench.py --only step,rollout > gpurun_out/kb_v5.json 2>gpurun_out/kb.err; cat gpurun_out/kb_v5.json; tail -3 gpurun_out/kb.err
python -c "import __graft_entry__ as g; g.smoke()"
