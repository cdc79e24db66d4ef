set -x
mkdir -p gpurun_out
timeout 1200 python -m pytest tests/test_gpu_guard_bands.py tests/test_selfplay.py -m gpu -x -q 2>&1 | tail -30 > gpurun_out/pytest_gpu_r2_v4_guard.txt; cat gpurun_out/pytest_gpu_r2_v4_guard.txt
