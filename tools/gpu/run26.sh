set -x
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py tests/test_gpu_guard_bands.py tests/test_gpu_puct.py -x -q -p no:cacheprovider 2>&1 | tail -5
timeout 600 python tools/geometry_bench.py 2>&1 | tee gpurun_out/geometry_r2_v10.txt
