set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_dropin_reference.py -m gpu -x -q 2>&1 | tail -15
