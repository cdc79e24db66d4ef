set -x
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_puct.py -x -q -p no:cacheprovider 2>&1 | tail -3 || exit 1
timeout 1200 python -m pytest tests -m gpu -x -q -p no:cacheprovider 2>&1 | tail -3 | tee gpurun_out/pytest_gpu_r2_v13.txt
timeout 900 python tools/profile_all.py --capture --tag r2_v13a --only step_bytes,step_bits,step_indices,step_unaligned,leaf_expand,observe,rollout 2>&1 | tail -3
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_r2_v13.csv python bench.py --steps 3 --warmup 3 --no-cpu --no-extra > gpurun_out/ncu_bench.log 2>&1
du -sh gpurun_out
