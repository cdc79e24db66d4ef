set -x
mkdir -p gpurun_out
T=r2_v7
timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -15 > gpurun_out/pytest_gpu_$T.txt; cat gpurun_out/pytest_gpu_$T.txt
python -c "import __graft_entry__ as g; g.smoke()"
timeout 900 python tools/profile_all.py --capture --tag $T 2>&1 | tail -4
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_$T.csv python bench.py --steps 3 --warmup 3 --no-cpu --no-extra > gpurun_out/ncu_bench.log 2>&1
