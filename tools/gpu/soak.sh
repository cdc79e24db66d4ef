mkdir -p gpurun_out
timeout 400 python tools/soak_parity.py 1 > gpurun_out/soak_parity_r2_v12.txt 2>&1; tail -8 gpurun_out/soak_parity_r2_v12.txt
