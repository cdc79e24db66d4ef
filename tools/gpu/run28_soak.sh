mkdir -p gpurun_out
timeout 1500 python tools/soak_parity.py 1.5 > gpurun_out/soak_parity_r2_v11.txt 2>&1; tail -12 gpurun_out/soak_parity_r2_v11.txt
