mkdir -p gpurun_out
timeout 1200 python tools/soak_parity.py 2 > gpurun_out/soak_parity_r2_v7_x2.txt 2>&1; tail -8 gpurun_out/soak_parity_r2_v7_x2.txt
