set -x
mkdir -p gpurun_out
timeout 900 python tools/profile_all.py --capture --tag r2_v13b --only search_b1,search_b4096,search_b1_lp8,puct,puct_b1,step_14_2_bytes,step_runtime_10_2_bytes,small7_bytes,small7_bits,small7_rollout 2>&1 | tail -3
du -sh gpurun_out
