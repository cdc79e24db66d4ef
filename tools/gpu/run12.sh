set -x
mkdir -p gpurun_out
python tools/kernel_bench.py --only small > gpurun_out/kb_s0.json 2>gpurun_out/kb.err; cat gpurun_out/kb_s0.json
for v in wb2 stag wb2stag; do BLOKUS_B200_LIB=build_exp/lib_$v.so python tools/kernel_bench.py --only small > gpurun_out/kb_$v.json 2>>gpurun_out/kb.err; cat gpurun_out/kb_$v.json; done
tail -3 gpurun_out/kb.err
