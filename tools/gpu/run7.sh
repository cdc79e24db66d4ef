set -x
mkdir -p gpurun_out
python tools/kernel_bench.py --only small > gpurun_out/kb_small_product.json 2>gpurun_out/kb.err; cat gpurun_out/kb_small_product.json
for v in small128 small64; do BLOKUS_B200_LIB=build_exp/lib_$v.so python tools/kernel_bench.py --only small > gpurun_out/kb_$v.json 2>>gpurun_out/kb.err; cat gpurun_out/kb_$v.json; done
tail -5 gpurun_out/kb.err
