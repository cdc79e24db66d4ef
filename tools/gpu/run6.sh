set -x
mkdir -p gpurun_out
for i in 1 2; do
python tools/kernel_bench.py --only step,rollout > gpurun_out/kb_product_$i.json 2>gpurun_out/kb.err; cat gpurun_out/kb_product_$i.json
BLOKUS_B200_LIB=build_exp/lib_guarded.so python tools/kernel_bench.py --only step,rollout > gpurun_out/kb_guarded_$i.json 2>>gpurun_out/kb.err; cat gpurun_out/kb_guarded_$i.json
BLOKUS_B200_ANY_ABI=1 BLOKUS_B200_LIB=build_exp/lib_old.so python tools/kernel_bench.py --only step,rollout > gpurun_out/kb_old_$i.json 2>>gpurun_out/kb.err; cat gpurun_out/kb_old_$i.json
done
tail -5 gpurun_out/kb.err
