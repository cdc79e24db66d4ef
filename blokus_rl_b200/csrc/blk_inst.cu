// blk_inst.cu -- instantiates the step / rollout / search kernels of ONE (N, P) specialisation.
// Compiled once per geometry: nvcc -DBLK_INST_N=20 -DBLK_INST_P=4 ... (see blokus_rl_b200/build.py).
#include "blk_kernels.cuh"
#include "blk_search.cuh"

#ifndef BLK_INST_N
#error "compile with -DBLK_INST_N=<board size or 0> -DBLK_INST_P=<players or 0>"
#endif
#define BLK_CAT_(a, b, c) kernels_##a##_##b
#define BLK_CAT(a, b) BLK_CAT_(a, b, )
#define BLK_SCAT_(a, b, c) search_kernels_##a##_##b
#define BLK_SCAT(a, b) BLK_SCAT_(a, b, )

namespace blk {
KernelSet BLK_CAT(BLK_INST_N, BLK_INST_P)() { return make_kernel_set<BLK_INST_N, BLK_INST_P>(); }
SearchKernelSet BLK_SCAT(BLK_INST_N, BLK_INST_P)() { return make_search_set<BLK_INST_N, BLK_INST_P>(); }
}  // namespace blk
