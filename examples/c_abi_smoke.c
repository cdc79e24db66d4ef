/* Plain-C user of the C ABI (include/blokus_b200.h): no torch, no C++.  Resets 1,024 envs, takes one step with the
 * first legal action's id for player 0 (action 0 = monomino on the start corner), reads back flags and counts.
 *
 *   gcc -std=c99 -I include -I /usr/local/cuda/include examples/c_abi_smoke.c -o /tmp/c_abi_smoke \
 *       -L blokus_rl_b200 -lblokus_b200 -L /usr/local/cuda/lib64 -lcudart -Wl,-rpath,$PWD/blokus_rl_b200
 */
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <cuda_runtime_api.h>

#include "blokus_b200.h"

#define CHECK(x) do { int rc_ = (x); if (rc_ != 0) { fprintf(stderr, "%s failed: %d %s\n", #x, rc_, blk_last_error()); return 1; } } while (0)

int main(void) {
    blk_config cfg = {20, 4, 0, 0};
    blk_engine *eng = NULL;
    blk_info info;
    const int64_t n = 1024;
    CHECK(blk_create(&cfg, &eng));
    CHECK(blk_get_info(eng, &info));
    printf("A=%d state_words=%d mask_bytes=%d\n", info.num_actions, info.state_words, info.mask_bytes);

    uint32_t *state; int32_t *action, *count; uint8_t *mask, *flags;
    cudaMalloc((void **)&state, n * info.state_words * 4);
    cudaMalloc((void **)&action, n * 4);
    cudaMalloc((void **)&count, n * 4);
    cudaMalloc((void **)&mask, n * (size_t)info.mask_bytes);
    cudaMalloc((void **)&flags, n);
    cudaMemset(action, 0, n * 4);                      /* action id 0 for every env */

    CHECK(blk_reset(eng, state, n, NULL));
    blk_step_args a;
    memset(&a, 0, sizeof a);
    a.n = n; a.state_in = state; a.state_out = state; a.action = action;
    a.mask = mask; a.mask_format = BLK_MASK_BYTES; a.mask_stride = info.mask_bytes;
    a.legal_count = count; a.flags = flags;
    CHECK(blk_step(eng, &a, NULL));

    int32_t h_count[4]; uint8_t h_flags[4];
    cudaMemcpy(h_count, count, sizeof h_count, cudaMemcpyDeviceToHost);
    cudaMemcpy(h_flags, flags, sizeof h_flags, cudaMemcpyDeviceToHost);
    printf("after one ply: player 1 has %d legal actions, flags=%d\n", h_count[0], h_flags[0]);
    int ok = h_count[0] == 58 && h_flags[0] == 0;

    int32_t meta[4], nc; uint8_t cells[10];
    CHECK(blk_action_to_cells(eng, 30432, meta, cells, &nc));
    printf("last action: piece %d at (%d,%d), %d cells\n", meta[0], meta[2], meta[3], nc);
    blk_destroy(eng);
    cudaFree(state); cudaFree(action); cudaFree(count); cudaFree(mask); cudaFree(flags);
    puts(ok ? "c_abi_smoke ok" : "c_abi_smoke FAILED");
    return ok ? 0 : 1;
}
