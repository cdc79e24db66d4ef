// Probe: can TMA bulk stores (cp.async.bulk.global.shared::cta) from a small per-warp staging buffer write the 2 GB
// byte-mask stream faster than st.global.cs.v4 from registers?  Same warp-per-row mapping as step_kernel.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o /tmp/tma_probe tools/probes/tma_store_probe.cu && /tmp/tma_probe
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

constexpr int kRow = 30464;            // bytes per env row (128 B aligned)

__device__ __forceinline__ uint32_t smem_u32(const void *p) { return static_cast<uint32_t>(__cvta_generic_to_shared(p)); }

// warp per row, registers -> st.global.cs.v4 (reference pattern)
__global__ void __launch_bounds__(256, 3) reg_store(unsigned char *out, int64_t rows) {
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    for (int64_t r = (int64_t)blockIdx.x * 8 + warp; r < rows; r += (int64_t)gridDim.x * 8) {
        unsigned char *row = out + r * kRow + 16 * lane;
#pragma unroll 6
        for (int c = 0; c < 60; ++c)
            if (c * 512 + 16 * lane < kRow) __stcs(reinterpret_cast<uint4 *>(row + c * 512), make_uint4(c, lane, 1, 0));
    }
}

// warp per row, registers -> shared staging (CH bytes, double buffered) -> one bulk store per chunk
template <int CH>
__global__ void __launch_bounds__(256, 3) tma_store(unsigned char *out, int64_t rows) {
    __shared__ __align__(128) unsigned char stage[8][2][CH];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    constexpr int kChunks = (kRow + CH - 1) / CH;
    int buf = 0;
    for (int64_t r = (int64_t)blockIdx.x * 8 + warp; r < rows; r += (int64_t)gridDim.x * 8) {
        unsigned char *row = out + r * kRow;
        for (int c = 0; c < kChunks; ++c) {
            const int bytes = min(CH, kRow - c * CH);
            // the buffer we are about to overwrite was handed to the bulk copy two chunks ago: wait until it was READ
            if (lane == 0) asm volatile("cp.async.bulk.wait_group.read 1;" ::: "memory");
            __syncwarp();
            unsigned char *s = stage[warp][buf];
#pragma unroll
            for (int i = 16 * lane; i < CH; i += 512) *reinterpret_cast<uint4 *>(s + i) = make_uint4(c, lane, 1, 0);
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
            __syncwarp();
            if (lane == 0) {
                asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(row + c * CH), "r"(smem_u32(s)), "r"(bytes) : "memory");
                asm volatile("cp.async.bulk.commit_group;" ::: "memory");
            }
            buf ^= 1;
        }
    }
    if (lane == 0) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
}

template <typename F>
float best_ms(F f) {
    cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
    float best = 1e9f;
    for (int i = 0; i < 12; ++i) {
        cudaEventRecord(a); f(); cudaEventRecord(b); cudaEventSynchronize(b);
        float ms; cudaEventElapsedTime(&ms, a, b); if (i >= 2 && ms < best) best = ms;
    }
    return best;
}

int main() {
    const int64_t rows = 65536, bytes = rows * kRow;
    unsigned char *out; cudaMalloc(&out, bytes);
    const int grid = 148 * 3;
    float ms;
    ms = best_ms([&] { reg_store<<<grid, 256>>>(out, rows); });
    printf("%-22s %.3f ms  %.0f GB/s\n", "st.global.cs.v4", ms, bytes / ms / 1e6);
    ms = best_ms([&] { tma_store<512><<<grid, 256>>>(out, rows); });
    printf("%-22s %.3f ms  %.0f GB/s\n", "bulk store 512 B", ms, bytes / ms / 1e6);
    ms = best_ms([&] { tma_store<1024><<<grid, 256>>>(out, rows); });
    printf("%-22s %.3f ms  %.0f GB/s\n", "bulk store 1 KB", ms, bytes / ms / 1e6);
    ms = best_ms([&] { tma_store<2048><<<grid, 256>>>(out, rows); });
    printf("%-22s %.3f ms  %.0f GB/s\n", "bulk store 2 KB", ms, bytes / ms / 1e6);
    ms = best_ms([&] { tma_store<2048><<<148 * 4, 256>>>(out, rows); });
    printf("%-22s %.3f ms  %.0f GB/s (4 blocks/SM)\n", "bulk store 2 KB", ms, bytes / ms / 1e6);
    ms = best_ms([&] { cudaMemsetAsync(out, 1, bytes); });
    printf("%-22s %.3f ms  %.0f GB/s\n", "cudaMemset", ms, bytes / ms / 1e6);
    printf("%s\n", cudaGetErrorString(cudaDeviceSynchronize()));
    return 0;
}
