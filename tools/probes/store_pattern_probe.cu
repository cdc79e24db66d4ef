// Probe: how much of the HBM write bandwidth survives different row-to-warp mappings of a 2 GB byte-mask write?
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o /tmp/store_probe tools/probes/store_pattern_probe.cu && /tmp/store_probe
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

constexpr int kRow = 30464;            // bytes per env row (128 B aligned)
constexpr int kChunks = kRow / 512;    // 59.5 -> 59 full + 1 partial; use 60 with guard

template <int POLICY>
__device__ __forceinline__ void st16(uint4 *p, uint4 v) {
    if (POLICY == 0) *p = v; else if (POLICY == 1) __stcs(p, v); else __stwt(p, v);
}

// warp per row (what step_kernel does): 3552 rows in flight
template <int POLICY>
__global__ void __launch_bounds__(256, 3) warp_per_row(unsigned char *out, int64_t rows) {
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    for (int64_t r = (int64_t)blockIdx.x * 8 + warp; r < rows; r += (int64_t)gridDim.x * 8) {
        unsigned char *row = out + r * kRow + 16 * lane;
#pragma unroll 6
        for (int c = 0; c < 60; ++c)
            if (c * 512 + 16 * lane < kRow) st16<POLICY>(reinterpret_cast<uint4 *>(row + c * 512), make_uint4(c, lane, 1, 0));
    }
}
// block per row: 8 warps share one row, 444 rows in flight
template <int POLICY>
__global__ void __launch_bounds__(256, 3) block_per_row(unsigned char *out, int64_t rows) {
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    for (int64_t r = blockIdx.x; r < rows; r += gridDim.x) {
        unsigned char *row = out + r * kRow + 16 * lane;
        for (int c = warp; c < 60; c += 8)
            if (c * 512 + 16 * lane < kRow) st16<POLICY>(reinterpret_cast<uint4 *>(row + c * 512), make_uint4(c, lane, 1, 0));
    }
}
// warp per row but rows interleaved in time: each warp writes ONE 1 KB piece of each of its rows per sweep (worst locality)
template <int POLICY>
__global__ void __launch_bounds__(256, 3) linear_fill(unsigned char *out, int64_t bytes) {
    int64_t i = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) * 16;
    const int64_t stride = (int64_t)gridDim.x * blockDim.x * 16;
    for (; i < bytes; i += stride) st16<POLICY>(reinterpret_cast<uint4 *>(out + i), make_uint4(1, 2, 3, 4));
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
    const char *names[3] = {"default", "cs", "wt"};
    float ms;
#define RUN(K, P, ARG, LABEL) ms = best_ms([&] { K<P><<<grid, 256>>>(out, ARG); }); \
    printf("%-14s %-8s %.3f ms  %.0f GB/s\n", LABEL, names[P], ms, bytes / ms / 1e6);
    RUN(warp_per_row, 0, rows, "warp_per_row") RUN(warp_per_row, 1, rows, "warp_per_row") RUN(warp_per_row, 2, rows, "warp_per_row")
    RUN(block_per_row, 0, rows, "block_per_row") RUN(block_per_row, 1, rows, "block_per_row")
    RUN(linear_fill, 0, bytes, "linear_fill") RUN(linear_fill, 1, bytes, "linear_fill")
    ms = best_ms([&] { linear_fill<0><<<148 * 16, 256>>>(out, bytes); });
    printf("%-14s %-8s %.3f ms  %.0f GB/s (grid 148*16)\n", "linear_fill", "default", ms, bytes / ms / 1e6);
    ms = best_ms([&] { cudaMemsetAsync(out, 1, bytes); });
    printf("%-14s %-8s %.3f ms  %.0f GB/s\n", "cudaMemset", "-", ms, bytes / ms / 1e6);
    printf("%s\n", cudaGetErrorString(cudaDeviceSynchronize()));
    return 0;
}
