// Probe: how long do N warps take when each one bumps the SAME global counter (a bump allocator) once or twice?
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o /tmp/atomic_probe tools/probes/atomic_probe.cu && /tmp/atomic_probe
#include <cstdio>
#include <cuda_runtime.h>
__global__ void bump(int *ctr, int *out, int n, int per_warp) {
    const int t = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
    if (t >= n) return;
    int v = 0;
    for (int k = 0; k < per_warp; ++k) {
        int r = 0;
        if (lane == 0) r = atomicAdd(ctr + k, 1 + (t & 7));
        v += __shfl_sync(0xffffffffu, r, 0);
    }
    if (lane == 0) out[t] = v;
}
int main() {
    int *ctr, *out; cudaMalloc(&ctr, 64); cudaMalloc(&out, 1 << 22);
    cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
    for (int n : {256, 4096, 16384, 65536})
        for (int per : {0, 1, 2}) {
            float best = 1e9f;
            for (int i = 0; i < 6; ++i) {
                cudaMemset(ctr, 0, 64);
                cudaEventRecord(a); bump<<<(n * 32 + 127) / 128, 128>>>(ctr, out, n, per); cudaEventRecord(b); cudaEventSynchronize(b);
                float ms; cudaEventElapsedTime(&ms, a, b); if (i && ms < best) best = ms;
            }
            printf("%6d warps x %d same-address atomics: %.1f us\n", n, per, best * 1e3f);
        }
    return 0;
}
