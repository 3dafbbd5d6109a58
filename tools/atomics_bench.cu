// Microbenchmark (not part of the library): throughput of random global atomics and scattered 8-byte stores, the two
// primitives an unstaged wide partition pass would be made of.  Build: nvcc -cudart shared -arch=sm_100a -O3 ...
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
__device__ __forceinline__ uint32_t mix(uint32_t x) { x ^= x >> 16; x *= 0x7feb352dU; x ^= x >> 15; x *= 0x846ca68bU; x ^= x >> 16; return x; }
__global__ void k_atomic(uint32_t* c, uint32_t mask, int64_t n, uint32_t* sink) {
    uint32_t acc = 0;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x)
        acc += atomicAdd(c + (mix((uint32_t)i) & mask), 1u);
    if (acc == 0xdeadbeef) *sink = acc;
}
__global__ void k_scatter(uint32_t* cur, uint32_t mask, uint32_t per, uint64_t* out, int64_t n) {
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        uint32_t b = mix((uint32_t)i) & mask;
        uint32_t p = atomicAdd(cur + b, 1u);
        if (p < per) out[(size_t)b * per + p] = (uint64_t)i;
    }
}
__global__ void k_store_only(uint32_t mask, uint32_t per, uint64_t* out, int64_t n) {
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        uint32_t b = mix((uint32_t)i) & mask;
        uint32_t p = (uint32_t)(i >> 0) / (mask + 1);  // ~ arrival order within the bucket
        if (p < per) out[(size_t)b * per + p] = (uint64_t)i;
    }
}
int main() {
    const int64_t n = 50000000;
    uint32_t *c, *sink; uint64_t* out;
    cudaMalloc(&c, (size_t)(1 << 24) * 4); cudaMalloc(&sink, 4);
    cudaMalloc(&out, (size_t)n * 8 * 2);
    cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
    for (int bits : {8, 12, 16, 19, 24}) {
        uint32_t mask = (1u << bits) - 1;
        float best = 1e9;
        for (int r = 0; r < 3; r++) {
            cudaMemset(c, 0, (size_t)(1 << 24) * 4);
            cudaEventRecord(a); k_atomic<<<148 * 8, 256>>>(c, mask, n, sink); cudaEventRecord(b); cudaEventSynchronize(b);
            float ms; cudaEventElapsedTime(&ms, a, b); if (ms < best) best = ms;
        }
        printf("atomicAdd(returning) 50M on 2^%d counters: %.3f ms = %.1f G/s\n", bits, best, n / best / 1e6);
        uint32_t per = (uint32_t)(2 * n / (mask + 1)) ; if (per < 4) per = 4;
        best = 1e9;
        for (int r = 0; r < 3; r++) {
            cudaMemset(c, 0, (size_t)(1 << 24) * 4);
            cudaEventRecord(a); k_scatter<<<148 * 8, 256>>>(c, mask, per, out, n); cudaEventRecord(b); cudaEventSynchronize(b);
            float ms; cudaEventElapsedTime(&ms, a, b); if (ms < best) best = ms;
        }
        printf("  atomic cursor + scattered 8-byte store: %.3f ms\n", best);
        best = 1e9;
        for (int r = 0; r < 3; r++) {
            cudaEventRecord(a); k_store_only<<<148 * 8, 256>>>(mask, per, out, n); cudaEventRecord(b); cudaEventSynchronize(b);
            float ms; cudaEventElapsedTime(&ms, a, b); if (ms < best) best = ms;
        }
        printf("  scattered 8-byte store only: %.3f ms\n", best);
    }
    printf("%s\n", cudaGetErrorString(cudaGetLastError()));
    return 0;
}
