// Microbenchmark: random gathers of 4/8/16/32 bytes per lane from an L2-resident table (sm_100a).
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
template <int BYTES>
__device__ __forceinline__ uint32_t gather(const unsigned char* t, uint64_t idx) {
    const unsigned char* p = t + idx * BYTES;
    if (BYTES == 4) return __ldg((const uint32_t*)p);
    if (BYTES == 8) { uint2 v = __ldg((const uint2*)p); return v.x ^ v.y; }
    if (BYTES == 16) { uint4 v = __ldg((const uint4*)p); return v.x ^ v.y ^ v.z ^ v.w; }
    uint32_t a, b, c, d, e, f, g, h;
    asm("ld.global.nc.v8.u32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];" : "=r"(a), "=r"(b), "=r"(c), "=r"(d), "=r"(e), "=r"(f), "=r"(g), "=r"(h) : "l"(p));
    return a ^ b ^ c ^ d ^ e ^ f ^ g ^ h;
}
template <int BYTES, int ILP>
__global__ void k(const unsigned char* t, uint32_t nrec, uint32_t* out, int64_t n) {
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    uint32_t acc = 0;
    for (int64_t base = i * ILP; base < n; base += (int64_t)gridDim.x * blockDim.x * ILP) {
        uint32_t v[ILP];
#pragma unroll
        for (int j = 0; j < ILP; j++) {
            uint64_t x = (uint64_t)(base + j) * 0x9E3779B97F4A7C15ull;
            x ^= x >> 29;
            v[j] = gather<BYTES>(t, (uint32_t)(x >> 20) % nrec);
        }
#pragma unroll
        for (int j = 0; j < ILP; j++) acc += v[j];
    }
    if (acc == 0x12345678) out[0] = acc;
}
template <int BYTES, int ILP>
void run(const unsigned char* t, size_t table_bytes, uint32_t* out, int64_t n, int blocks, int threads) {
    uint32_t nrec = (uint32_t)(table_bytes / BYTES);
    cudaEvent_t a, b;
    cudaEventCreate(&a); cudaEventCreate(&b);
    for (int w = 0; w < 2; w++) k<BYTES, ILP><<<blocks, threads>>>(t, nrec, out, n);
    cudaEventRecord(a);
    for (int r = 0; r < 5; r++) k<BYTES, ILP><<<blocks, threads>>>(t, nrec, out, n);
    cudaEventRecord(b);
    cudaEventSynchronize(b);
    float ms; cudaEventElapsedTime(&ms, a, b);
    double us = ms * 1e3 / 5;
    printf("bytes=%2d ilp=%d blocks=%5d thr=%4d table=%3zu MB : %8.1f us for %lld gathers = %6.1f G/s, %6.1f GB/s gathered\n",
           BYTES, ILP, blocks, threads, table_bytes >> 20, us, (long long)n, n / us / 1e3, n * (double)BYTES / us / 1e3);
}
int main() {
    size_t tb = 24u << 20;
    unsigned char* t; uint32_t* out;
    cudaMalloc(&t, 256u << 20); cudaMemset(t, 1, 256u << 20); cudaMalloc(&out, 64);
    int64_t n = 10000000;
    for (size_t tbytes : {(size_t)12 << 20, (size_t)24 << 20, (size_t)96 << 20}) {
        run<4, 4>(t, tbytes, out, n, 148 * 8, 256);
        run<8, 4>(t, tbytes, out, n, 148 * 8, 256);
        run<16, 4>(t, tbytes, out, n, 148 * 8, 256);
        run<32, 4>(t, tbytes, out, n, 148 * 8, 256);
    }
    run<32, 1>(t, tb, out, n, 148 * 8, 256);
    run<32, 2>(t, tb, out, n, 148 * 8, 256);
    run<32, 8>(t, tb, out, n, 148 * 4, 256);
    run<32, 4>(t, tb, out, n, 148 * 2, 256);
    run<32, 4>(t, tb, out, n, 148 * 4, 512);
    run<16, 8>(t, tb, out, n, 148 * 8, 256);
    run<16, 2>(t, tb, out, n, 148 * 8, 256);
    run<4, 8>(t, tb, out, n, 148 * 8, 256);
    printf("%s\n", cudaGetErrorString(cudaGetLastError()));
    return 0;
}
