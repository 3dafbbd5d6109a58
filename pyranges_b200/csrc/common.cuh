// common.cuh -- shared host/device helpers of libpyranges_b200 (sm_100a).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <string.h>

#include "pyranges_b200.h"

namespace iv {

// ---- error plumbing ---------------------------------------------------------------------------
void set_error(const char* fmt, ...);

#define IV_CUDA(call)                                                                              \
    do {                                                                                           \
        cudaError_t _e = (call);                                                                   \
        if (_e != cudaSuccess) {                                                                   \
            iv::set_error("%s:%d: %s -> %s", __FILE__, __LINE__, #call, cudaGetErrorString(_e));   \
            return IV_ERR_CUDA;                                                                    \
        }                                                                                          \
    } while (0)

// every kernel launch of the library passes through here; iv_launch_count() reports the tally
extern unsigned long long g_launches;
#define IV_LAUNCH_CHECK()                                                                          \
    do {                                                                                           \
        __atomic_fetch_add(&iv::g_launches, 1ull, __ATOMIC_RELAXED);                               \
        IV_CUDA(cudaGetLastError());                                                               \
    } while (0)

#define IV_REQUIRE(cond, code, ...)                                                                \
    do {                                                                                           \
        if (!(cond)) {                                                                             \
            iv::set_error(__VA_ARGS__);                                                            \
            return (code);                                                                         \
        }                                                                                          \
    } while (0)

// ---- workspace bump allocator (caller-owned memory, 256-byte aligned slices) --------------------
struct Bump {
    char* p;
    size_t off, cap;
    Bump(void* base, size_t bytes) : p((char*)base), off(0), cap(bytes) {}
    template <typename T>
    T* take(size_t n) {
        size_t bytes = (n * sizeof(T) + 255) & ~(size_t)255;
        T* r = (T*)(p ? p + off : nullptr);
        off += bytes;
        return r;
    }
    bool fits() const { return off <= cap; }
};

static inline int64_t cdiv(int64_t a, int64_t b) { return (a + b - 1) / b; }
static inline int bits_for(uint64_t v) {  // number of bits needed to represent v (v=0 -> 1)
    int b = 1;
    while (b < 64 && (v >> b)) b++;
    return b;
}

// ---- internal radix sort (sort.cu) ----------------------------------------------------------------
size_t radix_sort_temp_bytes(int64_t n);
// number of scatter passes of a sort of n keys on nbits bits; the sorted data ends in buffer (passes & 1)
int radix_sort_passes(int64_t n, int nbits);
// Sorts keys (+ optional 4/8-byte values) on bits [begin_bit, end_bit) ping-ponging between
// (k0,v0) and (k1,v1); input is in (k0,v0).  *result_buf = 0/1 says where the sorted data is.
// several independent sorts of the same length in one launch sequence (blockIdx.y = job)
constexpr int RS_MAX_JOBS = 2;
struct SortJob {
    uint64_t *k0, *k1;
    void *v0, *v1;
    int vb;
};
int radix_sort_multi(const SortJob* jobs, int njobs, int64_t n, int begin_bit, int end_bit, void* temp,
                     size_t temp_bytes, cudaStream_t stream, int* result_buf);
int radix_sort_pingpong(uint64_t* k0, uint64_t* k1, void* v0, void* v1, int value_bytes, int64_t n,
                        int begin_bit, int end_bit, void* temp, size_t temp_bytes, cudaStream_t stream,
                        int* result_buf);

// single-block exclusive scans over small/medium arrays (scan.cu)
int scan_exclusive_i64_inplace(int64_t* data, int64_t n, int64_t* out_total, cudaStream_t stream);
// long arrays: three launches, partials = scan_i64_large_temp_elems(n) int64 of scratch
size_t scan_i64_large_temp_elems(int64_t n);
int scan_exclusive_i64_large(int64_t* data, int64_t n, int64_t* partials, int64_t* out_total, cudaStream_t stream);
int scan_exclusive_u32_large(uint32_t* data, int64_t n, uint32_t* temp_partials, cudaStream_t stream);
size_t scan_u32_large_temp_elems(int64_t n);

#ifdef __CUDACC__
// ---- device helpers ---------------------------------------------------------------------------------
constexpr unsigned FULL = 0xffffffffu;

__device__ __forceinline__ int lane_id() { return threadIdx.x & 31; }

// largest k in [0, nseg) with off[k] <= i   (off[0] = 0 <= i < off[nseg])
__device__ __forceinline__ int find_seg(const int64_t* __restrict__ off, int nseg, int64_t i) {
    int lo = 0, hi = nseg;  // answer in [lo, hi)
    while (hi - lo > 1) {
        int mid = (lo + hi) >> 1;
        if (__ldg(off + mid) <= i) lo = mid; else hi = mid;
    }
    return lo;
}
__device__ __forceinline__ int find_seg_range(const int64_t* __restrict__ off, int klo, int khi, int64_t i) {
    int lo = klo, hi = khi + 1;
    while (hi - lo > 1) {
        int mid = (lo + hi) >> 1;
        if (__ldg(off + mid) <= i) lo = mid; else hi = mid;
    }
    return lo;
}

// Block-cooperative segment lookup: the block covers rows [first, last]; most blocks sit inside
// one segment, so two searches by thread 0 replace one per thread.  smem2: int[2].
struct SegSpan {
    int klo, khi;
    __device__ __forceinline__ int of(const int64_t* __restrict__ off, int64_t i) const {
        return klo == khi ? klo : find_seg_range(off, klo, khi, i);
    }
};
__device__ __forceinline__ SegSpan block_seg_span(const int64_t* __restrict__ off, int nseg, int64_t first,
                                                  int64_t last, int* smem2) {
    if (threadIdx.x == 0) {
        smem2[0] = find_seg(off, nseg, first);
        smem2[1] = find_seg(off, nseg, last);
    }
    __syncthreads();
    SegSpan s;
    s.klo = smem2[0];
    s.khi = smem2[1];
    return s;
}

// ---- warp / block scans -------------------------------------------------------------------------------
template <typename T>
__device__ __forceinline__ T warp_incl_sum(T v) {
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        T t = __shfl_up_sync(FULL, v, o);
        if (lane_id() >= o) v += t;
    }
    return v;
}
template <typename T>
__device__ __forceinline__ T warp_incl_max(T v) {
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        T t = __shfl_up_sync(FULL, v, o);
        if (lane_id() >= o) v = t > v ? t : v;
    }
    return v;
}
template <typename T>
__device__ __forceinline__ T warp_sum(T v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(FULL, v, o);
    return v;
}
template <typename T>
__device__ __forceinline__ T warp_min(T v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        T t = __shfl_xor_sync(FULL, v, o);
        v = t < v ? t : v;
    }
    return v;
}
template <typename T>
__device__ __forceinline__ T warp_max(T v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        T t = __shfl_xor_sync(FULL, v, o);
        v = t > v ? t : v;
    }
    return v;
}

// exclusive block sum of one value per thread; sm: T[THREADS/32 + 1]; returns exclusive prefix, sets total.
// Contains __syncthreads(); sm may be reused after the call returns + one more barrier by the caller.
template <typename T, int THREADS>
__device__ __forceinline__ T block_excl_sum(T v, T* sm, T& total) {
    constexpr int W = THREADS / 32;
    T inc = warp_incl_sum(v);
    int w = threadIdx.x >> 5;
    if (lane_id() == 31) sm[w] = inc;
    __syncthreads();
    if (w == 0) {
        T x = lane_id() < W ? sm[lane_id()] : T(0);
        T xi = warp_incl_sum(x);
        if (lane_id() < W) sm[lane_id()] = xi - x;
        if (lane_id() == W - 1) sm[W] = xi;
    }
    __syncthreads();
    T r = sm[w] + inc - v;
    total = sm[W];
    __syncthreads();
    return r;
}
// inclusive block max; sm: T[THREADS/32 + 1]; ident = identity element
template <typename T, int THREADS>
__device__ __forceinline__ T block_incl_max(T v, T* sm, T ident, T& total) {
    constexpr int W = THREADS / 32;
    T inc = warp_incl_max(v);
    int w = threadIdx.x >> 5;
    if (lane_id() == 31) sm[w] = inc;
    __syncthreads();
    if (w == 0) {
        T x = lane_id() < W ? sm[lane_id()] : ident;
        T xi = warp_incl_max(x);
        T xe = __shfl_up_sync(FULL, xi, 1);
        if (lane_id() == 0) xe = ident;
        if (lane_id() < W) sm[lane_id()] = xe;
        if (lane_id() == W - 1) sm[W] = xi;
    }
    __syncthreads();
    T carry = sm[w];
    T r = carry > inc ? carry : inc;
    total = sm[W];
    __syncthreads();
    return r;
}
#endif  // __CUDACC__

}  // namespace iv
