// query.cuh -- query-side geometry shared by the overlap kernels (overlap.cu: rank index) and the fused join
// (join.cu: list index): the partner group's window of the flattened axis.
#pragma once
#include "common.cuh"

namespace iv {

// x' = base + clamp(x - lo, 0, width): the partner group's window of the flattened axis
template <typename X> struct Flat {
    int64_t lo;
    uint64_t width;
    X base;
    int g;  // partner group, -1 = none (lo = width = base = 0: every count comes out 0)
    __device__ __forceinline__ X at(int64_t x) const {
        uint64_t d = (uint64_t)(x - lo);
        d = x < lo ? 0ull : d;
        d = d > width ? width : d;
        return base + (X)d;
    }
    __device__ __forceinline__ bool outside(int64_t x) const { return x < lo || (uint64_t)(x - lo) > width; }
    __device__ __forceinline__ int64_t raw(uint64_t xf) const { return (int64_t)(xf - (uint64_t)base) + lo; }
};
// G: the groups of the subject index, n_subjects: its row count (0: nothing can overlap)
template <typename X>
__device__ __forceinline__ Flat<X> load_flat_groups(const iv_groups_t& G, int64_t n_subjects, const iv_queries_t& q, int seg) {
    Flat<X> f;
    f.g = __ldg(q.seg_group + seg);
    f.lo = 0; f.width = 0; f.base = 0;
    if (f.g >= 0 && n_subjects > 0) {
        f.lo = __ldg(G.lo + f.g);
        f.width = (uint64_t)(__ldg(G.hi + f.g) - f.lo);
        f.base = (X)__ldg(G.base + f.g);
    } else {
        f.g = -1;
    }
    return f;
}
__device__ __forceinline__ void apply_slack(int64_t slack, int64_t& s, int64_t& e) {
    if (slack != 0) {  // join(slack=k): pyranges/multithreaded.py:420-423
        s -= slack;
        if (s < 0) s = 0;
        e += slack;
    }
}

}  // namespace iv
