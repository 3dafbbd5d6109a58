// split.cu -- PyRanges.split for all keys at once.
//
// Replaces pyranges/methods/split.py:4-24 (per key: concat(Start, End).sort_values().drop_duplicates(), then every pair
// of consecutive points is one feature) -- the feature table pr.count_overlaps builds when none is given
// (pyranges/multioverlap.py:146-147).
//
//   keys    2n flattened points  base[g] + (x - lo[g])  (group origins keep the groups apart and in order)
//   sort    LSD radix sort of the points, no payload (sort.cu)
//   flags   point p opens a feature [key[p-1], key[p]) iff it differs from its predecessor and is not the first
//           point of its group (sorted positions [2 row_off[g], 2 row_off[g+1]) belong to group g)
//   scan    exclusive scan of the flags = output slot; the value at a group's first point = features before the group
//   emit    un-flatten both ends of every flagged pair
// 16 B/row read + 2n x 8 B x (2 x passes + 3) of sort/scan traffic: HBM-bound, dominated by the sort.
#include "common.cuh"

namespace iv {

constexpr int SP_THREADS = 256;

__global__ void __launch_bounds__(SP_THREADS)
k_split_keys(const int64_t* __restrict__ s, const int64_t* __restrict__ e, int64_t n, iv_groups_t G,
             uint64_t* __restrict__ keys) {
    __shared__ int span[2];
    const int64_t first = (int64_t)blockIdx.x * SP_THREADS;
    const int64_t last = first + SP_THREADS - 1 < n - 1 ? first + SP_THREADS - 1 : n - 1;
    const SegSpan sp = block_seg_span(G.row_off, G.n_groups, first, last, span);
    const int64_t i = first + threadIdx.x;
    if (i >= n) return;
    const int g = sp.of(G.row_off, i);
    const int64_t off = __ldg(G.base + g) - __ldg(G.lo + g);
    // the two points of a row sit next to each other: positions [2 row_off[g], 2 row_off[g+1]) hold group g
    reinterpret_cast<ulonglong2*>(keys)[i] = make_ulonglong2((uint64_t)(s[i] + off), (uint64_t)(e[i] + off));
}

// EMIT=false: pos[p] = flag.  EMIT=true: pos[] holds the exclusive scan; write the features.
template <bool EMIT>
__global__ void __launch_bounds__(SP_THREADS)
k_split_sweep(const uint64_t* __restrict__ keys, int64_t m, iv_groups_t G, int64_t* __restrict__ pos,
              int64_t* __restrict__ out_start, int64_t* __restrict__ out_end, int32_t* __restrict__ out_group) {
    __shared__ int span[2];
    const int64_t first = (int64_t)blockIdx.x * SP_THREADS;
    const int64_t last = first + SP_THREADS - 1 < m - 1 ? first + SP_THREADS - 1 : m - 1;
    const SegSpan sp = block_seg_span(G.row_off, G.n_groups, first >> 1, last >> 1, span);
    const int64_t p = first + threadIdx.x;
    if (p >= m) return;
    const int g = sp.of(G.row_off, p >> 1);
    const uint64_t k = keys[p];
    const bool open = p > 2 * __ldg(G.row_off + g) && keys[p - 1] != k;
    if (!EMIT) {
        pos[p] = open ? 1 : 0;
    } else if (open) {
        const int64_t j = pos[p];
        const int64_t off = __ldg(G.lo + g) - __ldg(G.base + g);
        out_start[j] = (int64_t)keys[p - 1] + off;
        out_end[j] = (int64_t)k + off;
        if (out_group) out_group[j] = g;
    }
}

// out[g] = features before group g (the scan value at the group's first point); out[n_groups] already holds the total
__global__ void k_split_group_off(iv_groups_t G, int64_t m, const int64_t* __restrict__ pos, int64_t* __restrict__ out) {
    const int g = blockIdx.x * blockDim.x + threadIdx.x;
    if (g >= G.n_groups) return;
    const int64_t p = 2 * __ldg(G.row_off + g);
    out[g] = p < m ? pos[p] : out[G.n_groups];
}

struct SplitLayout {
    uint64_t *k0, *k1;
    int64_t *pos, *partials;
    void* sort_temp;
};

static void plan_split(int64_t n, Bump& b, SplitLayout& L) {
    const size_t m = (size_t)(n > 0 ? 2 * n : 2);
    L.k0 = b.take<uint64_t>(m);
    L.k1 = b.take<uint64_t>(m);
    L.pos = b.take<int64_t>(m);
    L.partials = b.take<int64_t>(scan_i64_large_temp_elems((int64_t)m));
    L.sort_temp = b.take<char>(radix_sort_temp_bytes((int64_t)m));
}

// which ping-pong buffer the LSD sort of m keys on `bits` bits ends in (sort.cu)
static int sorted_buf(int64_t m, int bits) { return radix_sort_passes(m, bits) & 1; }

}  // namespace iv

using namespace iv;

extern "C" size_t iv_split_workspace_bytes(int64_t n) {
    Bump b(nullptr, 0);
    SplitLayout L;
    plan_split(n, b, L);
    return b.off + 256;
}

extern "C" int iv_split_count(const int64_t* starts, const int64_t* ends, int64_t n, const iv_groups_t* groups, void* ws,
                              size_t ws_bytes, int64_t* out_group_off, void* stream) {
    cudaStream_t st = (cudaStream_t)stream;
    IV_REQUIRE(groups && n >= 0 && out_group_off, IV_ERR_INVALID, "iv_split_count: bad arguments");
    IV_REQUIRE(n < ((int64_t)1 << 30), IV_ERR_RANGE, "iv_split_count: more than 2^30-1 rows");
    const iv_groups_t& G = *groups;
    if (n == 0) {
        IV_CUDA(cudaMemsetAsync(out_group_off, 0, (size_t)(G.n_groups + 1) * 8, st));
        return IV_OK;
    }
    IV_REQUIRE(starts && ends && ws, IV_ERR_INVALID, "iv_split_count: null argument");
    const int bits = bits_for((uint64_t)(G.total_span > 0 ? G.total_span : 1));
    IV_REQUIRE(bits <= 63, IV_ERR_RANGE, "iv_split_count: span needs %d bits", bits);
    Bump b(ws, ws_bytes);
    SplitLayout L;
    plan_split(n, b, L);
    IV_REQUIRE(b.fits(), IV_ERR_WORKSPACE, "iv_split_count: workspace %zu < %zu", ws_bytes, b.off);
    const int64_t m = 2 * n;
    k_split_keys<<<(unsigned)cdiv(n, SP_THREADS), SP_THREADS, 0, st>>>(starts, ends, n, G, L.k0);
    IV_LAUNCH_CHECK();
    int res = 0;
    int rc = radix_sort_pingpong(L.k0, L.k1, nullptr, nullptr, 0, m, 0, bits, L.sort_temp, radix_sort_temp_bytes(m), st,
                                 &res);
    if (rc != IV_OK) return rc;
    IV_REQUIRE(res == sorted_buf(m, bits), IV_ERR_INVALID, "iv_split_count: unexpected sort buffer");
    const unsigned nb = (unsigned)cdiv(m, SP_THREADS);
    k_split_sweep<false><<<nb, SP_THREADS, 0, st>>>(res ? L.k1 : L.k0, m, G, L.pos, nullptr, nullptr, nullptr);
    IV_LAUNCH_CHECK();
    rc = scan_exclusive_i64_large(L.pos, m, L.partials, out_group_off + G.n_groups, st);
    if (rc != IV_OK) return rc;
    k_split_group_off<<<(unsigned)cdiv(G.n_groups, 256), 256, 0, st>>>(G, m, L.pos, out_group_off);
    IV_LAUNCH_CHECK();
    return IV_OK;
}

extern "C" int iv_split_emit(int64_t n, const iv_groups_t* groups, const void* ws, int64_t n_features,
                             int64_t* out_start, int64_t* out_end, int32_t* out_group, void* stream) {
    cudaStream_t st = (cudaStream_t)stream;
    IV_REQUIRE(groups && n >= 0 && n_features >= 0 && n_features < 2 * (n > 0 ? n : 1), IV_ERR_INVALID,
               "iv_split_emit: bad arguments");
    if (n == 0 || n_features == 0) return IV_OK;
    IV_REQUIRE(ws && out_start && out_end, IV_ERR_INVALID, "iv_split_emit: null argument");
    const iv_groups_t& G = *groups;
    Bump b((void*)ws, (size_t)-1);
    SplitLayout L;
    plan_split(n, b, L);
    const int64_t m = 2 * n;
    const int bits = bits_for((uint64_t)(G.total_span > 0 ? G.total_span : 1));
    const uint64_t* keys = sorted_buf(m, bits) ? L.k1 : L.k0;
    k_split_sweep<true><<<(unsigned)cdiv(m, SP_THREADS), SP_THREADS, 0, st>>>(keys, m, G, L.pos, out_start, out_end,
                                                                              out_group);
    IV_LAUNCH_CHECK();
    return IV_OK;
}
