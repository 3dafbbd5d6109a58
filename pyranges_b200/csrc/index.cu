// index.cu -- subject index of the binary operations (replaces NCLS construction,
// reference call sites pyranges/methods/join.py:12, intersection.py:19,71, coverage.py:20,58).
//
// All keys (chromosome/strand groups) are indexed at once in a FLATTENED coordinate:
//     x' = clamp(x, lo[g], hi[g]) - lo[g] + base[g]
// so that one global sorted array serves every group and a query never needs a per-group search.
// Pieces:  key_s/row_s  starts sorted (+ subject row),  end_s  ends in start order,
//          key_e[/row_e] ends sorted,  max1/max2  block maxima of end_s (skip index for the
//          stabbing scan),  bin_s/bin_e  direct-address rank tables (bin = x' >> shift).
#include "common.cuh"

namespace iv {

constexpr int IX_THREADS = 256;

// ---- per-segment bounds (ingest-time metadata) ------------------------------------------------------------------
__global__ void k_bounds_init(int64_t* lo, int64_t* hi, int64_t* ml, int n) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) {
        lo[i] = INT64_MAX;
        hi[i] = INT64_MIN;
        ml[i] = 0;
    }
}

constexpr int BD_SPLIT = 64;
__global__ void __launch_bounds__(IX_THREADS)
k_bounds(const int64_t* __restrict__ s, const int64_t* __restrict__ e, const int64_t* __restrict__ seg_off,
         int64_t* lo, int64_t* hi, int64_t* ml) {
    __shared__ long long sm[3][IX_THREADS / 32];
    int seg = blockIdx.x;
    int64_t a = seg_off[seg], b = seg_off[seg + 1];
    int64_t len = b - a;
    int64_t chunk = (len + BD_SPLIT - 1) / BD_SPLIT;
    int64_t ca = a + chunk * blockIdx.y, cb = ca + chunk < b ? ca + chunk : b;
    long long mn = INT64_MAX, mx = INT64_MIN, m = 0;
    for (int64_t i = ca + threadIdx.x; i < cb; i += IX_THREADS) {
        long long x = s[i], y = e[i];
        mn = x < mn ? x : mn;
        mx = y > mx ? y : mx;
        m = (y - x) > m ? (y - x) : m;
    }
    mn = warp_min(mn);
    mx = warp_max(mx);
    m = warp_max(m);
    int w = threadIdx.x >> 5;
    if (lane_id() == 0) { sm[0][w] = mn; sm[1][w] = mx; sm[2][w] = m; }
    __syncthreads();
    if (threadIdx.x == 0) {
        for (int k = 1; k < IX_THREADS / 32; k++) {
            mn = sm[0][k] < mn ? sm[0][k] : mn;
            mx = sm[1][k] > mx ? sm[1][k] : mx;
            m = sm[2][k] > m ? sm[2][k] : m;
        }
        if (ca < cb) {
            atomicMin((long long*)lo + seg, mn);
            atomicMax((long long*)hi + seg, mx);
            atomicMax((long long*)ml + seg, m);
        }
    }
}

// ---- flatten subjects ----------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(IX_THREADS)
k_flatten_subjects(const int64_t* __restrict__ s, const int64_t* __restrict__ e, int64_t n, iv_groups_t G,
                   uint64_t* __restrict__ key_s, uint32_t* __restrict__ row_s, uint64_t* __restrict__ key_e,
                   uint32_t* __restrict__ row_e) {
    __shared__ int span[2];
    int64_t first = (int64_t)blockIdx.x * IX_THREADS;
    int64_t last = first + IX_THREADS - 1 < n - 1 ? first + IX_THREADS - 1 : n - 1;
    SegSpan sp = block_seg_span(G.row_off, G.n_groups, first, last, span);
    int64_t i = first + threadIdx.x;
    if (i >= n) return;
    int g = sp.of(G.row_off, i);
    int64_t off = __ldg(G.base + g) - __ldg(G.lo + g);
    key_s[i] = (uint64_t)(s[i] + off);
    key_e[i] = (uint64_t)(e[i] + off);
    row_s[i] = (uint32_t)i;
    if (row_e) row_e[i] = (uint32_t)i;
}

// end_s[p] = flattened end of the subject at start-sorted position p; block maxima per 64 / 4096
__global__ void __launch_bounds__(IX_THREADS)
k_gather_ends(const int64_t* __restrict__ e, int64_t n, iv_groups_t G, const uint32_t* __restrict__ row_s,
              uint64_t* __restrict__ end_s, uint64_t* __restrict__ max1, uint64_t* __restrict__ max2) {
    __shared__ int span[2];
    __shared__ unsigned long long m1[64];
    int64_t first = (int64_t)blockIdx.x * 4096;
    int64_t last = first + 4095 < n - 1 ? first + 4095 : n - 1;
    // sorted position p of group g lies in the same row range as the group's rows
    SegSpan sp = block_seg_span(G.row_off, G.n_groups, first, last, span);
    if (threadIdx.x < 64) m1[threadIdx.x] = 0;
    __syncthreads();
#pragma unroll 4
    for (int it = 0; it < 16; it++) {
        int li = it * IX_THREADS + threadIdx.x;
        int64_t p = first + li;
        unsigned long long v = 0;
        if (p < n) {
            int g = sp.of(G.row_off, p);
            v = (unsigned long long)(e[__ldg(row_s + p)] + (__ldg(G.base + g) - __ldg(G.lo + g)));
            end_s[p] = v;
        }
        // 64-entry chunk = two warps; reduce per warp then one smem atomic
        unsigned long long wm = warp_max(v);
        if (lane_id() == 0) atomicMax(&m1[li >> 6], wm);
    }
    __syncthreads();
    if (threadIdx.x < 64) {
        int64_t c = (first >> 6) + threadIdx.x;
        if ((c << 6) < n) max1[c] = m1[threadIdx.x];
    }
    if (threadIdx.x < 32) {
        unsigned long long a = m1[threadIdx.x], b = m1[threadIdx.x + 32];
        a = a > b ? a : b;
        a = warp_max(a);
        if (threadIdx.x == 0) max2[blockIdx.x] = a;
    }
}

// bin[b] = #keys < (b << shift)   for b in [0, n_bins]
__global__ void __launch_bounds__(IX_THREADS)
k_build_bins(const uint64_t* __restrict__ key_s, const uint64_t* __restrict__ key_e, int64_t n, int shift,
             int64_t n_bins, uint32_t* __restrict__ bin_s, uint32_t* __restrict__ bin_e) {
    int64_t b = (int64_t)blockIdx.x * IX_THREADS + threadIdx.x;
    if (b > n_bins) return;
    uint64_t x = (uint64_t)b << shift;
    const uint64_t* keys = blockIdx.y == 0 ? key_s : key_e;
    int64_t lo = 0, hi = n;
    while (lo < hi) {
        int64_t mid = (lo + hi) >> 1;
        if (__ldg(keys + mid) < x) lo = mid + 1; else hi = mid;
    }
    (blockIdx.y == 0 ? bin_s : bin_e)[b] = (uint32_t)lo;
}

struct IndexLayout {
    uint64_t *ks0, *ks1, *ke0, *ke1, *end_s, *max1, *max2;
    uint32_t *rs0, *rs1, *re0, *re1, *bin_s, *bin_e;
    void* sort_temp;
    int shift;
    int64_t n_bins;
};

static void choose_bins(int64_t n, int64_t span, int* shift, int64_t* n_bins) {
    // about two bins per subject, between 2^10 and 2^26 bins
    int64_t target = 1024;
    while (target < 2 * n && target < ((int64_t)1 << 26)) target <<= 1;
    int s = 0;
    while ((span >> s) > target) s++;
    *shift = s;
    *n_bins = (span >> s) + 1;
}

static void plan_index(int64_t n, int64_t span, int flags, Bump& b, IndexLayout& L) {
    size_t nn = (size_t)(n > 0 ? n : 1);
    choose_bins(n, span, &L.shift, &L.n_bins);
    L.ks0 = b.take<uint64_t>(nn);
    L.ks1 = b.take<uint64_t>(nn);
    L.rs0 = b.take<uint32_t>(nn);
    L.rs1 = b.take<uint32_t>(nn);
    L.ke0 = b.take<uint64_t>(nn);
    L.ke1 = b.take<uint64_t>(nn);
    if (flags & IV_INDEX_ENDS_PERM) {
        L.re0 = b.take<uint32_t>(nn);
        L.re1 = b.take<uint32_t>(nn);
    } else {
        L.re0 = L.re1 = nullptr;
    }
    L.end_s = b.take<uint64_t>(nn);
    L.max1 = b.take<uint64_t>((size_t)cdiv(nn, 64) + 64);
    L.max2 = b.take<uint64_t>((size_t)cdiv(nn, 4096) + 1);
    L.bin_s = b.take<uint32_t>((size_t)L.n_bins + 2);
    L.bin_e = b.take<uint32_t>((size_t)L.n_bins + 2);
    L.sort_temp = b.take<char>(radix_sort_temp_bytes(n));
}

}  // namespace iv

using namespace iv;

extern "C" int iv_segment_bounds(const int64_t* starts, const int64_t* ends, const int64_t* seg_off, int32_t n_seg,
                                 int64_t* out_lo, int64_t* out_hi, int64_t* out_maxlen, void* stream) {
    cudaStream_t st = (cudaStream_t)stream;
    IV_REQUIRE(n_seg >= 0, IV_ERR_INVALID, "iv_segment_bounds: n_seg < 0");
    if (n_seg == 0) return IV_OK;
    k_bounds_init<<<(n_seg + 255) / 256, 256, 0, st>>>(out_lo, out_hi, out_maxlen, n_seg);
    IV_LAUNCH_CHECK();
    k_bounds<<<dim3((unsigned)n_seg, BD_SPLIT), IX_THREADS, 0, st>>>(starts, ends, seg_off, out_lo, out_hi, out_maxlen);
    IV_LAUNCH_CHECK();
    return IV_OK;
}

extern "C" size_t iv_index_workspace_bytes(int64_t n, int64_t total_span, int32_t flags) {
    Bump b(nullptr, 0);
    IndexLayout L;
    plan_index(n, total_span > 0 ? total_span : 1, flags, b, L);
    return b.off + 256;
}

extern "C" int iv_index_build(const int64_t* starts, const int64_t* ends, int64_t n, const iv_groups_t* groups,
                              int32_t flags, void* ws, size_t ws_bytes, iv_index_t* out, void* stream) {
    cudaStream_t st = (cudaStream_t)stream;
    IV_REQUIRE(groups && out && n >= 0, IV_ERR_INVALID, "iv_index_build: null argument");
    IV_REQUIRE(n < ((int64_t)1 << 31), IV_ERR_RANGE, "iv_index_build: more than 2^31-1 subjects");
    const iv_groups_t& G = *groups;
    int64_t span = G.total_span > 0 ? G.total_span : 1;
    IV_REQUIRE(span < ((int64_t)1 << 62), IV_ERR_RANGE, "iv_index_build: flattened span too large");
    Bump b(ws, ws_bytes);
    IndexLayout L;
    plan_index(n, span, flags, b, L);
    IV_REQUIRE(b.fits(), IV_ERR_WORKSPACE, "iv_index_build: workspace %zu < %zu", ws_bytes, b.off);

    memset(out, 0, sizeof(*out));
    out->n = n;
    out->total_span = span;
    out->shift = L.shift;
    out->flags = flags;
    out->n_bins = L.n_bins;
    out->groups = G;
    out->key_s = L.ks0; out->row_s = L.rs0; out->key_e = L.ke0; out->row_e = L.re0;
    out->end_s = L.end_s; out->max1 = L.max1; out->max2 = L.max2; out->bin_s = L.bin_s; out->bin_e = L.bin_e;
    if (n == 0) {
        IV_CUDA(cudaMemsetAsync(L.bin_s, 0, ((size_t)L.n_bins + 2) * 4, st));
        IV_CUDA(cudaMemsetAsync(L.bin_e, 0, ((size_t)L.n_bins + 2) * 4, st));
        return IV_OK;
    }
    unsigned nb = (unsigned)cdiv(n, IX_THREADS);
    k_flatten_subjects<<<nb, IX_THREADS, 0, st>>>(starts, ends, n, G, L.ks0, L.rs0, L.ke0, L.re0);
    IV_LAUNCH_CHECK();
    int bits = bits_for((uint64_t)span);
    int res = 0;
    size_t tb = radix_sort_temp_bytes(n);
    int rc = radix_sort_pingpong(L.ks0, L.ks1, L.rs0, L.rs1, 4, n, 0, bits, L.sort_temp, tb, st, &res);
    if (rc != IV_OK) return rc;
    if (res) { out->key_s = L.ks1; out->row_s = L.rs1; }
    if (flags & IV_INDEX_ENDS_PERM)
        rc = radix_sort_pingpong(L.ke0, L.ke1, L.re0, L.re1, 4, n, 0, bits, L.sort_temp, tb, st, &res);
    else
        rc = radix_sort_pingpong(L.ke0, L.ke1, nullptr, nullptr, 0, n, 0, bits, L.sort_temp, tb, st, &res);
    if (rc != IV_OK) return rc;
    if (res) { out->key_e = L.ke1; out->row_e = (flags & IV_INDEX_ENDS_PERM) ? L.re1 : nullptr; }
    k_gather_ends<<<(unsigned)cdiv(n, 4096), IX_THREADS, 0, st>>>(ends, n, G, out->row_s, L.end_s, L.max1, L.max2);
    IV_LAUNCH_CHECK();
    k_build_bins<<<dim3((unsigned)cdiv(L.n_bins + 1, IX_THREADS), 2), IX_THREADS, 0, st>>>(
        out->key_s, out->key_e, n, L.shift, L.n_bins, L.bin_s, L.bin_e);
    IV_LAUNCH_CHECK();
    return IV_OK;
}
