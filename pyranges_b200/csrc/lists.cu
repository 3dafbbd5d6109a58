// lists.cu -- the list index of the join / count_overlaps fast path (replaces NCLS construction, reference call
// sites pyranges/methods/join.py:12, coverage.py:20, intersection.py:71).
//
// bin = flattened coordinate >> shift.  Every bin owns its TOUCH LIST: the subjects intersecting the bin -- those
// that started before it and are still open at its origin ("crossing"), then those that start inside it -- as entries
// {subject row; test word} in two parallel 32-bit arrays.  The test word holds the start and end offsets from the bin
// origin as two 15-bit fields under guard bits, laid out so that ONE 32-bit subtraction of the query's word decides
// both compares of the overlap predicate (pyranges_b200.h).  A query starting in bin b overlaps the entries of list b
// that pass, plus the non-crossing entries of the later bins it reaches (join.cu).
//
// No sort: the list lengths are the running sum of a difference array (+1 at the subject's first bin, -1 behind its
// last one: two atomics per subject), the list offsets the running sum of the lengths (both in one second-order scan
// pass), the fill takes slots with an atomic cursor per bin and a last pass orders every (short) list canonically --
// crossing entries by row, then starts by (start, row) -- so that the output order of a join is deterministic.  The
// same pass assembles the 32-byte bin records: list offset, length and the first three entries inline.
#include "common.cuh"

namespace iv {

constexpr int LS_THREADS = 256;
constexpr int LS_ITEMS = 8;
constexpr int LS_TILE = LS_THREADS * LS_ITEMS;  // bins per scan block
constexpr int LS_SORT_MAX = 24;                 // lists up to this length are ordered by one thread
constexpr int LS_BIG_SMEM = 2048;               // longer ones by a block: in shared memory up to this length

// ---- K1: difference array of the list lengths -------------------------------------------------------------------------
__global__ void __launch_bounds__(LS_THREADS)
k_lists_hist(const int64_t* __restrict__ s, const int64_t* __restrict__ e, int64_t n, iv_groups_t G, int shift,
             uint32_t reach, uint32_t* __restrict__ diff) {
    __shared__ int span[2];
    const int64_t first = (int64_t)blockIdx.x * LS_THREADS;
    const int64_t last = first + LS_THREADS - 1 < n - 1 ? first + LS_THREADS - 1 : n - 1;
    const SegSpan sp = block_seg_span(G.row_off, G.n_groups, first, last, span);
    const int64_t i = first + threadIdx.x;
    if (i >= n) return;
    const int g = sp.of(G.row_off, i);
    const int64_t off = __ldg(G.base + g) - __ldg(G.lo + g);
    const uint64_t S = (uint64_t)(__ldg(s + i) + off), E = (uint64_t)(__ldg(e + i) + off);
    const uint64_t b0 = S >> shift, b1 = E > S ? (E - 1) >> shift : b0;
    // a subject that starts within `reach` of its bin's origin is also listed in the bin before
    const uint64_t bf = (b0 > 0 && (uint32_t)(S - (b0 << shift)) < reach) ? b0 - 1 : b0;
    atomicAdd(diff + bf, 1u);
    atomicAdd(diff + b1 + 1, 0xffffffffu);
}

// ---- K2: second-order scan: length[b] = sum(diff[0..b]), offset[b] = sum(length[0..b-1]) ---------------------------------
// A run of bins is summarised by (D = sum of its diffs, O = sum of its lengths for an entering length of 0, L = bins);
// two consecutive runs combine to (D1 + D2, O1 + O2 + D1 * L2, L1 + L2).
struct T3 { long long D, O, L; };
__device__ __forceinline__ T3 comb(const T3& a, const T3& b) { return T3{a.D + b.D, a.O + b.O + a.D * b.L, a.L + b.L}; }
__device__ __forceinline__ T3 shfl_up3(const T3& v, int o) {
    return T3{__shfl_up_sync(FULL, v.D, o), __shfl_up_sync(FULL, v.O, o), __shfl_up_sync(FULL, v.L, o)};
}
__device__ __forceinline__ T3 shfl3(const T3& v, int l) {
    return T3{__shfl_sync(FULL, v.D, l), __shfl_sync(FULL, v.O, l), __shfl_sync(FULL, v.L, l)};
}
__device__ __forceinline__ T3 warp_incl3(T3 v) {
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const T3 t = shfl_up3(v, o);
        if (lane_id() >= o) v = comb(t, v);
    }
    return v;
}
// exclusive block scan under comb; sm: T3[THREADS / 32]; carry = what enters the block
template <int THREADS>
__device__ __forceinline__ T3 block_excl3(const T3& v, T3* sm, const T3& carry) {
    constexpr int W = THREADS / 32;
    const int w = threadIdx.x >> 5, lane = lane_id();
    const T3 inc = warp_incl3(v);
    if (lane == 31) sm[w] = inc;
    __syncthreads();
    if (w == 0) {
        T3 x = lane < W ? sm[lane] : T3{0, 0, 0};
        const T3 xi = warp_incl3(x);
        T3 xe = shfl_up3(xi, 1);
        if (lane == 0) xe = T3{0, 0, 0};
        if (lane < W) sm[lane] = xe;
    }
    __syncthreads();
    T3 ex = shfl_up3(inc, 1);
    if (lane == 0) ex = T3{0, 0, 0};
    const T3 r = comb(comb(carry, sm[w]), ex);
    __syncthreads();
    return r;
}

__global__ void __launch_bounds__(LS_THREADS)
k_lists_reduce(const uint32_t* __restrict__ diff, int64_t nb, long long* __restrict__ tD, long long* __restrict__ tO) {
    __shared__ long long sm[2][LS_THREADS / 32];
    const int64_t base = (int64_t)blockIdx.x * LS_TILE;
    const int64_t L = nb - base < LS_TILE ? nb - base : LS_TILE;
    long long D = 0, O = 0;
#pragma unroll
    for (int it = 0; it < LS_ITEMS; it++) {
        const int64_t idx = it * LS_THREADS + threadIdx.x;
        if (idx < L) {
            const long long d = (int)__ldg(diff + base + idx);
            D += d;
            O += d * (L - idx);
        }
    }
    D = warp_sum(D);
    O = warp_sum(O);
    if (lane_id() == 0) { sm[0][threadIdx.x >> 5] = D; sm[1][threadIdx.x >> 5] = O; }
    __syncthreads();
    if (threadIdx.x == 0) {
        for (int k = 1; k < LS_THREADS / 32; k++) { D += sm[0][k]; O += sm[1][k]; }
        tD[blockIdx.x] = D;
        tO[blockIdx.x] = O;
    }
}

// single block: (length, offset) entering every tile
__global__ void __launch_bounds__(1024)
k_lists_scan_tiles(const long long* __restrict__ tD, const long long* __restrict__ tO, int64_t ntiles, int64_t nb,
                   long long* __restrict__ tC, long long* __restrict__ tOin) {
    __shared__ T3 sm[32];
    const int64_t per = (ntiles + 1023) / 1024;
    const int64_t a = (int64_t)threadIdx.x * per, b = a + per < ntiles ? a + per : ntiles;
    auto len = [&](int64_t t) { return (long long)(nb - t * LS_TILE < LS_TILE ? nb - t * LS_TILE : LS_TILE); };
    T3 mine{0, 0, 0};
    for (int64_t t = a; t < b; t++) mine = comb(mine, T3{tD[t], tO[t], len(t)});
    T3 run = block_excl3<1024>(mine, sm, T3{0, 0, 0});
    for (int64_t t = a; t < b; t++) {
        tC[t] = run.D;
        tOin[t] = run.O;
        run = comb(run, T3{tD[t], tO[t], len(t)});
    }
}

// offsets of every bin; the difference array is cleared on the way (it becomes the fill cursor)
__global__ void __launch_bounds__(LS_THREADS)
k_lists_apply(uint32_t* __restrict__ diff, int64_t nb, const long long* __restrict__ tC, const long long* __restrict__ tOin,
              uint32_t* __restrict__ co) {
    __shared__ T3 sm[LS_THREADS / 32];
    const int64_t b0 = (int64_t)blockIdx.x * LS_TILE + (int64_t)threadIdx.x * LS_ITEMS;
    int d[LS_ITEMS];
    {
        // the arrays are padded to a whole tile: two 16-byte loads, always in bounds
        const uint4 x = *reinterpret_cast<const uint4*>(diff + b0), y = *reinterpret_cast<const uint4*>(diff + b0 + 4);
        d[0] = (int)x.x; d[1] = (int)x.y; d[2] = (int)x.z; d[3] = (int)x.w;
        d[4] = (int)y.x; d[5] = (int)y.y; d[6] = (int)y.z; d[7] = (int)y.w;
    }
    T3 mine{0, 0, LS_ITEMS};
#pragma unroll
    for (int j = 0; j < LS_ITEMS; j++) {
        if (b0 + j >= nb) d[j] = 0;
        mine.D += d[j];
        mine.O += (long long)d[j] * (LS_ITEMS - j);
    }
    const T3 in = block_excl3<LS_THREADS>(mine, sm, T3{__ldg(tC + blockIdx.x), __ldg(tOin + blockIdx.x), 0});
    long long C = in.D, O = in.O;
    uint32_t o[LS_ITEMS];
#pragma unroll
    for (int j = 0; j < LS_ITEMS; j++) {
        C += d[j];
        o[j] = (uint32_t)O;
        O += C;
    }
    *reinterpret_cast<uint4*>(co + b0) = make_uint4(o[0], o[1], o[2], o[3]);
    *reinterpret_cast<uint4*>(co + b0 + 4) = make_uint4(o[4], o[5], o[6], o[7]);
    *reinterpret_cast<uint4*>(diff + b0) = make_uint4(0u, 0u, 0u, 0u);
    *reinterpret_cast<uint4*>(diff + b0 + 4) = make_uint4(0u, 0u, 0u, 0u);
}

// ---- K3: fill ---------------------------------------------------------------------------------------------------------------
// Every (subject, touched bin) pair is one unit of work, spread over the block's threads whatever the subject lengths
// (a 2 Mb gene touches a thousand bins, an exon one).
__global__ void __launch_bounds__(LS_THREADS)
k_lists_fill(const int64_t* __restrict__ s, const int64_t* __restrict__ e, int64_t n, iv_groups_t G, int shift,
             uint32_t reach, const uint32_t* __restrict__ co, uint32_t* __restrict__ cursor, int64_t cap, uint32_t* __restrict__ cand_se,
             uint32_t* __restrict__ cand_row) {
    __shared__ int span[2];
    __shared__ uint32_t pre[LS_THREADS + 1];
    __shared__ uint32_t wsum[LS_THREADS / 32 + 1];
    __shared__ uint64_t sS[LS_THREADS], sE[LS_THREADS];
    const int64_t first = (int64_t)blockIdx.x * LS_THREADS;
    const int64_t last = first + LS_THREADS - 1 < n - 1 ? first + LS_THREADS - 1 : n - 1;
    const SegSpan sp = block_seg_span(G.row_off, G.n_groups, first, last, span);
    const int64_t i = first + threadIdx.x;
    uint32_t cnt = 0;
    if (i < n) {
        const int g = sp.of(G.row_off, i);
        const int64_t off = __ldg(G.base + g) - __ldg(G.lo + g);
        const uint64_t S = (uint64_t)(__ldg(s + i) + off), E = (uint64_t)(__ldg(e + i) + off);
        sS[threadIdx.x] = S;
        sE[threadIdx.x] = E;
        const uint64_t b0 = S >> shift, b1 = E > S ? (E - 1) >> shift : b0;
        const uint64_t bf = (b0 > 0 && (uint32_t)(S - (b0 << shift)) < reach) ? b0 - 1 : b0;
        cnt = (uint32_t)(b1 - bf + 1);
    }
    uint32_t tot;
    pre[threadIdx.x] = block_excl_sum<uint32_t, LS_THREADS>(cnt, wsum, tot);
    if (threadIdx.x == 0) pre[LS_THREADS] = tot;
    __syncthreads();
    for (uint32_t u = threadIdx.x; u < tot; u += LS_THREADS) {
        int lo = 0, hi = LS_THREADS;  // owner = last thread whose prefix <= u
        while (hi - lo > 1) {
            const int mid = (lo + hi) >> 1;
            if (pre[mid] <= u) lo = mid; else hi = mid;
        }
        const uint64_t S = sS[lo], E = sE[lo];
        const uint64_t b0 = S >> shift;
        const uint64_t bf = (b0 > 0 && (uint32_t)(S - (b0 << shift)) < reach) ? b0 - 1 : b0;
        const uint64_t bb = bf + (u - pre[lo]);
        const uint64_t origin = bb << shift;
        // 0 = starts before the bin; 1 .. width: starts inside; beyond: starts in the next bin, within reach
        const uint32_t sp = bb > b0 ? 0u : (uint32_t)(S - origin) + 1u;
        const uint64_t d = E - origin;
        const uint32_t eo = d > (uint64_t)IV_L_EMAX ? IV_L_EMAX : (uint32_t)d;
        const int64_t at = (int64_t)__ldg(co + bb) + atomicAdd(cursor + bb, 1u);
        if (at < cap) {
            cand_row[at] = (uint32_t)(first + lo);
            cand_se[at] = IV_L_GUARD | (IV_L_SMAX - sp) | (eo << 16);
        }
    }
}

// ---- K4: canonical order inside every list + the bin records -------------------------------------------------------------
// crossing entries by row, then starts by (start offset, row)
__device__ __forceinline__ unsigned long long entry_key(const uint2& en) {
    return ((unsigned long long)(IV_L_SMAX - (en.y & 0x7fffu)) << 32) | en.x;  // (start offset + 1, 0 = crossing; row)
}
__device__ __forceinline__ void write_rec(iv_lrec_t* __restrict__ rec, int64_t b, uint32_t c0, uint32_t n, const uint2* first3) {
    const uint2 z = make_uint2(0u, IV_L_GUARD);  // an entry no query hits (end offset 0)
    const uint2 a = n > 0 ? first3[0] : z, bq = n > 1 ? first3[1] : z, c = n > 2 ? first3[2] : z;
    uint4* out = reinterpret_cast<uint4*>(rec + b);
    out[0] = make_uint4(c0, n < 255u ? n : 255u, a.x, a.y);
    out[1] = make_uint4(bq.x, bq.y, c.x, c.y);
}

__global__ void __launch_bounds__(LS_THREADS)
k_lists_finalize(int64_t n_bins, const uint32_t* __restrict__ co, uint32_t* __restrict__ cand_se,
                 uint32_t* __restrict__ cand_row, iv_lrec_t* __restrict__ rec, uint32_t* __restrict__ big,
                 uint32_t* __restrict__ nbig) {
    const int64_t b = (int64_t)blockIdx.x * LS_THREADS + threadIdx.x;
    if (b > n_bins + 1) return;
    const uint32_t c0 = __ldg(co + b);
    const uint32_t n = b <= n_bins ? __ldg(co + b + 1) - c0 : 0u;
    uint2 a[LS_SORT_MAX];
    if (n > (uint32_t)LS_SORT_MAX) {  // a block orders it and writes the record (k_lists_big)
        big[atomicAdd(nbig, 1u)] = (uint32_t)b;
        return;
    }
    for (uint32_t j = 0; j < n; j++) {  // insertion sort: the lists are a handful of entries long
        const uint2 x = make_uint2(cand_row[c0 + j], cand_se[c0 + j]);
        const unsigned long long kx = entry_key(x);
        uint32_t p = j;
        while (p > 0 && entry_key(a[p - 1]) > kx) { a[p] = a[p - 1]; p--; }
        a[p] = x;
    }
    if (n > 1)
        for (uint32_t j = 0; j < n; j++) { cand_row[c0 + j] = a[j].x; cand_se[c0 + j] = a[j].y; }
    write_rec(rec, b, c0, n, a);
}

// long lists (pile-ups): one block per list, normalised bitonic network (every compare-exchange ascending, so the
// virtual padding to a power of two needs no sentinels), in shared memory when the list fits
__global__ void __launch_bounds__(LS_THREADS)
k_lists_big(const uint32_t* __restrict__ co, uint32_t* __restrict__ cand_se, uint32_t* __restrict__ cand_row,
            iv_lrec_t* __restrict__ rec, const uint32_t* __restrict__ big, const uint32_t* __restrict__ nbig) {
    __shared__ uint2 sh[LS_BIG_SMEM];
    const uint32_t total = *nbig;
    for (uint32_t k = blockIdx.x; k < total; k += gridDim.x) {
        const int64_t b = big[k];
        const uint32_t c0 = co[b], n = co[b + 1] - c0;
        uint32_t* gr = cand_row + c0;
        uint32_t* gs = cand_se + c0;
        const bool in_smem = n <= (uint32_t)LS_BIG_SMEM;
        if (in_smem)
            for (uint32_t i = threadIdx.x; i < n; i += LS_THREADS) sh[i] = make_uint2(gr[i], gs[i]);
        __syncthreads();
        uint32_t P = 1;
        while (P < n) P <<= 1;
        auto step = [&](uint32_t mask) {
            for (uint32_t i = threadIdx.x; i < n; i += LS_THREADS) {
                const uint32_t p = i ^ mask;
                if (p > i && p < n) {
                    if (in_smem) {
                        const uint2 x = sh[i], y = sh[p];
                        if (entry_key(x) > entry_key(y)) { sh[i] = y; sh[p] = x; }
                    } else {
                        const uint2 x = make_uint2(gr[i], gs[i]), y = make_uint2(gr[p], gs[p]);
                        if (entry_key(x) > entry_key(y)) { gr[i] = y.x; gs[i] = y.y; gr[p] = x.x; gs[p] = x.y; }
                    }
                }
            }
            __syncthreads();
        };
        for (uint32_t kk = 2; kk <= P; kk <<= 1) {
            step(kk - 1);  // first step of a merge: the partner is the mirror position inside the block of kk
            for (uint32_t j = kk >> 2; j > 0; j >>= 1) step(j);
        }
        if (in_smem)
            for (uint32_t i = threadIdx.x; i < n; i += LS_THREADS) { gr[i] = sh[i].x; gs[i] = sh[i].y; }
        __syncthreads();
        if (threadIdx.x == 0) {
            uint2 f3[3];
            for (uint32_t j = 0; j < 3 && j < n; j++) f3[j] = make_uint2(gr[j], gs[j]);
            write_rec(rec, b, c0, n, f3);
        }
        __syncthreads();
    }
}

struct ListsLayout {
    int shift;
    uint32_t reach;
    int64_t n_bins, cand_cap, nb, ntiles;
    iv_lrec_t* rec;
    uint32_t *diff, *nbig, *co, *big;
    long long *tD, *tO, *tC, *tOin;
    uint32_t *cand_se, *cand_row;
};

// ~2 bins per subject, at most 2^14 coordinates wide (entry offsets are 14-bit + a guard bit), coarser when long subjects would blow
// the lists up.  false = cannot be indexed this way.
static bool choose_lists(int64_t n, int64_t span, int64_t sum_len, int64_t reach_in, int* shift, int64_t* n_bins,
                         int64_t* cand_cap, uint32_t* reach) {
    if (span < 1) span = 1;
    if (sum_len < 0) sum_len = 0;
    const int64_t target = 2 * n > 4096 ? 2 * n : 4096;
    int s = 0;
    while (s < IV_L_MAX_SHIFT && (span >> s) > target) s++;
    while (s < IV_L_MAX_SHIFT && (sum_len >> s) > 8 * (n > 1024 ? n : 1024)) s++;
    // queries up to `reach` long are answered from the list of their first bin alone: subjects starting within reach of
    // a bin's origin are listed in the bin before as well.  At most one bin wide, and start offsets must stay below 2^14.
    int64_t r = reach_in > 0 ? reach_in : 0;
    if (r > ((int64_t)1 << s)) r = (int64_t)1 << s;
    if (r > (int64_t)IV_L_SMAX - 1 - ((int64_t)1 << s)) r = (int64_t)IV_L_SMAX - 1 - ((int64_t)1 << s);
    if (r < 0) r = 0;
    *reach = (uint32_t)r;
    *shift = s;
    *n_bins = (span >> s) + 1;
    *cand_cap = (sum_len >> s) + 3 * n + 16;  // a subject touches at most len / binwidth + 2 bins (+ 1 within reach)
    return *n_bins <= ((int64_t)1 << 26) && *cand_cap < ((int64_t)1 << 31) && n < ((int64_t)1 << 31);
}

static bool plan_lists(int64_t n, int64_t span, int64_t sum_len, int64_t reach, Bump& b, ListsLayout& L) {
    if (!choose_lists(n, span, sum_len, reach, &L.shift, &L.n_bins, &L.cand_cap, &L.reach)) return false;
    L.nb = L.n_bins + 2;  // bins 0 .. n_bins - 1, the bin behind the last one, a sentinel
    L.ntiles = cdiv(L.nb, LS_TILE);
    const size_t padded = (size_t)L.ntiles * LS_TILE;
    L.rec = b.take<iv_lrec_t>((size_t)L.nb);
    L.diff = b.take<uint32_t>(padded + 64);  // + the counter of long lists: one memset clears both
    L.nbig = L.diff + padded;
    L.co = b.take<uint32_t>(padded + 8);
    L.tD = b.take<long long>((size_t)L.ntiles);
    L.tO = b.take<long long>((size_t)L.ntiles);
    L.tC = b.take<long long>((size_t)L.ntiles);
    L.tOin = b.take<long long>((size_t)L.ntiles);
    L.big = b.take<uint32_t>((size_t)(L.cand_cap / (LS_SORT_MAX + 1) + 1));
    L.cand_se = b.take<uint32_t>((size_t)L.cand_cap + 16);  // + slack: the join reads whole 8-entry chunks
    L.cand_row = b.take<uint32_t>((size_t)L.cand_cap + 16);
    return true;
}

}  // namespace iv

using namespace iv;

extern "C" size_t iv_lists_workspace_bytes(int64_t n, int64_t total_span, int64_t sum_len) {
    Bump b(nullptr, 0);
    ListsLayout L;
    if (!plan_lists(n > 0 ? n : 0, total_span, sum_len, 0, b, L)) return 0;  // the layout does not depend on the reach
    return b.off + 256;
}

extern "C" int32_t iv_lists_shift(int64_t n, int64_t total_span, int64_t sum_len) {
    int s;
    int64_t nb, cc;
    uint32_t r;
    return choose_lists(n > 0 ? n : 0, total_span, sum_len, 0, &s, &nb, &cc, &r) ? s : -1;
}

extern "C" int iv_lists_build(const int64_t* starts, const int64_t* ends, int64_t n, const iv_groups_t* groups,
                              int64_t reach, void* ws, size_t ws_bytes, iv_lists_t* out, void* stream) {
    cudaStream_t st = (cudaStream_t)stream;
    IV_REQUIRE(groups && out && n >= 0, IV_ERR_INVALID, "iv_lists_build: null argument");
    IV_REQUIRE(ws && (((uintptr_t)ws) & 255) == 0, IV_ERR_INVALID, "iv_lists_build: workspace null or not 256-byte aligned");
    const iv_groups_t& G = *groups;
    const int64_t span = G.total_span > 0 ? G.total_span : 1;
    Bump b(ws, ws_bytes);
    ListsLayout L;
    IV_REQUIRE(plan_lists(n, span, G.sum_len, reach, b, L), IV_ERR_RANGE,
               "iv_lists_build: span / subject lengths beyond the list index (use iv_index_build)");
    IV_REQUIRE(b.fits(), IV_ERR_WORKSPACE, "iv_lists_build: workspace %zu < %zu", ws_bytes, b.off);
    memset(out, 0, sizeof(*out));
    out->n = n;
    out->total_span = span;
    out->shift = L.shift;
    out->reach = (int32_t)L.reach;
    out->n_bins = L.n_bins;
    out->cand_cap = L.cand_cap;
    out->rec = L.rec;
    out->cand_se = L.cand_se;
    out->cand_row = L.cand_row;
    out->groups = G;
    if (n == 0) {
        IV_CUDA(cudaMemsetAsync(L.rec, 0, (size_t)L.nb * sizeof(iv_lrec_t), st));
        return IV_OK;
    }
    const size_t padded = (size_t)L.ntiles * LS_TILE;
    IV_CUDA(cudaMemsetAsync(L.diff, 0, (padded + 64) * sizeof(uint32_t), st));
    const unsigned nbk = (unsigned)cdiv(n, LS_THREADS);
    k_lists_hist<<<nbk, LS_THREADS, 0, st>>>(starts, ends, n, G, L.shift, L.reach, L.diff);
    IV_LAUNCH_CHECK();
    k_lists_reduce<<<(unsigned)L.ntiles, LS_THREADS, 0, st>>>(L.diff, L.nb, L.tD, L.tO);
    IV_LAUNCH_CHECK();
    k_lists_scan_tiles<<<1, 1024, 0, st>>>(L.tD, L.tO, L.ntiles, L.nb, L.tC, L.tOin);
    IV_LAUNCH_CHECK();
    k_lists_apply<<<(unsigned)L.ntiles, LS_THREADS, 0, st>>>(L.diff, L.nb, L.tC, L.tOin, L.co);
    IV_LAUNCH_CHECK();
    k_lists_fill<<<nbk, LS_THREADS, 0, st>>>(starts, ends, n, G, L.shift, L.reach, L.co, L.diff, L.cand_cap, L.cand_se, L.cand_row);
    IV_LAUNCH_CHECK();
    k_lists_finalize<<<(unsigned)cdiv(L.nb, LS_THREADS), LS_THREADS, 0, st>>>(L.n_bins, L.co, L.cand_se, L.cand_row, L.rec, L.big,
                                                                          L.nbig);
    IV_LAUNCH_CHECK();
    k_lists_big<<<64, LS_THREADS, 0, st>>>(L.co, L.cand_se, L.cand_row, L.rec, L.big, L.nbig);
    IV_LAUNCH_CHECK();
    return IV_OK;
}
