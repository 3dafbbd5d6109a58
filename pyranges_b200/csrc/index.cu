// index.cu -- subject index of the binary operations (replaces NCLS construction,
// reference call sites pyranges/methods/join.py:12, intersection.py:19,71, coverage.py:20,58).
//
// All keys (chromosome/strand groups) are indexed at once in a FLATTENED coordinate:
//     x' = clamp(x, lo[g], hi[g]) - lo[g] + base[g]
// so that one global sorted array serves every group and a query never needs a per-group search.
// Pieces:  key_s/row_s  starts sorted (+ subject row),  end_s  ends in start order,  key_e[/row_e] ends
//          sorted;  soff/seoff/eoff  the same values as 32-bit offsets from their bin origin;
//          bins[]   one 32-byte record per coordinate bin (bin = x' >> shift): start rank, end rank and
//                   stabbing-list range at the bin's origin and at the next one (one L2 sector answers
//                   "how many starts / ends lie before x" up to a scan of the bin's own few entries);
//          cross[]  per-bin stabbing lists: the subjects that start before the bin and are still open at
//                   its origin.  A query starting in bin b overlaps, besides the starts inside
//                   [s, e), exactly the members of cross[b] that end after s and the bin's own earlier
//                   starts that end after s -- no scan over non-hits (the NCLS nesting replaced by
//                   direct addressing).  sum(len)/binwidth + n bounds the list storage.
#include <mutex>
#include <stdlib.h>

#include "common.cuh"

namespace iv {

constexpr int IX_THREADS = 256;

// ---- per-segment bounds (ingest-time metadata) ------------------------------------------------------------------
__global__ void k_bounds_init(int64_t* lo, int64_t* hi, int64_t* ml, int64_t* sl, int64_t* mn, int n) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) {
        lo[i] = INT64_MAX;
        hi[i] = INT64_MIN;
        ml[i] = 0;
        sl[i] = 0;
        if (mn) mn[i] = INT64_MAX;
    }
}

constexpr int BD_SPLIT = 64;
__global__ void __launch_bounds__(IX_THREADS)
k_bounds(const int64_t* __restrict__ s, const int64_t* __restrict__ e, const int64_t* __restrict__ seg_off,
         int64_t* lo, int64_t* hi, int64_t* ml, int64_t* sl, int64_t* minl) {
    __shared__ long long sm[5][IX_THREADS / 32];
    int seg = blockIdx.x;
    int64_t a = seg_off[seg], b = seg_off[seg + 1];
    int64_t len = b - a;
    int64_t chunk = (len + BD_SPLIT - 1) / BD_SPLIT;
    int64_t ca = a + chunk * blockIdx.y, cb = ca + chunk < b ? ca + chunk : b;
    long long mn = INT64_MAX, mx = INT64_MIN, m = 0, sum = 0, sh = INT64_MAX;
    for (int64_t i = ca + threadIdx.x; i < cb; i += IX_THREADS) {
        long long x = s[i], y = e[i];
        mn = x < mn ? x : mn;
        mx = y > mx ? y : mx;
        m = (y - x) > m ? (y - x) : m;
        sh = (y - x) < sh ? (y - x) : sh;
        sum += y > x ? y - x : 0;
    }
    mn = warp_min(mn);
    mx = warp_max(mx);
    m = warp_max(m);
    sh = warp_min(sh);
    sum = warp_sum(sum);
    int w = threadIdx.x >> 5;
    if (lane_id() == 0) { sm[0][w] = mn; sm[1][w] = mx; sm[2][w] = m; sm[3][w] = sum; sm[4][w] = sh; }
    __syncthreads();
    if (threadIdx.x == 0) {
        for (int k = 1; k < IX_THREADS / 32; k++) {
            mn = sm[0][k] < mn ? sm[0][k] : mn;
            mx = sm[1][k] > mx ? sm[1][k] : mx;
            m = sm[2][k] > m ? sm[2][k] : m;
            sh = sm[4][k] < sh ? sm[4][k] : sh;
            sum += sm[3][k];
        }
        if (ca < cb) {
            atomicMin((long long*)lo + seg, mn);
            atomicMax((long long*)hi + seg, mx);
            atomicMax((long long*)ml + seg, m);
            atomicAdd((unsigned long long*)sl + seg, (unsigned long long)sum);
            if (minl) atomicMin((long long*)minl + seg, sh);
        }
    }
}

// ---- flatten subjects ----------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(IX_THREADS)
k_flatten_subjects(const int64_t* __restrict__ s, const int64_t* __restrict__ e, int64_t n, iv_groups_t G,
                   uint64_t* __restrict__ key_s, uint32_t* __restrict__ row_s, uint64_t* __restrict__ key_e,
                   uint32_t* __restrict__ row_e) {  // row_e is always written (the batched sort carries it)
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

// Per start-sorted position p: exact flattened end, 32-bit bin offsets, and the rank fields of every bin
// whose origin falls between the previous key and this one (bs of bin b = first p with bin(key) >= b).
// blockIdx.y == 0: start order; == 1: end order.
__global__ void __launch_bounds__(IX_THREADS)
k_index_arrays(const int64_t* __restrict__ e, int64_t n, iv_groups_t G, const uint64_t* __restrict__ key_s,
               const uint32_t* __restrict__ row_s, const uint64_t* __restrict__ key_e, int shift, int64_t n_bins,
               uint64_t* __restrict__ end_s, uint32_t* __restrict__ soff, uint32_t* __restrict__ seoff,
               uint32_t* __restrict__ eoff, uint32_t* __restrict__ bsarr, uint32_t* __restrict__ bearr) {
    __shared__ int span[2];
    const int64_t first = (int64_t)blockIdx.x * IX_THREADS;
    const int64_t p = first + threadIdx.x;
    if (blockIdx.y == 0) {
        // sorted position p of group g lies in the same row range as the group's rows
        const int64_t last = first + IX_THREADS - 1 < n - 1 ? first + IX_THREADS - 1 : n - 1;
        SegSpan sp = block_seg_span(G.row_off, G.n_groups, first, last, span);
        if (p >= n) return;
        const int g = sp.of(G.row_off, p);
        const uint64_t k = key_s[p];
        const uint64_t E = (uint64_t)(e[__ldg(row_s + p)] + (__ldg(G.base + g) - __ldg(G.lo + g)));
        end_s[p] = E;
        const int64_t b = (int64_t)(k >> shift);
        const uint64_t origin = (uint64_t)b << shift;
        soff[p] = (uint32_t)(k - origin);
        const uint64_t d = E - origin;
        seoff[p] = d > 0xffffffffull ? 0xffffffffu : (uint32_t)d;
        const int64_t bprev = p ? (int64_t)(key_s[p - 1] >> shift) : -1;
        for (int64_t bb = bprev + 1; bb <= b; bb++) bsarr[bb] = (uint32_t)p;
        if (p == n - 1)
            for (int64_t bb = b + 1; bb <= n_bins + 1; bb++) bsarr[bb] = (uint32_t)n;
    } else {
        if (p >= n) return;
        const uint64_t k = key_e[p];
        const int64_t b = (int64_t)(k >> shift);
        eoff[p] = (uint32_t)(k - ((uint64_t)b << shift));
        const int64_t bprev = p ? (int64_t)(key_e[p - 1] >> shift) : -1;
        for (int64_t bb = bprev + 1; bb <= b; bb++) bearr[bb] = (uint32_t)p;
        if (p == n - 1)
            for (int64_t bb = b + 1; bb <= n_bins + 1; bb++) bearr[bb] = (uint32_t)n;
    }
}

// Stabbing list of bin b = the subjects with S' < origin(b) <= E'.  Its length needs no counting pass:
// #{S' < o} - #{E' < o} = bsarr[b] - bearr[b].  (A member with E' == o can never be a hit -- its eoff is 0 --
// and is filtered like every member that ends before the query starts.)  The list offsets are the exclusive scan of
// those lengths over the bins: per-tile sums here (the same pass clears the fill cursors), a single-block scan of
// the sums, and the final pass inside k_bins_finalize.
constexpr int BT_ITEMS = 8;
constexpr int BT_TILE = IX_THREADS * BT_ITEMS;  // bins per block

__global__ void __launch_bounds__(IX_THREADS)
k_cross_tile_sums(int64_t n_bins, const uint32_t* __restrict__ bsarr, const uint32_t* __restrict__ bearr,
                  uint32_t* __restrict__ cursor, uint32_t* __restrict__ partials) {
    __shared__ uint32_t sm[IX_THREADS / 32];
    const int64_t base = (int64_t)blockIdx.x * BT_TILE;
    uint32_t s = 0;
#pragma unroll
    for (int it = 0; it < BT_ITEMS; it++) {
        const int64_t b = base + it * IX_THREADS + threadIdx.x;
        if (b <= n_bins + 1) {
            cursor[b] = 0u;
            if (b <= n_bins) s += bsarr[b] - bearr[b];
        }
    }
    s = warp_sum(s);
    if (lane_id() == 0) sm[threadIdx.x >> 5] = s;
    __syncthreads();
    if (threadIdx.x < 32) {
        uint32_t x = threadIdx.x < IX_THREADS / 32 ? sm[threadIdx.x] : 0u;
        x = warp_sum(x);
        if (threadIdx.x == 0) partials[blockIdx.x] = x;
    }
}

// Fill: every (subject, crossed bin) pair is one unit of work, spread over the block's threads whatever the
// subject lengths (a 2 Mb gene crosses hundreds of bins, an exon none).  Entry order inside a list is arbitrary
// (atomic cursor).
__global__ void __launch_bounds__(IX_THREADS)
k_cross_fill(int64_t n, const uint64_t* __restrict__ key_s, const uint64_t* __restrict__ end_s,
             const uint32_t* __restrict__ row_s, int shift, const iv_binrec_t* __restrict__ bins,
             uint32_t* __restrict__ cursor, int64_t cap, iv_cross_t* __restrict__ cross, uint32_t* __restrict__ cross_p) {
    __shared__ uint32_t pre[IX_THREADS + 1];
    __shared__ uint32_t wsum[IX_THREADS / 32 + 1];
    __shared__ uint64_t sE[IX_THREADS];
    __shared__ int64_t sB[IX_THREADS];
    __shared__ uint32_t sRow[IX_THREADS];
    const int64_t p = (int64_t)blockIdx.x * IX_THREADS + threadIdx.x;
    uint32_t cnt = 0;
    if (p < n) {
        const uint64_t k = key_s[p], E = end_s[p];
        const int64_t b = (int64_t)(k >> shift);
        sE[threadIdx.x] = E;
        sB[threadIdx.x] = b + 1;
        sRow[threadIdx.x] = __ldg(row_s + p);
        if (E > k) cnt = (uint32_t)((int64_t)(E >> shift) - b);
    }
    uint32_t tot;
    pre[threadIdx.x] = block_excl_sum<uint32_t, IX_THREADS>(cnt, wsum, tot);
    if (threadIdx.x == 0) pre[IX_THREADS] = tot;
    __syncthreads();
    for (uint32_t u = threadIdx.x; u < tot; u += IX_THREADS) {
        int lo = 0, hi = IX_THREADS;  // owner = last thread whose prefix <= u
        while (hi - lo > 1) {
            int mid = (lo + hi) >> 1;
            if (pre[mid] <= u) lo = mid; else hi = mid;
        }
        const int64_t bb = sB[lo] + (u - pre[lo]);
        const uint64_t d = sE[lo] - ((uint64_t)bb << shift);
        const int64_t at = (int64_t)__ldg(&bins[bb].co) + atomicAdd(cursor + bb, 1u);
        if (at < cap) {
            iv_cross_t c;
            c.row = sRow[lo];
            c.eoff = d > 0xffffffffull ? 0xffffffffu : (uint32_t)d;
            cross[at] = c;
            cross_p[at] = (uint32_t)(blockIdx.x * IX_THREADS + lo);
        }
    }
}

// assemble the 32-byte bin records: ranks at the bin origin, the stabbing-list offset (block scan of the list lengths
// on top of the tile's scanned partial) and the first offsets of each bin.  BT_ITEMS consecutive bins per thread.
__global__ void __launch_bounds__(IX_THREADS)
k_bins_finalize(int64_t n_bins, int shift, const uint32_t* __restrict__ bsarr, const uint32_t* __restrict__ bearr,
                const uint32_t* __restrict__ partials, const uint32_t* __restrict__ soff,
                const uint32_t* __restrict__ eoff, iv_binrec_t* __restrict__ bins) {
    __shared__ uint32_t wsum[IX_THREADS / 32 + 1];
    const int64_t b0 = (int64_t)blockIdx.x * BT_TILE + (int64_t)threadIdx.x * BT_ITEMS;
    uint32_t bs[BT_ITEMS + 1], be[BT_ITEMS + 1];
#pragma unroll
    for (int j = 0; j <= BT_ITEMS; j++) {
        const int64_t b = b0 + j;
        bs[j] = b <= n_bins + 1 ? bsarr[b] : 0u;
        be[j] = b <= n_bins + 1 ? bearr[b] : 0u;
    }
    uint32_t mine = 0;
#pragma unroll
    for (int j = 0; j < BT_ITEMS; j++)
        if (b0 + j <= n_bins) mine += bs[j] - be[j];
    uint32_t tot;
    uint32_t co = __ldg(partials + blockIdx.x) + block_excl_sum<uint32_t, IX_THREADS>(mine, wsum, tot);
    const bool inl = shift <= 15;  // inline offsets are 15-bit (SWAR compares in the count kernel)
#pragma unroll
    for (int j = 0; j < BT_ITEMS; j++) {
        const int64_t b = b0 + j;
        if (b > n_bins + 1) break;
        // bin n_bins + 1 is the sentinel a look-up of "the next bin" may touch: no members, list offset = total
        const uint32_t ns = b <= n_bins ? bs[j + 1] - bs[j] : 0u, ne = b <= n_bins ? be[j + 1] - be[j] : 0u;
        const uint32_t nc = b <= n_bins ? bs[j] - be[j] : 0u;
        uint32_t so[4], eo[4];
#pragma unroll
        for (int i = 0; i < 4; i++) {
            so[i] = (inl && (uint32_t)i < ns) ? __ldg(soff + bs[j] + i) : 0x7fffu;
            eo[i] = (inl && (uint32_t)i < ne) ? __ldg(eoff + be[j] + i) : 0x7fffu;
        }
        uint4 a, c;
        a.x = bs[j]; a.y = be[j]; a.z = co;
        a.w = (ns < 255 ? ns : 255u) | ((ne < 255 ? ne : 255u) << 8) | ((nc < 255 ? nc : 255u) << 16) |
              ((inl ? (uint32_t)IV_BIN_INLINE | ((ns <= 4 && ne <= 4) ? (uint32_t)IV_BIN_FAST : 0u) : 0u) << 24);
        c.x = so[0] | (so[1] << 16); c.y = so[2] | (so[3] << 16);
        c.z = eo[0] | (eo[1] << 16); c.w = eo[2] | (eo[3] << 16);
        uint4* out = reinterpret_cast<uint4*>(bins + b);
        out[0] = a;
        out[1] = c;
        co += nc;
    }
}

struct IndexLayout {
    uint64_t *ks0, *ks1, *ke0, *ke1, *end_s;
    uint32_t *rs0, *rs1, *re0, *re1, *soff, *seoff, *eoff, *bsarr, *bearr, *cursor, *cross_p, *scan_tmp;
    iv_binrec_t* bins;
    iv_cross_t* cross;
    void* sort_temp;
    int shift;
    int64_t n_bins, cross_cap;
};

// about two bins per subject (2^10 .. 2^26 bins), coarser when long subjects would blow the stabbing lists up
static void choose_bins(int64_t n, int64_t span, int64_t sum_len, int* shift, int64_t* n_bins, int64_t* cross_cap) {
    int64_t target = 1024;
    while (target < 4 * n && target < ((int64_t)1 << 26)) target <<= 1;
    int s = 0;
    while ((span >> s) > target) s++;
    if (sum_len < 0) sum_len = 0;
    while (s < 31 && (sum_len >> s) > 6 * (n > 1024 ? n : 1024)) s++;
    if (s > 31) s = 31;  // bin offsets are 32-bit
    *shift = s;
    *n_bins = (span >> s) + 1;
    *cross_cap = (sum_len >> s) + 2 * n + 16;  // bin(E) - bin(S) <= len / binwidth + 1 per subject
}

static void plan_index(int64_t n, int64_t span, int64_t sum_len, int flags, Bump& b, IndexLayout& L) {
    size_t nn = (size_t)(n > 0 ? n : 1);
    choose_bins(n, span, sum_len, &L.shift, &L.n_bins, &L.cross_cap);
    L.ks0 = b.take<uint64_t>(nn);
    L.ks1 = b.take<uint64_t>(nn);
    L.rs0 = b.take<uint32_t>(nn);
    L.rs1 = b.take<uint32_t>(nn);
    L.ke0 = b.take<uint64_t>(nn);
    L.ke1 = b.take<uint64_t>(nn);
    L.re0 = b.take<uint32_t>(nn);
    L.re1 = b.take<uint32_t>(nn);
    L.end_s = b.take<uint64_t>(nn);
    L.soff = b.take<uint32_t>(nn);
    L.seoff = b.take<uint32_t>(nn);
    L.eoff = b.take<uint32_t>(nn);
    L.bins = b.take<iv_binrec_t>((size_t)L.n_bins + 2);
    L.bsarr = b.take<uint32_t>((size_t)L.n_bins + 4);
    L.bearr = b.take<uint32_t>((size_t)L.n_bins + 4);
    L.cursor = b.take<uint32_t>((size_t)L.n_bins + 4);
    {
        const int64_t nbt = cdiv(L.n_bins + 2, BT_TILE);
        L.scan_tmp = b.take<uint32_t>((size_t)((nbt + 63) & ~(int64_t)63) + scan_u32_large_temp_elems(nbt) + 64);
    }
    L.cross = b.take<iv_cross_t>((size_t)L.cross_cap);
    L.cross_p = b.take<uint32_t>((size_t)L.cross_cap);
    L.sort_temp = b.take<char>(radix_sort_temp_bytes(n));
}

}  // namespace iv

using namespace iv;

extern "C" int iv_segment_bounds(const int64_t* starts, const int64_t* ends, const int64_t* seg_off, int32_t n_seg,
                                 int64_t* out_lo, int64_t* out_hi, int64_t* out_maxlen, int64_t* out_sumlen,
                                 int64_t* out_minlen, void* stream) {
    cudaStream_t st = (cudaStream_t)stream;
    IV_REQUIRE(n_seg >= 0, IV_ERR_INVALID, "iv_segment_bounds: n_seg < 0");
    if (n_seg == 0) return IV_OK;
    IV_REQUIRE(out_lo && out_hi && out_maxlen && out_sumlen, IV_ERR_INVALID, "iv_segment_bounds: null output");
    k_bounds_init<<<(n_seg + 255) / 256, 256, 0, st>>>(out_lo, out_hi, out_maxlen, out_sumlen, out_minlen, n_seg);
    IV_LAUNCH_CHECK();
    k_bounds<<<dim3((unsigned)n_seg, BD_SPLIT), IX_THREADS, 0, st>>>(starts, ends, seg_off, out_lo, out_hi, out_maxlen,
                                                                      out_sumlen, out_minlen);
    IV_LAUNCH_CHECK();
    return IV_OK;
}

extern "C" size_t iv_index_workspace_bytes(int64_t n, int64_t total_span, int64_t sum_len, int32_t flags) {
    Bump b(nullptr, 0);
    IndexLayout L;
    plan_index(n, total_span > 0 ? total_span : 1, sum_len, flags, b, L);
    return b.off + 256;
}

// The launch sequence of a build (~20 small kernels) is fixed by its arguments.  pyranges rebuilds the index of
// the same columns on every call (pyranges/methods/join.py:12), so the second identical call captures the
// sequence into a CUDA graph and later ones replay it with a single launch: the host enqueue cost of the build
// disappears and the kernels run back to back.
namespace iv {
struct BuildKey {
    const void *starts, *ends, *ws, *row_off, *lo, *hi, *base;
    void* stream;
    int64_t n, span, sum_len, max_len;
    size_t ws_bytes;
    int flags, n_groups;
    bool operator==(const BuildKey& o) const { return memcmp(this, &o, sizeof(BuildKey)) == 0; }
};
struct BuildEntry {
    BuildKey key;
    cudaGraphExec_t exec;
    iv_index_t out;
    unsigned long long launches;
    int seen;
};
constexpr int BUILD_CACHE = 8;
static BuildEntry g_builds[BUILD_CACHE];
static int g_build_next = 0;
static std::mutex g_build_mu;
}  // namespace iv

static int index_build_direct(const int64_t* starts, const int64_t* ends, int64_t n, const iv_groups_t* groups,
                              int32_t flags, void* ws, size_t ws_bytes, iv_index_t* out, void* stream);

extern "C" int iv_index_build(const int64_t* starts, const int64_t* ends, int64_t n, const iv_groups_t* groups,
                              int32_t flags, void* ws, size_t ws_bytes, iv_index_t* out, void* stream) {
    cudaStream_t st = (cudaStream_t)stream;
    IV_REQUIRE(groups && out && n >= 0, IV_ERR_INVALID, "iv_index_build: null argument");
    // replay of a captured launch sequence is OPT-IN (IV_INDEX_GRAPH); IV_NO_GRAPH in the environment overrides it
    static const bool no_graphs = getenv("IV_NO_GRAPH") != nullptr;
    if (!(flags & IV_INDEX_GRAPH) || no_graphs || n == 0 || st == nullptr)
        return index_build_direct(starts, ends, n, groups, flags, ws, ws_bytes, out, stream);
    BuildKey key;
    memset(&key, 0, sizeof(key));
    key.starts = starts; key.ends = ends; key.ws = ws; key.row_off = groups->row_off; key.lo = groups->lo;
    key.hi = groups->hi; key.base = groups->base; key.stream = stream; key.n = n; key.span = groups->total_span;
    key.sum_len = groups->sum_len; key.max_len = groups->max_len; key.ws_bytes = ws_bytes; key.flags = flags;
    key.n_groups = groups->n_groups;
    std::lock_guard<std::mutex> lock(g_build_mu);
    BuildEntry* e = nullptr;
    for (int i = 0; i < BUILD_CACHE; i++)
        if (g_builds[i].seen && g_builds[i].key == key) e = &g_builds[i];
    if (e && e->exec) {  // replay
        IV_CUDA(cudaGraphLaunch(e->exec, st));
        __atomic_fetch_add(&iv::g_launches, e->launches, __ATOMIC_RELAXED);
        *out = e->out;
        return IV_OK;
    }
    if (!e) {  // first sighting: plain launches, remember the arguments
        e = &g_builds[g_build_next];
        g_build_next = (g_build_next + 1) % BUILD_CACHE;
        if (e->exec) cudaGraphExecDestroy(e->exec);
        memset(e, 0, sizeof(*e));
        e->key = key;
        e->seen = 1;
        return index_build_direct(starts, ends, n, groups, flags, ws, ws_bytes, out, stream);
    }
    // second sighting: capture, instantiate, launch
    const unsigned long long before = __atomic_load_n(&iv::g_launches, __ATOMIC_RELAXED);
    if (cudaStreamBeginCapture(st, cudaStreamCaptureModeThreadLocal) != cudaSuccess) {
        cudaGetLastError();
        return index_build_direct(starts, ends, n, groups, flags, ws, ws_bytes, out, stream);
    }
    int rc = index_build_direct(starts, ends, n, groups, flags, ws, ws_bytes, out, stream);
    cudaGraph_t graph = nullptr;
    cudaError_t ce = cudaStreamEndCapture(st, &graph);
    const unsigned long long captured = __atomic_load_n(&iv::g_launches, __ATOMIC_RELAXED) - before;
    __atomic_fetch_sub(&iv::g_launches, captured, __ATOMIC_RELAXED);  // captured, not yet run
    if (rc != IV_OK || ce != cudaSuccess || !graph) {
        cudaGetLastError();
        if (graph) cudaGraphDestroy(graph);
        e->seen = 0;
        return index_build_direct(starts, ends, n, groups, flags, ws, ws_bytes, out, stream);
    }
    cudaGraphExec_t exec = nullptr;
    ce = cudaGraphInstantiate(&exec, graph, 0);
    cudaGraphDestroy(graph);
    if (ce != cudaSuccess || !exec) {
        cudaGetLastError();
        e->seen = 0;
        return index_build_direct(starts, ends, n, groups, flags, ws, ws_bytes, out, stream);
    }
    e->exec = exec;
    e->out = *out;
    e->launches = captured;
    IV_CUDA(cudaGraphLaunch(exec, st));
    __atomic_fetch_add(&iv::g_launches, captured, __ATOMIC_RELAXED);
    return IV_OK;
}

static int index_build_direct(const int64_t* starts, const int64_t* ends, int64_t n, const iv_groups_t* groups,
                              int32_t flags, void* ws, size_t ws_bytes, iv_index_t* out, void* stream) {
    cudaStream_t st = (cudaStream_t)stream;
    IV_REQUIRE(groups && out && n >= 0, IV_ERR_INVALID, "iv_index_build: null argument");
    IV_REQUIRE(n < ((int64_t)1 << 31), IV_ERR_RANGE, "iv_index_build: more than 2^31-1 subjects");
    const iv_groups_t& G = *groups;
    int64_t span = G.total_span > 0 ? G.total_span : 1;
    IV_REQUIRE(span < ((int64_t)1 << 56), IV_ERR_RANGE, "iv_index_build: flattened span too large");
    Bump b(ws, ws_bytes);
    IndexLayout L;
    plan_index(n, span, G.sum_len, flags, b, L);
    IV_REQUIRE(b.fits(), IV_ERR_WORKSPACE, "iv_index_build: workspace %zu < %zu", ws_bytes, b.off);
    IV_REQUIRE(L.cross_cap < ((int64_t)1 << 32), IV_ERR_RANGE, "iv_index_build: stabbing lists exceed 2^32 entries");
    IV_REQUIRE(L.n_bins <= ((int64_t)1 << 28), IV_ERR_RANGE, "iv_index_build: flattened span needs too many bins");

    memset(out, 0, sizeof(*out));
    out->n = n;
    out->total_span = span;
    out->shift = L.shift;
    out->flags = flags;
    out->n_bins = L.n_bins;
    out->cross_cap = L.cross_cap;
    out->groups = G;
    out->key_s = L.ks0; out->row_s = L.rs0; out->key_e = L.ke0; out->row_e = L.re0;
    out->end_s = L.end_s; out->soff = L.soff; out->seoff = L.seoff; out->eoff = L.eoff;
    out->bins = L.bins; out->cross = L.cross; out->cross_p = L.cross_p;
    if (n == 0) {
        IV_CUDA(cudaMemsetAsync(L.bins, 0, ((size_t)L.n_bins + 2) * sizeof(iv_binrec_t), st));
        return IV_OK;
    }
    unsigned nb = (unsigned)cdiv(n, IX_THREADS);
    k_flatten_subjects<<<nb, IX_THREADS, 0, st>>>(starts, ends, n, G, L.ks0, L.rs0, L.ke0, L.re0);
    IV_LAUNCH_CHECK();
    int bits = bits_for((uint64_t)span);
    int res = 0;
    SortJob jobs[2] = {{L.ks0, L.ks1, L.rs0, L.rs1, 4},
                       {L.ke0, L.ke1, L.re0, L.re1, (flags & IV_INDEX_ENDS_PERM) ? 4 : 0}};
    int rc = radix_sort_multi(jobs, 2, n, 0, bits, L.sort_temp, radix_sort_temp_bytes(n), st, &res);
    if (rc != IV_OK) return rc;
    if (res) { out->key_s = L.ks1; out->row_s = L.rs1; out->key_e = L.ke1; out->row_e = L.re1; }
    if (!(flags & IV_INDEX_ENDS_PERM)) out->row_e = nullptr;
    k_index_arrays<<<dim3(nb, 2), IX_THREADS, 0, st>>>(ends, n, G, out->key_s, out->row_s, out->key_e, L.shift, L.n_bins,
                                                       L.end_s, L.soff, L.seoff, L.eoff, L.bsarr, L.bearr);
    IV_LAUNCH_CHECK();
    // stabbing-list offsets: tile sums of the list lengths -> single-block scan -> finished inside k_bins_finalize
    const int64_t nbt = cdiv(L.n_bins + 2, BT_TILE);
    k_cross_tile_sums<<<(unsigned)nbt, IX_THREADS, 0, st>>>(L.n_bins, L.bsarr, L.bearr, L.cursor, L.scan_tmp);
    IV_LAUNCH_CHECK();
    rc = scan_exclusive_u32_large(L.scan_tmp, nbt, L.scan_tmp + ((nbt + 63) & ~(int64_t)63), st);
    if (rc != IV_OK) return rc;
    k_bins_finalize<<<(unsigned)nbt, IX_THREADS, 0, st>>>(L.n_bins, L.shift, L.bsarr, L.bearr, L.scan_tmp, L.soff, L.eoff,
                                                         L.bins);
    IV_LAUNCH_CHECK();
    k_cross_fill<<<nb, IX_THREADS, 0, st>>>(n, out->key_s, L.end_s, out->row_s, L.shift, L.bins, L.cursor, L.cross_cap,
                                            L.cross, L.cross_p);
    IV_LAUNCH_CHECK();
    return IV_OK;
}
