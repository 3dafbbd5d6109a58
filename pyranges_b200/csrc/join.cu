// join.cu -- the fused join: count, output offsets and emission of all overlap pairs in ONE kernel over the list
// index (lists.cu).  Replaces NCLS.all_overlaps_both (pyranges/methods/join.py:15,23) and, in its count-only form,
// all_overlaps_both + value_counts / has_overlaps (pyranges/methods/coverage.py:26-40, intersection.py:78).
//
// A block owns a tile of 2048 consecutive query rows (tiles are handed out by an atomic ticket, so a tile only ever
// waits for tiles that are already running), a warp 256 consecutive rows of it, a lane two rows per iteration:
//   1. 16-byte streaming loads of Start / End, 32-bit clamp into the partner group's window of the flattened axis,
//      ONE 256-bit gather of the record of the bin the query starts in (list offset, length, first three entries);
//   2. count sweep: two 16-bit compares per entry (inline entries from registers; longer lists and the later bins of a
//      query that reaches beyond its first bin are walked in global memory, L2-resident);
//   3. shuffle scan of the counts -> the lane's position in the warp's STAGING area in shared memory; emit sweep: the
//      same compares again, hits go to the staging area as (row in tile: 16 bit, subject row: 32 bit);
//   4. the tile's pair total is published and its output offset obtained by a decoupled look-back over the preceding
//      tiles (status word = 62-bit value + 2-bit flag, one relaxed 64-bit access each);
//   5. every warp copies its staged pairs to out_q / out_s with fully coalesced 8-byte streaming stores.
// A warp whose pairs do not fit its staging area repeats its rows once the offset is known and writes directly.
// DRAM sees the query columns once and the pair columns once; the index stays in L2.
#include "common.cuh"
#include "query.cuh"

namespace iv {

constexpr int JN_THREADS = 256;
constexpr int JN_WARPS = JN_THREADS / 32;
constexpr int JN_ITERS = 4;                       // iterations of 64 rows per warp
constexpr int JN_WROWS = 64 * JN_ITERS;           // 256 rows per warp
constexpr int JN_TILE = JN_WROWS * JN_WARPS;      // 2048 rows per block
constexpr int JN_WCAP = 768;                      // staged pairs per warp (3 per row; the rest goes the direct way)

constexpr int JM_STAGE = 0, JM_DIRECT = 1, JM_COUNT = 2, JM_ANY = 3;

struct LRec { uint32_t w[8]; };  // w0 = co, w1 = n8, (w2, w3) (w4, w5) (w6, w7) = entries {row, so | eo << 16}

__device__ __forceinline__ LRec load_lrec(const iv_lrec_t* __restrict__ rec, uint64_t b) {
    LRec r;
    asm("ld.global.nc.v8.u32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
        : "=r"(r.w[0]), "=r"(r.w[1]), "=r"(r.w[2]), "=r"(r.w[3]), "=r"(r.w[4]), "=r"(r.w[5]), "=r"(r.w[6]), "=r"(r.w[7])
        : "l"(rec + b));
    return r;
}

// FIRST bin of a query (the one its start falls into): hit <=> (crossing or so < rel_e) and eo > rel_s
// later bins: only subjects that START there: hit <=> not crossing and so < rel_e
template <bool FIRST>
__device__ __forceinline__ bool entry_hit(uint32_t se, uint32_t rel_s, uint32_t rel_e) {
    const uint32_t so = se & 0xffffu, eo = se >> 16;
    if (FIRST) return ((so & IV_L_CROSS) || so < rel_e) && eo > rel_s;
    return so < rel_e;  // a crossing entry has so >= 0x8000 > any rel_e of a later bin (<= 2^15); an open-ended
                        // rel_e (0xffffffff) is never used for later bins
}

// visit the hits of one bin's list
template <bool FIRST, typename F>
__device__ __forceinline__ void walk_bin(const iv_lists_t& ix, uint64_t b, const LRec& r, uint32_t rel_s, uint32_t rel_e, F&& f) {
    uint32_t n = r.w[1] & 0xffu;
    if (n == 255u) n = __ldg(&ix.rec[b + 1].co) - r.w[0];
    if (n > 0 && entry_hit<FIRST>(r.w[3], rel_s, rel_e)) f(r.w[2]);
    if (n > 1 && entry_hit<FIRST>(r.w[5], rel_s, rel_e)) f(r.w[4]);
    if (n > 2 && entry_hit<FIRST>(r.w[7], rel_s, rel_e)) f(r.w[6]);
    if (n > 3) {
        const uint2* __restrict__ c = reinterpret_cast<const uint2*>(ix.cand) + r.w[0];
#pragma unroll 4
        for (uint32_t j = 3; j < n; j++) {
            const uint2 en = __ldg(c + j);
            if (entry_hit<FIRST>(en.y, rel_s, rel_e)) f(en.x);
        }
    }
}

// all hits of a flattened query [sc, ec): r = record of bin(sc)
template <typename X, typename F>
__device__ __forceinline__ void walk_query(const iv_lists_t& ix, X sc, X ec, const LRec& r, F&& f) {
    const int shift = ix.shift;
    const uint64_t b = (uint64_t)sc >> shift;
    const uint32_t rel_s = (uint32_t)(sc - (X)(b << shift));
    const uint64_t d = (uint64_t)(ec - (X)(b << shift));  // ec >= sc: no wrap
    const bool single = d <= (1ull << shift);
    walk_bin<true>(ix, b, r, rel_s, single ? (uint32_t)d : 0x7fffffffu, f);
    if (!single) {
        const uint64_t bl = (uint64_t)(ec - 1) >> shift;
        for (uint64_t bb = b + 1; bb <= bl; bb++) {
            const LRec rr = load_lrec(ix.rec, bb);
            const uint32_t re = bb < bl ? 0x8000u : (uint32_t)(ec - (X)(bl << shift));
            walk_bin<false>(ix, bb, rr, 0u, re, f);
        }
    }
}

// ---- decoupled look-back over the tile status words -------------------------------------------------------------------------
__device__ __forceinline__ unsigned long long ld_relaxed(const unsigned long long* p) {
    unsigned long long v;
    asm volatile("ld.relaxed.gpu.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void st_relaxed(unsigned long long* p, unsigned long long v) {
    asm volatile("st.relaxed.gpu.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
// called by one full warp; -> pairs of all tiles before `tile`
__device__ __forceinline__ unsigned long long lookback(unsigned long long* state, int64_t tile, unsigned long long aggregate) {
    const int lane = lane_id();
    if (lane == 0) st_relaxed(state + tile, (aggregate << 2) | (tile == 0 ? 2ull : 1ull));
    unsigned long long excl = 0;
    int64_t look = tile - 1;
    while (look >= 0) {
        const int64_t idx = look - lane;
        unsigned long long v = 2ull;  // before the first tile: an inclusive prefix of 0
        if (idx >= 0) {
            v = ld_relaxed(state + idx);
            while ((v & 3ull) == 0ull) { __nanosleep(20); v = ld_relaxed(state + idx); }
        }
        const unsigned incl = __ballot_sync(FULL, (v & 3ull) == 2ull);
        const int first = incl ? __ffs(incl) - 1 : 32;  // nearest predecessor that already knows its inclusive prefix
        excl += warp_sum(lane <= first ? (v >> 2) : 0ull);
        if (incl) break;
        look -= 32;
    }
    if (lane == 0 && tile > 0) st_relaxed(state + tile, ((excl + aggregate) << 2) | 2ull);
    return excl;
}

// ---- the rows of one warp ----------------------------------------------------------------------------------------------------
struct WarpOut {
    unsigned long long total;  // pairs of the warp's rows
    bool overflow;             // they did not fit the staging area (JM_STAGE)
};

template <typename X, int MODE, bool SLACK>
__device__ __forceinline__ WarpOut warp_rows(const iv_lists_t& ix, const iv_queries_t& q, int64_t wbase, bool vec,
                                             uint32_t* __restrict__ st_s, uint16_t* __restrict__ st_q, uint32_t qloc0,
                                             int64_t* __restrict__ out_q, int64_t* __restrict__ out_s,
                                             const int64_t* __restrict__ s_label, int64_t cap, unsigned long long obase,
                                             int64_t* __restrict__ out_count) {
    const int lane = lane_id();
    const int shift = ix.shift;
    WarpOut wo{0ull, false};
    if (wbase >= q.n) return wo;
    int seg = find_seg(q.seg_off, q.n_seg, wbase), f_seg = -1;
    Flat<X> f0;
    uint32_t lo32 = 0, hi32 = 0, off32 = 0;
    bool geo32 = false;
#pragma unroll 1
    for (int it = 0; it < JN_ITERS; it++) {
        const int64_t ibase = wbase + it * 64;
        if (ibase >= q.n) break;
        while (seg + 1 < q.n_seg && __ldg(q.seg_off + seg + 1) <= ibase) seg++;
        const bool uniform = __ldg(q.seg_off + seg + 1) >= ibase + 64;
        if (seg != f_seg) {
            f0 = load_flat_groups<X>(ix.groups, ix.n, q, seg);
            f_seg = seg;
            const uint64_t hi = (uint64_t)f0.lo + f0.width;
            geo32 = sizeof(X) == 4 && f0.lo >= 0 && hi <= 0xffffffffull;
            lo32 = (uint32_t)f0.lo;
            hi32 = (uint32_t)hi;
            off32 = (uint32_t)f0.base - lo32;
        }
        const int64_t r0 = ibase + 2 * lane;  // rows r0, r0 + 1
        int64_t s[2] = {0, 0}, e[2] = {0, 0};
        if (vec && r0 + 1 < q.n) {
            const longlong2 s2 = __ldcs(reinterpret_cast<const longlong2*>(q.start + r0));
            const longlong2 e2 = __ldcs(reinterpret_cast<const longlong2*>(q.end + r0));
            s[0] = s2.x; s[1] = s2.y; e[0] = e2.x; e[1] = e2.y;
        } else {
            if (r0 < q.n) { s[0] = __ldcs(q.start + r0); e[0] = __ldcs(q.end + r0); }
            if (r0 + 1 < q.n) { s[1] = __ldcs(q.start + r0 + 1); e[1] = __ldcs(q.end + r0 + 1); }
        }
        X sc[2], ec[2];
        LRec r[2];
        bool live[2];
#pragma unroll
        for (int k = 0; k < 2; k++) {
            if (SLACK) apply_slack(q.slack, s[k], e[k]);
            bool partner;
            if (uniform && geo32 && (((uint64_t)s[k] | (uint64_t)e[k]) >> 32) == 0) {
                sc[k] = (X)(min(max((uint32_t)s[k], lo32), hi32) + off32);
                ec[k] = (X)(min(max((uint32_t)e[k], lo32), hi32) + off32);
                partner = f0.g >= 0;
            } else {
                Flat<X> f = f0;
                if (!uniform) {
                    const int sg = find_seg_range(q.seg_off, seg, q.n_seg - 1, r0 + k < q.n ? r0 + k : q.n - 1);
                    if (sg != seg) f = load_flat_groups<X>(ix.groups, ix.n, q, sg);
                }
                sc[k] = f.at(s[k]);
                ec[k] = f.at(e[k]);
                partner = f.g >= 0;
            }
            // rows beyond the end, keys without a partner and malformed rows (end < start) have no hits
            live[k] = r0 + k < q.n && partner && ec[k] >= sc[k];
            if (!live[k]) { sc[k] = 0; ec[k] = 0; }
            r[k] = load_lrec(ix.rec, (uint64_t)sc[k] >> shift);
        }
        // count sweep
        uint32_t c[2];
#pragma unroll
        for (int k = 0; k < 2; k++) {
            uint32_t n = 0;
            if (live[k]) walk_query<X>(ix, sc[k], ec[k], r[k], [&](uint32_t) { n++; });
            c[k] = MODE == JM_ANY ? (n > 0 ? 1u : 0u) : n;
        }
        if (MODE == JM_COUNT || MODE == JM_ANY) {
            if (vec && r0 + 1 < q.n) {
                __stcs(reinterpret_cast<longlong2*>(out_count + r0), make_longlong2((long long)c[0], (long long)c[1]));
            } else {
                if (r0 < q.n) out_count[r0] = c[0];
                if (r0 + 1 < q.n) out_count[r0 + 1] = c[1];
            }
            continue;
        }
        // position of the lane's pairs among the warp's
        const uint32_t mine = c[0] + c[1];
        const uint32_t inc = warp_incl_sum(mine);
        const uint32_t itot = __shfl_sync(FULL, inc, 31);
        if (itot == 0) continue;
        if (MODE == JM_STAGE) {
            if (!wo.overflow && wo.total + itot <= (unsigned long long)JN_WCAP) {
                uint32_t pos = (uint32_t)wo.total + inc - mine;
#pragma unroll
                for (int k = 0; k < 2; k++) {
                    if (c[k]) {
                        const uint16_t ql = (uint16_t)(qloc0 + it * 64 + 2 * lane + k);
                        walk_query<X>(ix, sc[k], ec[k], r[k], [&](uint32_t row) {
                            st_s[pos] = row;
                            st_q[pos] = ql;
                            pos++;
                        });
                    }
                }
            } else {
                wo.overflow = true;
            }
        } else {  // JM_DIRECT: the output offset of the warp is known
            unsigned long long o = obase + wo.total + inc - mine;
#pragma unroll
            for (int k = 0; k < 2; k++) {
                if (c[k]) {
                    const int64_t row_q = r0 + k;
                    const int64_t ql = q.label ? __ldg(q.label + row_q) : row_q;
                    walk_query<X>(ix, sc[k], ec[k], r[k], [&](uint32_t row) {
                        if ((int64_t)o < cap) {
                            if (out_q) out_q[o] = ql;
                            if (out_s) out_s[o] = s_label ? __ldg(s_label + row) : (int64_t)row;
                        }
                        o++;
                    });
                }
            }
        }
        wo.total += itot;
    }
    return wo;
}

template <typename X, bool SLACK>
__global__ void __launch_bounds__(JN_THREADS, 4)
k_join(iv_lists_t ix, iv_queries_t q, bool vec, int64_t* __restrict__ out_q, int64_t* __restrict__ out_s,
       const int64_t* __restrict__ s_label, int64_t cap, unsigned long long* __restrict__ state,
       unsigned int* __restrict__ ticket, int64_t n_tiles, int64_t* __restrict__ out_total) {
    __shared__ uint32_t st_s[JN_WARPS][JN_WCAP];
    __shared__ uint16_t st_q[JN_WARPS][JN_WCAP];
    __shared__ unsigned long long s_wtot[JN_WARPS];
    __shared__ unsigned long long s_base;
    __shared__ unsigned int s_tile;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (threadIdx.x == 0) s_tile = atomicAdd(ticket, 1u);
    __syncthreads();
    const int64_t tile = s_tile;
    const int64_t tbase = tile * JN_TILE;
    const int64_t wbase = tbase + (int64_t)warp * JN_WROWS;
    const WarpOut wo = warp_rows<X, JM_STAGE, SLACK>(ix, q, wbase, vec, st_s[warp], st_q[warp], (uint32_t)(warp * JN_WROWS),
                                                     nullptr, nullptr, nullptr, 0, 0ull, nullptr);
    if (lane == 0) s_wtot[warp] = wo.total;
    __syncthreads();
    if (warp == 0) {
        unsigned long long agg = 0;
#pragma unroll
        for (int w = 0; w < JN_WARPS; w++) agg += s_wtot[w];
        const unsigned long long excl = lookback(state, tile, agg);
        if (lane == 0) {
            s_base = excl;
            if (tile == n_tiles - 1) *out_total = (int64_t)(excl + agg);
        }
    }
    __syncthreads();
    unsigned long long obase = s_base;
    for (int w = 0; w < warp; w++) obase += s_wtot[w];
    if (!wo.overflow) {
        const uint32_t n = (uint32_t)wo.total;
        for (uint32_t k = lane; k < n; k += 32) {
            const int64_t o = (int64_t)(obase + k);
            if (o < cap) {
                const int64_t row_q = tbase + st_q[warp][k];
                const uint32_t row_s = st_s[warp][k];
                if (out_q) __stcs(out_q + o, q.label ? __ldg(q.label + row_q) : row_q);
                if (out_s) __stcs(out_s + o, s_label ? __ldg(s_label + row_s) : (int64_t)row_s);
            }
        }
    } else {
        warp_rows<X, JM_DIRECT, SLACK>(ix, q, wbase, vec, nullptr, nullptr, 0u, out_q, out_s, s_label, cap, obase, nullptr);
    }
}

template <typename X, int MODE, bool SLACK>
__global__ void __launch_bounds__(JN_THREADS, 4)
k_join_count(iv_lists_t ix, iv_queries_t q, bool vec, int64_t* __restrict__ out_count) {
    const int warp = threadIdx.x >> 5;
    const int64_t wbase = (int64_t)blockIdx.x * JN_TILE + (int64_t)warp * JN_WROWS;
    warp_rows<X, MODE, SLACK>(ix, q, wbase, vec, nullptr, nullptr, 0u, nullptr, nullptr, nullptr, 0, 0ull, out_count);
}

// pairs per query segment of a pair list in query order (ascending query rows): lower bounds of the segment boundaries
__global__ void k_sorted_pair_segments(const int64_t* __restrict__ pair_q, const int64_t* __restrict__ total,
                                       int64_t capacity, const int64_t* __restrict__ seg_off, int n_seg,
                                       int64_t* __restrict__ out) {
    const int k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= n_seg) return;
    const int64_t n = *total < capacity ? *total : capacity;
    auto lower = [&](int64_t row) {  // first pair whose query row >= row
        int64_t lo = 0, hi = n;
        while (lo < hi) {
            const int64_t mid = (lo + hi) >> 1;
            if (__ldg(pair_q + mid) < row) lo = mid + 1; else hi = mid;
        }
        return lo;
    };
    out[k] = lower(__ldg(seg_off + k + 1)) - lower(__ldg(seg_off + k));
}

static bool aligned16(const void* p) { return (((uintptr_t)p) & 15) == 0; }
static bool narrow_span(const iv_lists_t* ix) { return ix->total_span < (int64_t)0xfffffff0ll; }

static int check_join(const iv_lists_t* ix, const iv_queries_t* q, const char* who) {
    IV_REQUIRE(ix && q, IV_ERR_INVALID, "%s: null argument", who);
    IV_REQUIRE(q->n >= 0 && (q->n == 0 || (q->start && q->end && q->seg_off && q->seg_group && q->n_seg > 0)),
               IV_ERR_INVALID, "%s: bad query descriptor", who);
    IV_REQUIRE(ix->rec && ix->shift >= 0 && ix->shift <= 15, IV_ERR_INVALID, "%s: not a list index", who);
    return IV_OK;
}

}  // namespace iv

using namespace iv;

extern "C" size_t iv_join_workspace_bytes(int64_t n_queries) {
    return ((size_t)cdiv(n_queries > 0 ? n_queries : 1, JN_TILE) + 2) * sizeof(unsigned long long) + 256;
}

extern "C" int iv_join_pairs(const iv_lists_t* lists, const iv_queries_t* q, int64_t* out_q, int64_t* out_s,
                             const int64_t* s_label, int64_t out_capacity, void* ws, size_t ws_bytes, int64_t* out_total,
                             void* stream) {
    cudaStream_t st = (cudaStream_t)stream;
    int rc = check_join(lists, q, "iv_join_pairs");
    if (rc != IV_OK) return rc;
    IV_REQUIRE(out_total, IV_ERR_INVALID, "iv_join_pairs: out_total is null");
    IV_REQUIRE(out_capacity >= 0, IV_ERR_INVALID, "iv_join_pairs: negative out_capacity");
    if (q->n == 0) {
        IV_CUDA(cudaMemsetAsync(out_total, 0, 8, st));
        return IV_OK;
    }
    IV_REQUIRE(ws && (((uintptr_t)ws) & 255) == 0 && ws_bytes >= iv_join_workspace_bytes(q->n), IV_ERR_WORKSPACE,
               "iv_join_pairs: workspace missing, misaligned or too small");
    const int64_t n_tiles = cdiv(q->n, JN_TILE);
    IV_REQUIRE(n_tiles < ((int64_t)1 << 31), IV_ERR_RANGE, "iv_join_pairs: too many rows");
    unsigned long long* state = (unsigned long long*)ws;
    unsigned int* ticket = (unsigned int*)(state + n_tiles);
    IV_CUDA(cudaMemsetAsync(ws, 0, (size_t)(n_tiles + 1) * sizeof(unsigned long long), st));
    const bool vec = aligned16(q->start) && aligned16(q->end);
    const bool slack = q->slack != 0;
    const unsigned nb = (unsigned)n_tiles;
    if (narrow_span(lists)) {
        if (slack) k_join<uint32_t, true><<<nb, JN_THREADS, 0, st>>>(*lists, *q, vec, out_q, out_s, s_label, out_capacity, state, ticket, n_tiles, out_total);
        else k_join<uint32_t, false><<<nb, JN_THREADS, 0, st>>>(*lists, *q, vec, out_q, out_s, s_label, out_capacity, state, ticket, n_tiles, out_total);
    } else {
        if (slack) k_join<uint64_t, true><<<nb, JN_THREADS, 0, st>>>(*lists, *q, vec, out_q, out_s, s_label, out_capacity, state, ticket, n_tiles, out_total);
        else k_join<uint64_t, false><<<nb, JN_THREADS, 0, st>>>(*lists, *q, vec, out_q, out_s, s_label, out_capacity, state, ticket, n_tiles, out_total);
    }
    IV_LAUNCH_CHECK();
    return IV_OK;
}

extern "C" int iv_join_count(const iv_lists_t* lists, const iv_queries_t* q, int32_t mode, int64_t* out_count, void* stream) {
    cudaStream_t st = (cudaStream_t)stream;
    int rc = check_join(lists, q, "iv_join_count");
    if (rc != IV_OK) return rc;
    IV_REQUIRE(mode == IV_MODE_ALL || mode == IV_MODE_ANY, IV_ERR_INVALID, "iv_join_count: mode must be ALL or ANY");
    if (q->n == 0) return IV_OK;
    IV_REQUIRE(out_count, IV_ERR_INVALID, "iv_join_count: out_count is null");
    const bool vec = aligned16(q->start) && aligned16(q->end) && aligned16(out_count);
    const bool slack = q->slack != 0;
    const unsigned nb = (unsigned)cdiv(q->n, JN_TILE);
#define IV_JC(X, M)                                                                                           \
    do {                                                                                                      \
        if (slack) k_join_count<X, M, true><<<nb, JN_THREADS, 0, st>>>(*lists, *q, vec, out_count);            \
        else k_join_count<X, M, false><<<nb, JN_THREADS, 0, st>>>(*lists, *q, vec, out_count);                 \
    } while (0)
    if (narrow_span(lists)) {
        if (mode == IV_MODE_ALL) IV_JC(uint32_t, JM_COUNT); else IV_JC(uint32_t, JM_ANY);
    } else {
        if (mode == IV_MODE_ALL) IV_JC(uint64_t, JM_COUNT); else IV_JC(uint64_t, JM_ANY);
    }
#undef IV_JC
    IV_LAUNCH_CHECK();
    return IV_OK;
}

extern "C" int iv_sorted_pair_segments(const int64_t* pair_q, const int64_t* total, int64_t capacity,
                                       const int64_t* seg_off, int32_t n_seg, int64_t* out_seg_pairs, void* stream) {
    cudaStream_t st = (cudaStream_t)stream;
    IV_REQUIRE(n_seg >= 0, IV_ERR_INVALID, "iv_sorted_pair_segments: n_seg < 0");
    if (n_seg == 0) return IV_OK;
    IV_REQUIRE(total && seg_off && out_seg_pairs, IV_ERR_INVALID, "iv_sorted_pair_segments: null argument");
    k_sorted_pair_segments<<<(n_seg + 127) / 128, 128, 0, st>>>(pair_q, total, capacity, seg_off, n_seg, out_seg_pairs);
    IV_LAUNCH_CHECK();
    return IV_OK;
}
