// join.cu -- the fused join: count, output offsets and emission of all overlap pairs in ONE kernel over the list
// index (lists.cu).  Replaces NCLS.all_overlaps_both (pyranges/methods/join.py:15,23) and, in its count-only form,
// all_overlaps_both + value_counts / has_overlaps (pyranges/methods/coverage.py:26-40, intersection.py:78).
//
// A block takes a ticket for 2048 consecutive query rows (so a tile only ever waits for tiles that are already running),
// a WARP is its own tile of 256 consecutive rows, a lane owns two rows per iteration:
//   1. 16-byte streaming loads of Start / End, 32-bit clamp into the partner group's window of the flattened axis,
//      ONE 256-bit gather of the record of the bin the query starts in (list offset, length, first three entries);
//   2. count sweep: two 16-bit compares per entry (inline entries from registers; longer lists and the later bins of a
//      query that reaches beyond its first bin are walked in global memory, L2-resident);
//   3. shuffle scan of the counts -> the lane's position in the warp's STAGING area in shared memory; emit sweep: the
//      same compares again, hits go to the staging area as (row in tile: 8 bit, subject row: 32 bit);
//   4. the tile's pair total is published and its output offset obtained by a decoupled look-back over the preceding
//      tiles (status word = 62-bit value + 2-bit flag, one relaxed 64-bit access each);
//   5. every warp copies its staged pairs to out_q / out_s with fully coalesced 8-byte streaming stores.
// A warp whose pairs do not fit its staging area repeats its rows once the offset is known and writes directly.
// DRAM sees the query columns once and the pair columns once; the index stays in L2.
#include <stdlib.h>

#include "common.cuh"
#include "query.cuh"

namespace iv {

constexpr int JN_THREADS = 256;
constexpr int JN_WARPS = JN_THREADS / 32;
constexpr int JN_ITERS = 4;                       // iterations of 64 rows per warp
constexpr int JN_WROWS = 64 * JN_ITERS;           // 256 rows per warp
constexpr int JN_TILE = JN_WROWS * JN_WARPS;      // 2048 rows per block
// Staged pairs per warp of k_join, 5 bytes each (iterations that do not fit are counted again once the offset is known).
// k_join serves position-ordered read sets (IV_JOIN_HINT_ORDERED): rows with many hits sit next to each other there,
// so the area is large -- measured on the sorted C3 reads: 576 pairs 3.12 ms, 896: 2.55, 1152: 2.39, 1536: 2.74 --
// while their record gathers coalesce and do not miss the L1 the shared memory takes away.  (For reads in random
// order it is the other way round: with 212 KB of shared memory per SM the count sweep ran 1.6x slower than with
// 100 KB; k_join_pipe keeps two areas of 640 pairs at three blocks per SM.)
constexpr int JN_WCAP = 1152;
constexpr size_t JN_SMEM = (size_t)JN_WARPS * JN_WCAP * 5;  // dynamic shared memory of k_join: staging areas

constexpr int JM_COUNT = 2, JM_ANY = 3;


struct LRec { uint32_t w[8]; };  // w0 = co, w1 = n8, (w2, w3) (w4, w5) (w6, w7) = entries {row, so | eo << 16}

__device__ __forceinline__ LRec load_lrec(const iv_lrec_t* __restrict__ rec, uint64_t b) {
    LRec r;
    asm("ld.global.nc.v8.u32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
        : "=r"(r.w[0]), "=r"(r.w[1]), "=r"(r.w[2]), "=r"(r.w[3]), "=r"(r.w[4]), "=r"(r.w[5]), "=r"(r.w[6]), "=r"(r.w[7])
        : "l"(rec + b));
    return r;
}

// ---- entry tests (encoding: pyranges_b200.h) ------------------------------------------------------------------------------
// first bin of a query (the one its start falls into): one subtraction decides both compares of the overlap predicate
__device__ __forceinline__ uint32_t hit_bit(uint32_t se, uint32_t qw) {  // bit 31 = hit
    const uint32_t t = se - qw;
    return t & (t << 16) & 0x80000000u;
}
// later bins: only subjects that START there, before the query's end: sp in [1, rel_e]
__device__ __forceinline__ bool later_hit(uint32_t se, uint32_t rel_e) {
    const uint32_t sp = IV_L_SMAX - (se & 0x7fffu);
    return sp >= 1u && sp <= rel_e;
}

// ---- what the count sweep learns about a query ------------------------------------------------------------------------------
// The first JN_MASKED entries of the list of the bin the query starts in are tested from registers: three from the bin
// record, eight from ONE batch of loads (all in flight together -- a short list costs one round trip, not one per
// entry).  mask = which of them are hits (bit j = entry j); extra = hits beyond: entries past the first eleven of a long
// list and the starts inside the later bins of a query that reaches beyond its first bin (walked entry by entry: rare).
constexpr int JN_MASKED = 11;

struct QPlan {
    uint32_t mask, extra;
    __device__ __forceinline__ uint32_t count() const { return __popc(mask) + extra; }
};

template <typename X> struct QGeom {
    X b;
    uint32_t qw, n;
    bool single;
    __device__ __forceinline__ QGeom(const iv_lists_t& ix, X sc, X ec, const LRec& r) {
        const int shift = ix.shift;
        b = sc >> shift;
        const X origin = b << shift;
        const uint32_t rel_s = (uint32_t)(sc - origin);
        const X d = ec - origin;  // ec >= sc: no wrap
        const uint32_t width = 1u << shift;
        // ends within bin width + reach: the list of this bin holds every candidate; else only this bin's own subjects
        // are taken from it (sp <= width) and the later bins are walked
        single = d <= (X)(width + (uint32_t)ix.reach);
        qw = (IV_L_SMAX - (single ? (uint32_t)d : width)) | ((rel_s + 1u) << 16);
        n = r.w[1] & 0xffu;
        if (n == 255u) n = __ldg(&ix.rec[(uint64_t)b + 1].co) - r.w[0];
    }
};

// hits beyond the masked entries: f(subject row)
template <typename X, typename F>
__device__ __forceinline__ void walk_extra(const iv_lists_t& ix, X ec, const LRec& r, const QGeom<X>& g, F&& f) {
    const int shift = ix.shift;
    // the rest of a long list, eight entries per round trip
    for (uint32_t j0 = JN_MASKED; j0 < g.n; j0 += 8) {
        const uint32_t* __restrict__ c = ix.cand_se + r.w[0] + j0;
        uint32_t w[8];
#pragma unroll
        for (int t = 0; t < 8; t++) w[t] = __ldg(c + t);  // (the arrays end in slack: reading past the list is safe)
#pragma unroll
        for (int t = 0; t < 8; t++)
            if (j0 + t < g.n && hit_bit(w[t], g.qw)) f(__ldg(ix.cand_row + r.w[0] + j0 + t));
    }
    if (!g.single) {
        const uint64_t bl = (uint64_t)(ec - 1) >> shift;
        for (uint64_t bb = (uint64_t)g.b + 1; bb <= bl; bb++) {
            const uint32_t c0 = __ldg(&ix.rec[bb].co), c1 = __ldg(&ix.rec[bb + 1].co);
            const uint32_t re = bb < bl ? (1u << shift) : (uint32_t)(ec - (X)(bl << shift));
            for (uint32_t j = c0; j < c1; j++)
                if (later_hit(__ldg(ix.cand_se + j), re)) f(__ldg(ix.cand_row + j));
        }
    }
}

template <typename X>
__device__ __forceinline__ QPlan plan_query(const iv_lists_t& ix, X sc, X ec, const LRec& r, bool live) {
    QPlan p{0u, 0u};
    if (!live) return p;
    const QGeom<X> g(ix, sc, ec, r);
    uint32_t m = 0;
    if (g.n > 3) {  // entries 3 .. 10: eight loads in flight together; processed last to first so that entry j lands on bit j
        const uint32_t* __restrict__ c = ix.cand_se + r.w[0] + 3u;
        uint32_t w[8];
#pragma unroll
        for (int t = 0; t < 8; t++) w[t] = __ldg(c + t);
#pragma unroll
        for (int t = 7; t >= 0; t--) m = __funnelshift_l(hit_bit(w[t], g.qw), m, 1);
    }
    m = __funnelshift_l(hit_bit(r.w[7], g.qw), m, 1);
    m = __funnelshift_l(hit_bit(r.w[5], g.qw), m, 1);
    m = __funnelshift_l(hit_bit(r.w[3], g.qw), m, 1);
    p.mask = m & (g.n >= (uint32_t)JN_MASKED ? (1u << JN_MASKED) - 1u : (1u << g.n) - 1u);
    if (g.n > (uint32_t)JN_MASKED || !g.single) {
        uint32_t x = 0;
        walk_extra<X>(ix, ec, r, g, [&](uint32_t) { x++; });
        p.extra = x;
    }
    return p;
}

// every hit of a planned query, in list order: f(subject row)
template <typename X, typename F>
__device__ __forceinline__ void emit_query(const iv_lists_t& ix, X sc, X ec, const LRec& r, const QPlan& p, F&& f) {
    uint32_t m = p.mask;
    if (m & 1u) f(r.w[2]);
    if (m & 2u) f(r.w[4]);
    if (m & 4u) f(r.w[6]);
    m >>= 3;
    if (m) {  // hits among entries 3 .. 10: their rows in ONE round trip (predicated loads, all in flight together)
        const uint32_t* __restrict__ c = ix.cand_row + r.w[0] + 3u;
        uint32_t row[8];
#pragma unroll
        for (int t = 0; t < 8; t++) row[t] = (m >> t) & 1u ? __ldg(c + t) : 0u;
#pragma unroll
        for (int t = 0; t < 8; t++)
            if ((m >> t) & 1u) f(row[t]);
    }
    if (p.extra) {
        const QGeom<X> g(ix, sc, ec, r);
        walk_extra<X>(ix, ec, r, g, f);
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
// called by one full warp: the tile's own total becomes visible to its successors
__device__ __forceinline__ void publish_total(unsigned long long* state, int64_t tile, unsigned long long aggregate) {
    if (lane_id() == 0) st_relaxed(state + tile, (aggregate << 2) | (tile == 0 ? 2ull : 1ull));
}
// called by one full warp after publish_total; -> pairs of all tiles before `tile`
__device__ __forceinline__ unsigned long long lookback(unsigned long long* state, int64_t tile, unsigned long long aggregate) {
    const int lane = lane_id();
    unsigned long long excl = 0;
    int64_t look = tile - 1;
    while (look >= 0) {
        const int64_t idx = look - lane;
        unsigned long long v = idx >= 0 ? ld_relaxed(state + idx) : 2ull;  // before the first tile: inclusive prefix 0
        // poll as a warp (one uniform loop, not 32 divergent ones), backing off while predecessors are still counting
        for (unsigned ns = 32; __any_sync(FULL, (v & 3ull) == 0ull); ns = ns < 512 ? ns * 2 : ns) {
            __nanosleep(ns);
            if ((v & 3ull) == 0ull) v = ld_relaxed(state + idx);
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
// segment geometry of the rows a warp walks over (monotone: re-derived only when the key changes)
template <typename X> struct Geo {
    int seg, f_seg;
    Flat<X> f0;
    uint32_t lo32, hi32, off32;
    bool geo32;
    __device__ __forceinline__ Geo(const iv_queries_t& q, int64_t row) : f_seg(-1), lo32(0), hi32(0), off32(0), geo32(false) {
        seg = find_seg(q.seg_off, q.n_seg, row);
    }
};

// raw Start / End of the two rows (ibase + 2 * lane, + 1) of a lane: issued one iteration ahead
struct Raw2 { int64_t s[2], e[2]; };
__device__ __forceinline__ Raw2 load_raw(const iv_queries_t& q, bool vec, int64_t ibase) {
    Raw2 w;
    w.s[0] = w.s[1] = w.e[0] = w.e[1] = 0;
    const int64_t r0 = ibase + 2 * lane_id();
    if (vec && r0 + 1 < q.n) {
        const longlong2 s2 = __ldcs(reinterpret_cast<const longlong2*>(q.start + r0));
        const longlong2 e2 = __ldcs(reinterpret_cast<const longlong2*>(q.end + r0));
        w.s[0] = s2.x; w.s[1] = s2.y; w.e[0] = e2.x; w.e[1] = e2.y;
    } else {
        if (r0 < q.n) { w.s[0] = __ldcs(q.start + r0); w.e[0] = __ldcs(q.end + r0); }
        if (r0 + 1 < q.n) { w.s[1] = __ldcs(q.start + r0 + 1); w.e[1] = __ldcs(q.end + r0 + 1); }
    }
    return w;
}

// flattened interval, liveness and the record of the bin the query starts in, for the lane's two rows
template <typename X> struct Rows2 {
    X sc[2], ec[2];
    LRec r[2];
    bool live[2];
};

template <typename X, bool SLACK>
__device__ __forceinline__ void flatten_rows(const iv_lists_t& ix, const iv_queries_t& q, int64_t ibase, Raw2 w, Geo<X>& G,
                                             Rows2<X>& R) {
    const int shift = ix.shift;
    while (G.seg + 1 < q.n_seg && __ldg(q.seg_off + G.seg + 1) <= ibase) G.seg++;
    const bool uniform = __ldg(q.seg_off + G.seg + 1) >= ibase + 64;
    if (G.seg != G.f_seg) {
        G.f0 = load_flat_groups<X>(ix.groups, ix.n, q, G.seg);
        G.f_seg = G.seg;
        const uint64_t hi = (uint64_t)G.f0.lo + G.f0.width;
        G.geo32 = sizeof(X) == 4 && G.f0.lo >= 0 && hi <= 0xffffffffull;
        G.lo32 = (uint32_t)G.f0.lo;
        G.hi32 = (uint32_t)hi;
        G.off32 = (uint32_t)G.f0.base - G.lo32;
    }
    const int64_t r0 = ibase + 2 * lane_id();
#pragma unroll
    for (int k = 0; k < 2; k++) {
        int64_t s = w.s[k], e = w.e[k];
        if (SLACK) apply_slack(q.slack, s, e);
        bool partner;
        if (uniform && G.geo32 && (((uint64_t)s | (uint64_t)e) >> 32) == 0) {
            R.sc[k] = (X)(min(max((uint32_t)s, G.lo32), G.hi32) + G.off32);
            R.ec[k] = (X)(min(max((uint32_t)e, G.lo32), G.hi32) + G.off32);
            partner = G.f0.g >= 0;
        } else {
            Flat<X> f = G.f0;
            if (!uniform) {
                const int sg = find_seg_range(q.seg_off, G.seg, q.n_seg - 1, r0 + k < q.n ? r0 + k : q.n - 1);
                if (sg != G.seg) f = load_flat_groups<X>(ix.groups, ix.n, q, sg);
            }
            R.sc[k] = f.at(s);
            R.ec[k] = f.at(e);
            partner = f.g >= 0;
        }
        // rows beyond the end, keys without a partner and malformed rows (end < start) have no hits
        R.live[k] = r0 + k < q.n && partner && R.ec[k] >= R.sc[k];
        if (!R.live[k]) { R.sc[k] = 0; R.ec[k] = 0; }
        R.r[k] = load_lrec(ix.rec, (uint64_t)(R.sc[k] >> shift));
    }
}

// count-only kernels: out_count[row] = hits (JM_COUNT) or 0 / 1 (JM_ANY)
template <typename X, int MODE, bool SLACK>
__global__ void __launch_bounds__(JN_THREADS, 4)
k_join_count(iv_lists_t ix, iv_queries_t q, bool vec, int64_t* __restrict__ out_count) {
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int64_t wbase = (int64_t)blockIdx.x * JN_TILE + (int64_t)warp * JN_WROWS;
    if (wbase >= q.n) return;
    Geo<X> G(q, wbase);
    Raw2 next = load_raw(q, vec, wbase);
#pragma unroll 1
    for (int it = 0; it < JN_ITERS; it++) {
        const int64_t ibase = wbase + it * 64;
        if (ibase >= q.n) break;
        const Raw2 cur = next;
        if (it + 1 < JN_ITERS && ibase + 64 < q.n) next = load_raw(q, vec, ibase + 64);
        Rows2<X> R;
        flatten_rows<X, SLACK>(ix, q, ibase, cur, G, R);
        uint32_t c[2];
#pragma unroll
        for (int k = 0; k < 2; k++) {
            c[k] = plan_query<X>(ix, R.sc[k], R.ec[k], R.r[k], R.live[k]).count();
            if (MODE == JM_ANY) c[k] = c[k] ? 1u : 0u;
        }
        const int64_t r0 = ibase + 2 * lane;
        if (vec && r0 + 1 < q.n) {
            __stcs(reinterpret_cast<longlong2*>(out_count + r0), make_longlong2((long long)c[0], (long long)c[1]));
        } else {
            if (r0 < q.n) out_count[r0] = c[0];
            if (r0 + 1 < q.n) out_count[r0 + 1] = c[1];
        }
    }
}

// ---- the fused join ---------------------------------------------------------------------------------------------------------------
// Every WARP is its own tile of 256 rows (tile id = block ticket * 8 + warp: the warps of a block are co-resident and
// blocks with smaller tickets started earlier, so a tile only waits for tiles that are running or done); no block-wide
// barrier after the ticket.  The warp stages the pairs of its iterations in shared memory while they fit, publishes its
// pair total, looks back for its output offset and copies the staged pairs out with coalesced stores; iterations that
// did not fit are repeated afterwards -- the offset is known by then -- through the same staging area, one at a time.
template <typename X, bool SLACK, bool ANY>
__global__ void __launch_bounds__(JN_THREADS, 4)
k_join(iv_lists_t ix, iv_queries_t q, bool vec, int64_t* __restrict__ out_q, int64_t* __restrict__ out_s,
       const int64_t* __restrict__ s_label, int64_t cap, unsigned long long* __restrict__ state,
       unsigned int* __restrict__ ticket, int64_t n_wtiles, int64_t* __restrict__ out_total, int variant) {
    extern __shared__ __align__(16) unsigned char jn_smem[];  // [JN_WARPS][JN_WCAP] subject rows, then query rows
    __shared__ uint32_t s_cnt[JN_WARPS][JN_ITERS];
    __shared__ unsigned int s_tile;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (threadIdx.x == 0) s_tile = atomicAdd(ticket, 1u);
    __syncthreads();
    const int64_t wtile = (int64_t)s_tile * JN_WARPS + warp;
    if (wtile >= n_wtiles) return;
    const int64_t wbase = wtile * JN_WROWS;
    uint32_t* __restrict__ st_s = reinterpret_cast<uint32_t*>(jn_smem) + warp * JN_WCAP;
    uint8_t* __restrict__ st_q = jn_smem + (size_t)JN_WARPS * JN_WCAP * 4 + warp * JN_WCAP;

    // coalesced copy of `n` staged pairs to the outputs at offset o0; rows are relative to rbase
    auto copy_out = [&](uint32_t n, unsigned long long o0, int64_t rbase) {
        for (uint32_t k = lane; k < n; k += 32) {
            const int64_t o = (int64_t)(o0 + k);
            if (o < cap) {
                const int64_t row_q = rbase + st_q[k];
                if (out_q) __stcs(out_q + o, q.label ? __ldg(q.label + row_q) : row_q);
                if (!ANY && out_s) {
                    const uint32_t row_s = st_s[k];
                    __stcs(out_s + o, s_label ? __ldg(s_label + row_s) : (int64_t)row_s);
                }
            }
        }
    };
    // emit sweep of one iteration into the staging area, the lane's first pair at `pos`
    auto stage_rows = [&](const Rows2<X>& R, const QPlan (&p)[2], uint32_t pos, uint32_t qloc) {
#pragma unroll
        for (int k = 0; k < 2; k++) {
            const uint8_t ql = (uint8_t)(qloc + k);
            if (ANY) {  // one entry per query with hits: the query row only
                if (p[k].count()) st_q[pos++] = ql;
            } else {
                emit_query<X>(ix, R.sc[k], R.ec[k], R.r[k], p[k], [&](uint32_t row) {
                    st_s[pos] = row;
                    st_q[pos] = ql;
                    pos++;
                });
            }
        }
    };
    auto pairs_of = [&](const QPlan (&p)[2]) -> uint32_t {
        return ANY ? (p[0].count() ? 1u : 0u) + (p[1].count() ? 1u : 0u) : p[0].count() + p[1].count();
    };

    Geo<X> G(q, wbase);
    unsigned long long total = 0;
    uint32_t staged = 0;
    int staged_iters = 0;
    bool open = true;
    Raw2 next = load_raw(q, vec, wbase);
#pragma unroll 1
    for (int it = 0; it < JN_ITERS; it++) {
        const int64_t ibase = wbase + it * 64;
        uint32_t itot = 0;
        if (ibase < q.n) {
            const Raw2 cur = next;
            if (it + 1 < JN_ITERS && ibase + 64 < q.n) next = load_raw(q, vec, ibase + 64);
            Rows2<X> R;
            flatten_rows<X, SLACK>(ix, q, ibase, cur, G, R);
            QPlan p[2];
#pragma unroll
            for (int k = 0; k < 2; k++) p[k] = plan_query<X>(ix, R.sc[k], R.ec[k], R.r[k], R.live[k]);
            const uint32_t mine = pairs_of(p);
            const uint32_t inc = warp_incl_sum(mine);
            itot = __shfl_sync(FULL, inc, 31);
            if (open && staged + itot <= (uint32_t)JN_WCAP) {
                if (itot) stage_rows(R, p, staged + inc - mine, (uint32_t)(it * 64 + 2 * lane));
                staged += itot;
                staged_iters = it + 1;
            } else {
                open = false;
            }
        }
        if (lane == 0) s_cnt[warp][it] = itot;
        total += itot;
    }
    __syncwarp();
    unsigned long long o;
    if (variant == 1) {  // EXPERIMENT: unordered output regions from an atomic cursor (no waiting)
        if (lane == 0) o = atomicAdd(reinterpret_cast<unsigned long long*>(out_total), total);
        o = __shfl_sync(FULL, o, 0);
    } else {
        publish_total(state, wtile, total);
        o = lookback(state, wtile, total);
        if (wtile == n_wtiles - 1 && lane == 0) *out_total = (int64_t)(o + total);
    }
    copy_out(staged, o, wbase);
    o += staged;
    // iterations that did not fit: once more, through the staging area (or straight to memory when a single iteration
    // exceeds it)
    for (int it = staged_iters; it < JN_ITERS; it++) {
        const uint32_t itot = s_cnt[warp][it];
        if (itot == 0) continue;
        const int64_t ibase = wbase + it * 64;
        __syncwarp();
        Geo<X> G2(q, ibase);
        Rows2<X> R;
        flatten_rows<X, SLACK>(ix, q, ibase, load_raw(q, vec, ibase), G2, R);
        QPlan p[2];
#pragma unroll
        for (int k = 0; k < 2; k++) p[k] = plan_query<X>(ix, R.sc[k], R.ec[k], R.r[k], R.live[k]);
        const uint32_t mine = pairs_of(p);
        const uint32_t inc = warp_incl_sum(mine);
        if (itot <= (uint32_t)JN_WCAP) {
            stage_rows(R, p, inc - mine, (uint32_t)(2 * lane));
            __syncwarp();
            copy_out(itot, o, ibase);
        } else {
            unsigned long long oo = o + inc - mine;
#pragma unroll
            for (int k = 0; k < 2; k++) {
                const int64_t row_q = ibase + 2 * lane + k;
                const int64_t ql = (q.label && row_q < q.n) ? __ldg(q.label + row_q) : row_q;
                if (ANY) {
                    if (p[k].count()) {
                        if ((int64_t)oo < cap && out_q) out_q[oo] = ql;
                        oo++;
                    }
                    continue;
                }
                emit_query<X>(ix, R.sc[k], R.ec[k], R.r[k], p[k], [&](uint32_t row) {
                    if ((int64_t)oo < cap) {
                        if (out_q) out_q[oo] = ql;
                        if (out_s) out_s[oo] = s_label ? __ldg(s_label + row) : (int64_t)row;
                    }
                    oo++;
                });
            }
        }
        o += itot;
    }
}

// ---- the fused join, pipelined ---------------------------------------------------------------------------------------
// k_join makes a tile wait for its output offset right after counting it: with thousands of tiles in flight the
// stragglers among a tile's predecessors cost about a third of the kernel.  Here every warp has TWO staging areas: it
// counts and stages tile j + 1 (and publishes its total) before it looks back for tile j, whose staged pairs wait in the
// other area -- the predecessors of tile j had a whole tile's time to publish.  One ticket per tile.  Measured at C3
// (100 M x 1 M): 2.31 ms against 2.60 ms for k_join; the staging capacity matters more than anything else (an iteration
// that does not fit is counted again after the look-back): 320 pairs 6.2 ms, 384: 4.4, 512: 2.8, 640 / 704 / 768 at three
// blocks per SM: 2.31 / 2.35 / 2.34, 896: 2.62; two tiles per warp 2.93, four 2.31, eight 2.29.  The grid is persistent
// (as many blocks as the device holds, tiles handed out one by one): no partly filled last wave at small sizes.
constexpr int JP_WCAP = 640;     // pairs per staging area
constexpr int JP_MINB = 3;       // blocks per SM: 2 x 8 x 640 x 5 bytes = 51 KB each
constexpr size_t JP_SMEM = (size_t)2 * JN_WARPS * JP_WCAP * 5;

template <typename X, bool SLACK>
__global__ void __launch_bounds__(JN_THREADS, JP_MINB)
k_join_pipe(iv_lists_t ix, iv_queries_t q, bool vec, int64_t* __restrict__ out_q, int64_t* __restrict__ out_s,
            const int64_t* __restrict__ s_label, int64_t cap, unsigned long long* __restrict__ state,
            unsigned int* __restrict__ ticket, int64_t n_wtiles, int64_t* __restrict__ out_total) {
    extern __shared__ __align__(16) unsigned char jn_smem[];  // [2][JN_WARPS][JP_WCAP] subject rows, then query rows
    __shared__ uint32_t s_cnt[2][JN_WARPS][JN_ITERS];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    auto area_s = [&](int a) { return reinterpret_cast<uint32_t*>(jn_smem) + (a * JN_WARPS + warp) * JP_WCAP; };
    auto area_q = [&](int a) { return jn_smem + (size_t)2 * JN_WARPS * JP_WCAP * 4 + (size_t)(a * JN_WARPS + warp) * JP_WCAP; };

    auto copy_out = [&](const uint32_t* st_s, const uint8_t* st_q, uint32_t n, unsigned long long o0, int64_t rbase) {
        for (uint32_t k = lane; k < n; k += 32) {
            const int64_t o = (int64_t)(o0 + k);
            if (o < cap) {
                const int64_t row_q = rbase + st_q[k];
                if (out_q) __stcs(out_q + o, q.label ? __ldg(q.label + row_q) : row_q);
                if (out_s) {
                    const uint32_t row_s = st_s[k];
                    __stcs(out_s + o, s_label ? __ldg(s_label + row_s) : (int64_t)row_s);
                }
            }
        }
    };
    auto stage_rows = [&](uint32_t* st_s, uint8_t* st_q, const Rows2<X>& R, const QPlan (&p)[2], uint32_t pos, uint32_t qloc) {
#pragma unroll
        for (int k = 0; k < 2; k++) {
            const uint8_t ql = (uint8_t)(qloc + k);
            emit_query<X>(ix, R.sc[k], R.ec[k], R.r[k], p[k], [&](uint32_t row) {
                st_s[pos] = row;
                st_q[pos] = ql;
                pos++;
            });
        }
    };
    // the tile waiting for its offset
    int64_t ptile = -1;
    unsigned long long ptotal = 0;
    uint32_t pstaged = 0;
    int pstaged_iters = 0, parea = 0;

    auto write_tile = [&](int64_t wtile, int area, unsigned long long total, uint32_t staged, int staged_iters) {
        unsigned long long o = lookback(state, wtile, total);
        if (wtile == n_wtiles - 1 && lane == 0) *out_total = (int64_t)(o + total);
        const int64_t wbase = wtile * JN_WROWS;
        uint32_t* st_s = area_s(area);
        uint8_t* st_q = area_q(area);
        copy_out(st_s, st_q, staged, o, wbase);
        o += staged;
        // iterations that did not fit: once more, through the staging area (or straight to memory)
        for (int it = staged_iters; it < JN_ITERS; it++) {
            const uint32_t itot = s_cnt[area][warp][it];
            if (itot == 0) continue;
            const int64_t ibase = wbase + it * 64;
            __syncwarp();
            Geo<X> G2(q, ibase);
            Rows2<X> R;
            flatten_rows<X, SLACK>(ix, q, ibase, load_raw(q, vec, ibase), G2, R);
            QPlan p[2];
#pragma unroll
            for (int k = 0; k < 2; k++) p[k] = plan_query<X>(ix, R.sc[k], R.ec[k], R.r[k], R.live[k]);
            const uint32_t mine = p[0].count() + p[1].count();
            const uint32_t inc = warp_incl_sum(mine);
            if (itot <= (uint32_t)JP_WCAP) {
                stage_rows(st_s, st_q, R, p, inc - mine, (uint32_t)(2 * lane));
                __syncwarp();
                copy_out(st_s, st_q, itot, o, ibase);
            } else {
                unsigned long long oo = o + inc - mine;
#pragma unroll
                for (int k = 0; k < 2; k++) {
                    const int64_t row_q = ibase + 2 * lane + k;
                    const int64_t ql = (q.label && row_q < q.n) ? __ldg(q.label + row_q) : row_q;
                    emit_query<X>(ix, R.sc[k], R.ec[k], R.r[k], p[k], [&](uint32_t row) {
                        if ((int64_t)oo < cap) {
                            if (out_q) out_q[oo] = ql;
                            if (out_s) out_s[oo] = s_label ? __ldg(s_label + row) : (int64_t)row;
                        }
                        oo++;
                    });
                }
            }
            o += itot;
        }
        __syncwarp();
    };

#pragma unroll 1
    for (int j = 0;; j++) {  // persistent: the grid is as large as the device holds, a warp takes tiles until none is left
        unsigned int tk = 0;
        if (lane == 0) tk = atomicAdd(ticket, 1u);
        const int64_t wtile = (int64_t)__shfl_sync(FULL, tk, 0);
        if (wtile >= n_wtiles) break;
        const int64_t wbase = wtile * JN_WROWS;
        const int area = j & 1;
        uint32_t* st_s = area_s(area);
        uint8_t* st_q = area_q(area);
        Geo<X> G(q, wbase);
        unsigned long long total = 0;
        uint32_t staged = 0;
        int staged_iters = 0;
        bool open = true;
        Raw2 next = load_raw(q, vec, wbase);
#pragma unroll 1
        for (int it = 0; it < JN_ITERS; it++) {
            const int64_t ibase = wbase + it * 64;
            uint32_t itot = 0;
            if (ibase < q.n) {
                const Raw2 cur = next;
                if (it + 1 < JN_ITERS && ibase + 64 < q.n) next = load_raw(q, vec, ibase + 64);
                Rows2<X> R;
                flatten_rows<X, SLACK>(ix, q, ibase, cur, G, R);
                QPlan p[2];
#pragma unroll
                for (int k = 0; k < 2; k++) p[k] = plan_query<X>(ix, R.sc[k], R.ec[k], R.r[k], R.live[k]);
                const uint32_t mine = p[0].count() + p[1].count();
                const uint32_t inc = warp_incl_sum(mine);
                itot = __shfl_sync(FULL, inc, 31);
                if (open && staged + itot <= (uint32_t)JP_WCAP) {
                    if (itot) stage_rows(st_s, st_q, R, p, staged + inc - mine, (uint32_t)(it * 64 + 2 * lane));
                    staged += itot;
                    staged_iters = it + 1;
                } else {
                    open = false;
                }
            }
            if (lane == 0) s_cnt[area][warp][it] = itot;
            total += itot;
        }
        __syncwarp();
        publish_total(state, wtile, total);
        if (ptile >= 0) write_tile(ptile, parea, ptotal, pstaged, pstaged_iters);
        ptile = wtile;
        parea = area;
        ptotal = total;
        pstaged = staged;
        pstaged_iters = staged_iters;
    }
    if (ptile >= 0) write_tile(ptile, parea, ptotal, pstaged, pstaged_iters);
}

// ---- overlap(how="first"): the rows with at least one hit -----------------------------------------------------------
// The same tiles and the same look-back as k_join, but a row contributes at most one entry -- its own number -- so
// nothing is staged: the hit flags of a tile's 256 rows live in registers (two ballots per iteration) and the positions
// inside the tile come from population counts.  That makes the ordering wait avoidable: a warp of the persistent grid takes tile after tile, one
// ticket each, and writes a tile only after it has counted and published the next one -- the tile's predecessors had
// a whole tile's time to publish theirs.

template <typename X, bool SLACK>
__global__ void __launch_bounds__(JN_THREADS, 4)
k_join_any(iv_lists_t ix, iv_queries_t q, bool vec, int64_t* __restrict__ out_q, int64_t cap,
           unsigned long long* __restrict__ state, unsigned int* __restrict__ ticket, int64_t n_wtiles,
           int64_t* __restrict__ out_total) {
    const int lane = threadIdx.x & 31;
    const unsigned below = (1u << lane) - 1u;
    unsigned p0[JN_ITERS], p1[JN_ITERS];  // hit flags of the tile waiting for its offset
    uint32_t ptotal = 0;
    int64_t ptile = -1;

    auto write_tile = [&](int64_t wtile, const unsigned (&b0)[JN_ITERS], const unsigned (&b1)[JN_ITERS], uint32_t total) {
        unsigned long long o = lookback(state, wtile, total);
        if (wtile == n_wtiles - 1 && lane == 0) *out_total = (int64_t)(o + total);
        if (!out_q) return;
        const int64_t wbase = wtile * JN_WROWS;
#pragma unroll
        for (int it = 0; it < JN_ITERS; it++) {
            // rows of the iteration in order: (lane 0: k = 0, 1), (lane 1: k = 0, 1) ...
            const uint32_t before = __popc(b0[it] & below) + __popc(b1[it] & below);
            const int64_t row = wbase + it * 64 + 2 * lane;
            if (b0[it] >> lane & 1u) {
                const int64_t at = (int64_t)o + before;
                if (at < cap) __stcs(out_q + at, q.label ? __ldg(q.label + row) : row);
            }
            if (b1[it] >> lane & 1u) {
                const int64_t at = (int64_t)o + before + (b0[it] >> lane & 1u);
                if (at < cap) __stcs(out_q + at, q.label ? __ldg(q.label + row + 1) : row + 1);
            }
            o += __popc(b0[it]) + __popc(b1[it]);
        }
    };

#pragma unroll 1
    for (;;) {  // persistent: a warp takes tiles until none is left
        // one ticket per tile: a tile only ever waits for tiles taken earlier, all of them by running warps
        unsigned int tk = 0;
        if (lane == 0) tk = atomicAdd(ticket, 1u);
        const int64_t wtile = (int64_t)__shfl_sync(FULL, tk, 0);
        if (wtile >= n_wtiles) break;
        const int64_t wbase = wtile * JN_WROWS;
        Geo<X> G(q, wbase);
        unsigned b0[JN_ITERS], b1[JN_ITERS];
        uint32_t total = 0;
        Raw2 next = load_raw(q, vec, wbase);
#pragma unroll
        for (int it = 0; it < JN_ITERS; it++) {
            const int64_t ibase = wbase + it * 64;
            bool h0 = false, h1 = false;
            if (ibase < q.n) {
                const Raw2 cur = next;
                if (it + 1 < JN_ITERS && ibase + 64 < q.n) next = load_raw(q, vec, ibase + 64);
                Rows2<X> R;
                flatten_rows<X, SLACK>(ix, q, ibase, cur, G, R);
                h0 = plan_query<X>(ix, R.sc[0], R.ec[0], R.r[0], R.live[0]).count() != 0;
                h1 = plan_query<X>(ix, R.sc[1], R.ec[1], R.r[1], R.live[1]).count() != 0;
            }
            b0[it] = __ballot_sync(FULL, h0);
            b1[it] = __ballot_sync(FULL, h1);
            total += __popc(b0[it]) + __popc(b1[it]);
        }
        publish_total(state, wtile, total);
        if (ptile >= 0) write_tile(ptile, p0, p1, ptotal);
#pragma unroll
        for (int it = 0; it < JN_ITERS; it++) { p0[it] = b0[it]; p1[it] = b1[it]; }
        ptotal = total;
        ptile = wtile;
    }
    if (ptile >= 0) write_tile(ptile, p0, p1, ptotal);
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
    IV_REQUIRE(ix->rec && ix->shift >= 0 && ix->shift <= IV_L_MAX_SHIFT, IV_ERR_INVALID, "%s: not a list index", who);
    return IV_OK;
}

}  // namespace iv

using namespace iv;

extern "C" size_t iv_join_workspace_bytes(int64_t n_queries) {
    return ((size_t)cdiv(n_queries > 0 ? n_queries : 1, JN_WROWS) + 2) * sizeof(unsigned long long) + 256;
}

extern "C" int iv_join_pairs(const iv_lists_t* lists, const iv_queries_t* q, int32_t mode, int64_t* out_q, int64_t* out_s,
                             const int64_t* s_label, int64_t out_capacity, void* ws, size_t ws_bytes, int64_t* out_total,
                             void* stream) {
    cudaStream_t st = (cudaStream_t)stream;
    int rc = check_join(lists, q, "iv_join_pairs");
    if (rc != IV_OK) return rc;
    IV_REQUIRE(out_total, IV_ERR_INVALID, "iv_join_pairs: out_total is null");
    const bool ordered = (mode & IV_JOIN_HINT_ORDERED) != 0;
    mode &= ~IV_JOIN_HINT_ORDERED;
    IV_REQUIRE(mode == IV_MODE_ALL || mode == IV_MODE_ANY, IV_ERR_INVALID, "iv_join_pairs: mode must be ALL or ANY");
    IV_REQUIRE(out_capacity >= 0, IV_ERR_INVALID, "iv_join_pairs: negative out_capacity");
    if (q->n == 0) {
        IV_CUDA(cudaMemsetAsync(out_total, 0, 8, st));
        return IV_OK;
    }
    IV_REQUIRE(ws && (((uintptr_t)ws) & 255) == 0 && ws_bytes >= iv_join_workspace_bytes(q->n), IV_ERR_WORKSPACE,
               "iv_join_pairs: workspace missing, misaligned or too small");
    const int64_t n_tiles = cdiv(q->n, JN_WROWS);  // one tile per warp
    IV_REQUIRE(n_tiles < ((int64_t)1 << 31), IV_ERR_RANGE, "iv_join_pairs: too many rows");
    unsigned long long* state = (unsigned long long*)ws;
    unsigned int* ticket = (unsigned int*)(state + n_tiles);
    IV_CUDA(cudaMemsetAsync(ws, 0, (size_t)(n_tiles + 1) * sizeof(unsigned long long), st));
    const bool vec = aligned16(q->start) && aligned16(q->end);
    const bool slack = q->slack != 0;
    const unsigned nb = (unsigned)cdiv(q->n, JN_TILE);
    static const int variant = getenv("IV_JOIN_VARIANT") ? atoi(getenv("IV_JOIN_VARIANT")) : 0;  // measurements only
    if (variant == 1) IV_CUDA(cudaMemsetAsync(out_total, 0, 8, st));
#define IV_JK(X, SL, AN)                                                                                                \
    do {                                                                                                               \
        static bool attr_done[64] = {false};                                                                           \
        int dev = 0;                                                                                                   \
        IV_CUDA(cudaGetDevice(&dev));                                                                                  \
        if (dev < 0 || dev >= 64 || !attr_done[dev]) {                                                                 \
            IV_CUDA(cudaFuncSetAttribute(k_join<X, SL, AN>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)JN_SMEM)); \
            if (dev >= 0 && dev < 64) attr_done[dev] = true;                                                           \
        }                                                                                                              \
        k_join<X, SL, AN><<<nb, JN_THREADS, JN_SMEM, st>>>(*lists, *q, vec, out_q, out_s, s_label, out_capacity, state,  \
                                                          ticket, n_tiles, out_total, variant);                        \
    } while (0)
#define IV_JK2(X, SL)                                                                                                  \
    do {                                                                                                               \
        if (any && variant != 3) {                                                                                     \
            static int fit_any[64] = {0};                                                                              \
            int dev = 0;                                                                                               \
            IV_CUDA(cudaGetDevice(&dev));                                                                              \
            if (dev < 0 || dev >= 64 || !fit_any[dev]) {                                                               \
                int sms = 0, per_sm = 0;                                                                               \
                IV_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev < 0 ? 0 : dev));               \
                IV_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, k_join_any<X, SL>, JN_THREADS, 0));      \
                if (dev >= 0 && dev < 64) fit_any[dev] = sms * per_sm > 0 ? sms * per_sm : 148;                        \
            }                                                                                                          \
            const int64_t fit = (dev >= 0 && dev < 64) ? fit_any[dev] : 148, need = cdiv(n_tiles, JN_WARPS);           \
            k_join_any<X, SL><<<(unsigned)(need < fit ? need : fit), JN_THREADS, 0, st>>>(                              \
                *lists, *q, vec, out_q, out_capacity, state, ticket, n_tiles, out_total);                              \
        } else if (any) IV_JK(X, SL, true);                                                                            \
        else if (variant != 2 && !ordered) {  /* variant 2: the unpipelined k_join for any input (measurements) */                                                                                       \
            static int fit_pipe[64] = {0};                                                                             \
            int dev = 0;                                                                                               \
            IV_CUDA(cudaGetDevice(&dev));                                                                              \
            if (dev < 0 || dev >= 64 || !fit_pipe[dev]) {                                                              \
                IV_CUDA(cudaFuncSetAttribute(k_join_pipe<X, SL>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)JP_SMEM)); \
                int sms = 0, per_sm = 0;                                                                               \
                IV_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev < 0 ? 0 : dev));               \
                IV_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, k_join_pipe<X, SL>, JN_THREADS, JP_SMEM)); \
                if (dev >= 0 && dev < 64) fit_pipe[dev] = sms * per_sm > 0 ? sms * per_sm : 148;                       \
            }                                                                                                          \
            const int64_t fit = (dev >= 0 && dev < 64) ? fit_pipe[dev] : 148, need = cdiv(n_tiles, JN_WARPS);          \
            k_join_pipe<X, SL><<<(unsigned)(need < fit ? need : fit), JN_THREADS, JP_SMEM, st>>>(                       \
                *lists, *q, vec, out_q, out_s, s_label, out_capacity, state, ticket, n_tiles, out_total);              \
        } else IV_JK(X, SL, false);                                                                                    \
    } while (0)
    const bool any = mode == IV_MODE_ANY;
    if (narrow_span(lists)) {
        if (slack) IV_JK2(uint32_t, true); else IV_JK2(uint32_t, false);
    } else {
        if (slack) IV_JK2(uint64_t, true); else IV_JK2(uint64_t, false);
    }
#undef IV_JK2
#undef IV_JK
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
    static const int xsmem = getenv("IV_COUNT_SMEM") ? atoi(getenv("IV_COUNT_SMEM")) : 0;  // measurements only
#define IV_JC(X, M)                                                                                           \
    do {                                                                                                      \
        if (xsmem > 0) cudaFuncSetAttribute(k_join_count<X, M, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, xsmem); \
        if (slack) k_join_count<X, M, true><<<nb, JN_THREADS, 0, st>>>(*lists, *q, vec, out_count);            \
        else k_join_count<X, M, false><<<nb, JN_THREADS, xsmem, st>>>(*lists, *q, vec, out_count);             \
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
