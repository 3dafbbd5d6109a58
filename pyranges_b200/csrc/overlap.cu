// overlap.cu -- batched overlap queries against the flattened subject index.
//
// Replaces (reference call sites; the native code itself lives in the un-vendored ncls /
// sorted_nearest packages):
//   NCLS.all_overlaps_both / all_overlaps_self / has_overlaps / all_containments_both /
//   first_overlap_both / last_overlap_both   pyranges/methods/join.py:15-23, intersection.py:73-78, coverage.py:26
//   NCLS.set_difference_helper  pyranges/methods/subtraction.py:63-65
//   NCLS.coverage               pyranges/methods/coverage.py:64
//   nearest_previous/next/_nonoverlapping   pyranges/methods/nearest.py:59,69,113
//
// count   = #{S' < e'} - #{E' <= s'}         ONE 256-bit gather per query: the 32-byte record of the bin of
//                                            e' holds both ranks at the bin origin plus the bin's first four
//                                            start / end offsets (s'+1 almost always lies in the same bin)
// hits    = class A: s' <= S' < e'  (contiguous in start order, all of them overlap)
//         + class B: S' < s' < E'   (stabbing set) = members of the bin's stabbing list that end after s'
//                                    + earlier starts of the same bin that end after s'
// emit    = the count pass leaves, per 64-row granule, the list of its queries WITH hits (row, count, flattened
//           interval) and the granule's pair sum; tile sums -> exclusive scan -> the emit pass walks the dense
//           lists, one warp per quarter tile, and never reads the query columns or the counts again.
//
// Kernels are instantiated for 32-bit flattened coordinates (total span < 2^32: any single genome) and for
// 64-bit ones; the 32-bit instantiation halves the integer work per query.
#include "common.cuh"
#include "query.cuh"

namespace iv {

constexpr int OV_THREADS = 256;
constexpr int OV_ITEMS = 4;
constexpr int OV_TILE = OV_THREADS * OV_ITEMS;  // 1024 queries per block
constexpr int OV_GRAN = 64;                       // rows per emit warp: the count kernels also leave their sums
constexpr int OV_BATCH = 2;                       // record gathers a thread keeps in flight
constexpr int OV_GPT = OV_TILE / OV_GRAN;         // 16 granules per tile

// ---- hit lists: what the count pass hands to the emit pass ------------------------------------------------------
// One entry per query WITH hits (a fraction of the rows of a sparse join), written by the count kernels straight
// from their registers: the row inside its tile, the hit count and the flattened interval.  The entries of a 64-row
// granule sit at the front of the granule's 64 slots, in row order; gran_hits[] says how many.  The emit pass then
// works on dense lists -- every lane owns a hit query -- and never touches the query columns or the counts again.
template <typename X> struct HitEntry;
template <> struct __align__(16) HitEntry<uint32_t> { uint32_t row, cnt, sc, ec; };
template <> struct __align__(16) HitEntry<uint64_t> { uint32_t row, cnt; uint64_t sc, ec, pad; };
template <typename X> __device__ __forceinline__ void put_entry(HitEntry<X>* p, uint32_t row, uint32_t cnt, X sc, X ec);
template <> __device__ __forceinline__ void put_entry<uint32_t>(HitEntry<uint32_t>* p, uint32_t row, uint32_t cnt,
                                                                 uint32_t sc, uint32_t ec) {
    *reinterpret_cast<uint4*>(p) = make_uint4(row, cnt, sc, ec);
}
template <> __device__ __forceinline__ void put_entry<uint64_t>(HitEntry<uint64_t>* p, uint32_t row, uint32_t cnt,
                                                                 uint64_t sc, uint64_t ec) {
    uint4* v = reinterpret_cast<uint4*>(p);
    v[0] = make_uint4(row, cnt, (uint32_t)sc, (uint32_t)(sc >> 32));
    v[1] = make_uint4((uint32_t)ec, (uint32_t)(ec >> 32), 0u, 0u);
}

template <typename X> struct Hit { uint32_t row, cnt; X sc, ec; };
template <typename X> __device__ __forceinline__ Hit<X> get_entry(const HitEntry<X>* p);
template <> __device__ __forceinline__ Hit<uint32_t> get_entry<uint32_t>(const HitEntry<uint32_t>* p) {
    const uint4 v = __ldg(reinterpret_cast<const uint4*>(p));
    return Hit<uint32_t>{v.x, v.y, v.z, v.w};
}
template <> __device__ __forceinline__ Hit<uint64_t> get_entry<uint64_t>(const HitEntry<uint64_t>* p) {
    const uint4 a = __ldg(reinterpret_cast<const uint4*>(p)), b = __ldg(reinterpret_cast<const uint4*>(p) + 1);
    return Hit<uint64_t>{a.x, a.y, (uint64_t)a.z | ((uint64_t)a.w << 32), (uint64_t)b.x | ((uint64_t)b.y << 32)};
}

// ---- bin records ----------------------------------------------------------------------------------------------
// w[0]=bs w[1]=be w[2]=co w[3]=ns|ne<<8|nc<<16|flags<<24 w[4..5]=so[0..3] w[6..7]=eo[0..3]
struct Rec { uint32_t w[8]; };

// one 256-bit load = one L2 sector per look-up (LDG.E.256 on sm_100a)
__device__ __forceinline__ Rec load_rec(const iv_index_t& ix, uint64_t b) {
    Rec r;
    const iv_binrec_t* p = ix.bins + b;
    asm("ld.global.nc.v8.u32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
        : "=r"(r.w[0]), "=r"(r.w[1]), "=r"(r.w[2]), "=r"(r.w[3]), "=r"(r.w[4]), "=r"(r.w[5]), "=r"(r.w[6]), "=r"(r.w[7])
        : "l"(p));
    return r;
}
// #{15-bit offsets of (a, b) that are < rel}, rel <= 0x8000: per 16-bit half, bit 15 of (rel + 0x7fff - x) is set iff
// x < rel (no borrow between the halves: x <= 0x7fff <= rel + 0x7fff)
__device__ __forceinline__ uint32_t swar_cnt4(uint32_t a, uint32_t b, uint32_t rel) {
    const uint32_t c = (rel + 0x7fffu) * 0x10001u;
    return __popc(((c - a) & 0x80008000u) | (((c - b) & 0x80008000u) >> 1));
}
__device__ __forceinline__ uint32_t cnt4(uint32_t a, uint32_t b, uint32_t rel) { return swar_cnt4(a, b, rel); }
// lo + #{off[p] < rel : p in [lo, hi)}   (off ascending inside a bin)
__device__ __noinline__ uint32_t rank_in(const uint32_t* __restrict__ off, uint32_t lo, uint32_t hi, uint32_t rel) {
    if (hi - lo <= 32) {  // independent loads over at most a few sectors: one round trip, not log2(n) dependent ones
        uint32_t r = lo;
#pragma unroll 8
        for (uint32_t p = lo; p < hi; p++) r += __ldg(off + p) < rel ? 1u : 0u;
        return r;
    }
    while (lo < hi) {
        uint32_t mid = (lo + hi) >> 1;
        if (__ldg(off + mid) < rel) lo = mid + 1; else hi = mid;
    }
    return lo;
}
__device__ __forceinline__ bool rec_inline_s(const Rec& r) { return (r.w[3] & (IV_BIN_INLINE << 24)) && (r.w[3] & 0xffu) <= 4; }
__device__ __forceinline__ bool rec_inline_e(const Rec& r) { return (r.w[3] & (IV_BIN_INLINE << 24)) && ((r.w[3] >> 8) & 0xffu) <= 4; }
// #starts < origin(b) + rel,  #ends < origin(b) + rel   (r = record of bin b)
__device__ __forceinline__ uint32_t rec_rank_s(const iv_index_t& ix, const Rec& r, uint64_t b, uint32_t rel) {
    if (rec_inline_s(r)) return r.w[0] + cnt4(r.w[4], r.w[5], rel);
    const uint32_t n = r.w[3] & 0xffu;
    const uint32_t hi = n < 255 ? r.w[0] + n : __ldg(&ix.bins[b + 1].bs);
    return rank_in(ix.soff, r.w[0], hi, rel);
}
__device__ __forceinline__ uint32_t rec_rank_e(const iv_index_t& ix, const Rec& r, uint64_t b, uint32_t rel) {
    if (rec_inline_e(r)) return r.w[1] + cnt4(r.w[6], r.w[7], rel);
    const uint32_t n = (r.w[3] >> 8) & 0xffu;
    const uint32_t hi = n < 255 ? r.w[1] + n : __ldg(&ix.bins[b + 1].be);
    return rank_in(ix.eoff, r.w[1], hi, rel);
}
template <typename X> __device__ __forceinline__ uint32_t rel_of(const iv_index_t& ix, X x) {
    return (uint32_t)x & ((1u << ix.shift) - 1u);  // shift <= 31
}
// #starts < x, #ends < x  (x: flattened coordinate)
template <typename X> __device__ __forceinline__ uint32_t rank_s_lt(const iv_index_t& ix, X x) {
    const uint64_t b = x >> ix.shift;
    return rec_rank_s(ix, load_rec(ix, b), b, rel_of(ix, x));
}
template <typename X> __device__ __forceinline__ uint32_t rank_e_lt(const iv_index_t& ix, X x) {
    const uint64_t b = x >> ix.shift;
    return rec_rank_e(ix, load_rec(ix, b), b, rel_of(ix, x));
}

// ---- query geometry (query.cuh) -------------------------------------------------------------------------------------
template <typename X> __device__ __forceinline__ Flat<X> load_flat(const iv_index_t& ix, const iv_queries_t& q, int seg) {
    return load_flat_groups<X>(ix.groups, ix.n, q, seg);
}
// per-thread cache: most blocks sit inside one query segment
template <typename X> struct FlatCache {
    SegSpan sp;
    Flat<X> f0;
    __device__ __forceinline__ FlatCache(const iv_index_t& ix, const iv_queries_t& q, SegSpan s) : sp(s) {
        f0 = load_flat<X>(ix, q, sp.klo);
    }
    __device__ __forceinline__ Flat<X> of(const iv_index_t& ix, const iv_queries_t& q, int64_t i, int* seg_out = nullptr) const {
        if (sp.klo == sp.khi) { if (seg_out) *seg_out = sp.klo; return f0; }
        int seg = find_seg_range(q.seg_off, sp.klo, sp.khi, i);
        if (seg_out) *seg_out = seg;
        return seg == sp.klo ? f0 : load_flat<X>(ix, q, seg);
    }
};
// count from the record of bin(ec); general fallback when s'+1 sits in another bin or a bin overflows
template <typename X>
__device__ __forceinline__ int64_t count_rec(const iv_index_t& ix, X sc, X ec, const Rec& r, uint64_t b1) {
    const X xs = sc + 1;
    const uint64_t b0 = xs >> ix.shift;
    if (b0 == b1 && rec_inline_s(r) && rec_inline_e(r)) {
        int c = (int)(r.w[0] + cnt4(r.w[4], r.w[5], rel_of(ix, ec))) - (int)(r.w[1] + cnt4(r.w[6], r.w[7], rel_of(ix, xs)));
        return c > 0 ? c : 0;
    }
    int64_t pe = rec_rank_s(ix, r, b1, rel_of(ix, ec));
    int64_t re = b0 == b1 ? rec_rank_e(ix, r, b1, rel_of(ix, xs)) : rank_e_lt<X>(ix, xs);
    return pe > re ? pe - re : 0;
}
template <typename X> __device__ __forceinline__ int64_t count_all(const iv_index_t& ix, X sc, X ec) {
    const uint64_t b1 = ec >> ix.shift;
    return count_rec<X>(ix, sc, ec, load_rec(ix, b1), b1);
}

// Visit every subject overlapping [sc, ec); r = record of bin(sc).
// WANT_P: f(position in start order), else f(subject row).
template <typename X, bool WANT_P, typename F>
__device__ __forceinline__ void for_each_hit_rec(const iv_index_t& ix, X sc, X ec, const Rec& r, F&& f) {
    const uint64_t b = sc >> ix.shift;
    const uint32_t rel = rel_of(ix, sc);
    const uint32_t pS = rec_rank_s(ix, r, b, rel);
    const uint32_t pE = (ec >> ix.shift) == b ? rec_rank_s(ix, r, b, rel_of(ix, ec)) : rank_s_lt<X>(ix, ec);
    for (uint32_t p = pS; p < pE; p++) f(WANT_P ? p : __ldg(ix.row_s + p));  // class A
    uint32_t pL = r.w[0];  // class B, started in this bin: only starts later than s' - max_len can still be open
    if (pS - pL > 8 && ix.groups.max_len > 0 && (int64_t)rel >= ix.groups.max_len)
        pL = rank_in(ix.soff, pL, pS, (uint32_t)((int64_t)rel - ix.groups.max_len + 1));
    for (uint32_t p = pL; p < pS; p++)
        if (__ldg(ix.seoff + p) > rel) f(WANT_P ? p : __ldg(ix.row_s + p));
    const uint32_t nc = (r.w[3] >> 16) & 0xffu;
    const uint32_t c0 = r.w[2], c1 = nc < 255 ? c0 + nc : __ldg(&ix.bins[b + 1].co);
    for (uint32_t j = c0; j < c1; j++) {                                      // class B, open at the bin origin
        const uint2 en = __ldg(reinterpret_cast<const uint2*>(ix.cross) + j);
        if (en.y > rel) f(WANT_P ? __ldg(ix.cross_p + j) : en.x);
    }
}
template <typename X, bool WANT_P, typename F>
__device__ __forceinline__ void for_each_hit(const iv_index_t& ix, X sc, X ec, F&& f) {
    for_each_hit_rec<X, WANT_P>(ix, sc, ec, load_rec(ix, sc >> ix.shift), f);
}

// first hit in nested-containment-list order: smallest start, then largest end, then smallest row --
// found without visiting every hit: in start order the stabbing-list members (started before the bin)
// come before the bin's earlier starts, and those before the starts inside the query; inside each class the smallest
// start-order position has the smallest start (and, among equal starts and ends, the smallest row: the sort is
// stable).  Only a run of EQUAL starts at the winning position needs the ends compared.  r = record of bin(sc); the
// query must have at least one hit.  -> subject row.
template <typename X>
__device__ __forceinline__ uint32_t first_hit_rec(const iv_index_t& ix, X sc, X ec, const Rec& r) {
    const uint64_t b = sc >> ix.shift;
    const uint32_t rel = rel_of(ix, sc);
    uint32_t lo = 0xffffffffu, hi = 0;  // candidate positions [lo, hi) of the winning class, start order
    // class B, open at the bin origin
    const uint32_t nc = (r.w[3] >> 16) & 0xffu;
    const uint32_t c0 = r.w[2], c1 = nc < 255 ? c0 + nc : __ldg(&ix.bins[b + 1].co);
    for (uint32_t j = c0; j < c1; j++) {
        const uint2 en = __ldg(reinterpret_cast<const uint2*>(ix.cross) + j);
        if (en.y > rel) {
            const uint32_t p = __ldg(ix.cross_p + j);
            lo = p < lo ? p : lo;
        }
    }
    if (lo != 0xffffffffu) {
        hi = r.w[0];  // members of the list all start before the bin: positions below the bin's first start
    } else {
        const uint32_t pS = rec_rank_s(ix, r, b, rel);
        uint32_t pL = r.w[0];
        if (pS - pL > 8 && ix.groups.max_len > 0 && (int64_t)rel >= ix.groups.max_len)
            pL = rank_in(ix.soff, pL, pS, (uint32_t)((int64_t)rel - ix.groups.max_len + 1));
        for (uint32_t p = pL; p < pS; p++)  // class B, started in this bin
            if (__ldg(ix.seoff + p) > rel) { lo = p; hi = pS; break; }
        if (lo == 0xffffffffu) {  // class A: the first start inside the query
            lo = pS;
            hi = (ec >> ix.shift) == b ? rec_rank_s(ix, r, b, rel_of(ix, ec)) : rank_s_lt<X>(ix, ec);
        }
    }
    // ties: the run of starts equal to key_s[lo] inside [lo, hi); every member of the run that is a hit competes on
    // (largest end, smallest position)
    uint32_t best = lo;
    if (lo + 1 < hi) {
        const uint64_t s0 = __ldg(ix.key_s + lo);
        if (__ldg(ix.key_s + lo + 1) == s0) {
            uint64_t be = __ldg(ix.end_s + lo);
            for (uint32_t p = lo + 1; p < hi && __ldg(ix.key_s + p) == s0; p++) {
                const uint64_t e = __ldg(ix.end_s + p);
                if (e > (uint64_t)sc && e > be) { be = e; best = p; }  // e > sc: it is a hit (its start is < ec already)
            }
        }
    }
    return __ldg(ix.row_s + best);
}

// NCLS.last_overlap_both: the hit that ENDS LAST (ties: smallest start, then smallest row).  ncls keeps the hit with the
// largest end while it walks the nested lists; the reference's own known answer pins it (tests/unit/k_nearest.py:163-211:
// the "last" overlap of [15, 20) among [11, 16) and [11, 20) is [11, 20)).  Every hit is visited: `last` is not a hot mode.
template <typename X>
__device__ __forceinline__ uint32_t last_hit_rec(const iv_index_t& ix, X sc, X ec, const Rec& r) {
    uint64_t be = 0, bs = 0;
    uint32_t brow = 0xffffffffu;
    for_each_hit_rec<X, true>(ix, sc, ec, r, [&](uint32_t p) {
        const uint64_t e = __ldg(ix.end_s + p), st = __ldg(ix.key_s + p);
        const uint32_t row = __ldg(ix.row_s + p);
        if (brow == 0xffffffffu || e > be || (e == be && (st < bs || (st == bs && row < brow)))) {
            be = e;
            bs = st;
            brow = row;
        }
    });
    return brow;
}

__device__ __forceinline__ void load_items(const int64_t* __restrict__ p, int64_t i0, int64_t n, bool vec, int64_t (&v)[OV_ITEMS]) {
    if (vec && i0 + OV_ITEMS <= n) {
        const longlong2* p2 = reinterpret_cast<const longlong2*>(p + i0);
        longlong2 a = __ldg(p2), b = __ldg(p2 + 1);
        v[0] = a.x; v[1] = a.y; v[2] = b.x; v[3] = b.y;
    } else {
#pragma unroll
        for (int j = 0; j < OV_ITEMS; j++) v[j] = (i0 + j < n) ? __ldg(p + i0 + j) : 0;
    }
}
__device__ __forceinline__ void store_items(int64_t* __restrict__ p, int64_t i0, int64_t n, bool vec, const int64_t (&v)[OV_ITEMS]) {
    if (vec && i0 + OV_ITEMS <= n) {
        longlong2* p2 = reinterpret_cast<longlong2*>(p + i0);
        p2[0] = make_longlong2(v[0], v[1]);
        p2[1] = make_longlong2(v[2], v[3]);
    } else {
#pragma unroll
        for (int j = 0; j < OV_ITEMS; j++) if (i0 + j < n) p[i0 + j] = v[j];
    }
}

// ---- count ----------------------------------------------------------------------------------------------------------
template <typename X, int MODE>
__global__ void __launch_bounds__(OV_THREADS)
k_count(iv_index_t ix, iv_queries_t q, bool vec, int64_t* __restrict__ out_count, int64_t* __restrict__ gran_sums,
        uint32_t* __restrict__ gran_hits, HitEntry<X>* __restrict__ list, int64_t first_tile) {
    __shared__ int span[2];
    const int64_t tile = first_tile + blockIdx.x;
    const int64_t tbase = tile * OV_TILE;
    const int64_t last = tbase + OV_TILE - 1 < q.n - 1 ? tbase + OV_TILE - 1 : q.n - 1;
    const FlatCache<X> fc(ix, q, block_seg_span(q.seg_off, q.n_seg, tbase, last, span));
    const int64_t i0 = tbase + (int64_t)threadIdx.x * OV_ITEMS;
    int64_t s[OV_ITEMS], e[OV_ITEMS], c[OV_ITEMS];
    load_items(q.start, i0, q.n, vec, s);
    load_items(q.end, i0, q.n, vec, e);
    // two half-batches: flatten, issue one record gather per query (all in flight together), then count;
    // dead rows use a harmless geometry
    int64_t tsum = 0;
    X fsc[OV_ITEMS], fec[OV_ITEMS];  // flattened intervals, kept for the hit list
#pragma unroll
    for (int h = 0; h < OV_ITEMS; h += OV_BATCH) {
        X sc[OV_BATCH], ec[OV_BATCH];
        Rec r[OV_BATCH];
        bool out[OV_BATCH];
#pragma unroll
        for (int k = 0; k < OV_BATCH; k++) {
            const int j = h + k;
            const int64_t i = i0 + j < q.n ? i0 + j : q.n - 1;
            const Flat<X> f = fc.of(ix, q, i);
            apply_slack(q.slack, s[j], e[j]);
            sc[k] = f.at(s[j]);
            ec[k] = f.at(e[j]);
            fsc[j] = sc[k];
            fec[j] = ec[k];
            out[k] = MODE == IV_MODE_CONTAIN ? (f.g < 0 || f.outside(s[j]) || f.outside(e[j])) : false;
            r[k] = load_rec(ix, ec[k] >> ix.shift);
        }
#pragma unroll
        for (int k = 0; k < OV_BATCH; k++) {
            const int j = h + k;
            const int64_t ca = count_rec<X>(ix, sc[k], ec[k], r[k], ec[k] >> ix.shift);
            if (MODE == IV_MODE_ALL) c[j] = ca;
            else if (MODE == IV_MODE_ANY || MODE == IV_MODE_LAST) c[j] = ca > 0 ? 1 : 0;
            else {
                // containment: subject.start <= q.start && q.end <= subject.end (raw coordinates)
                int64_t cnt = 0;
                if (ca > 0 && !out[k]) {
                    const uint64_t qs = sc[k], qe = ec[k];
                    for_each_hit<X, true>(ix, sc[k], ec[k], [&](uint32_t p) {
                        cnt += (__ldg(ix.key_s + p) <= qs && qe <= __ldg(ix.end_s + p)) ? 1 : 0;
                    });
                }
                c[j] = cnt;
            }
            if (i0 + j >= q.n) c[j] = 0;
            tsum += c[j];
        }
    }
    if (out_count) store_items(out_count, i0, q.n, vec, c);
    if (list) {  // a granule = the 64 rows of 16 consecutive threads: its hit entries, in row order
        const int nh = (c[0] > 0) + (c[1] > 0) + (c[2] > 0) + (c[3] > 0);
        int inc = nh;
#pragma unroll
        for (int o = 1; o < 16; o <<= 1) {
            const int t = __shfl_up_sync(FULL, inc, o, 16);
            if ((lane_id() & 15) >= o) inc += t;
        }
        const int64_t gran = tile * OV_GPT + (threadIdx.x >> 4);
        HitEntry<X>* slot = list + gran * OV_GRAN + (inc - nh);
#pragma unroll
        for (int j = 0; j < OV_ITEMS; j++)
            if (c[j] > 0) put_entry<X>(slot++, (uint32_t)(threadIdx.x * OV_ITEMS + j), (uint32_t)c[j], fsc[j], fec[j]);
        int64_t gs = tsum;
#pragma unroll
        for (int o = 1; o < 16; o <<= 1) gs += __shfl_xor_sync(FULL, gs, o);
        if ((lane_id() & 15) == 15 && gran * OV_GRAN < q.n) {
            gran_hits[gran] = (uint32_t)inc;
            gran_sums[gran] = gs;
        }
    }
}

// tile_sums[t] = sum of the tile's 16 granule sums (granules beyond the last row do not exist: 0)
__global__ void __launch_bounds__(256)
k_tile_sums(const int64_t* __restrict__ gran_sums, int64_t n, int64_t ntiles, int64_t* __restrict__ tile_sums) {
    const int64_t g = (int64_t)blockIdx.x * 256 + threadIdx.x;  // 16 consecutive lanes = one tile
    int64_t v = (g < ntiles * OV_GPT && g * OV_GRAN < n) ? __ldg(gran_sums + g) : 0;
#pragma unroll
    for (int o = 1; o < OV_GPT; o <<= 1) v += __shfl_xor_sync(FULL, v, o);
    if ((threadIdx.x & (OV_GPT - 1)) == 0 && g / OV_GPT < ntiles) tile_sums[g / OV_GPT] = v;
}

// ---- count, lean: the throughput kernel for full, 16-byte aligned tiles ----------------------------------------------
// The record gather runs at ~1 sector / cycle / SM whatever its width (tools/gather_bench.cu: 253 G gathers/s),
// so the budget per query is set by INSTRUCTIONS: this kernel keeps them to a few dozen -- 32-bit clamps when the
// coordinates fit, one 256-bit record gather, SWAR compares of the inline 15-bit offsets, vector loads / stores --
// and keeps 48+ warps per SM resident.  The rare queries that need more than the one record (s'+1 in another bin, a
// bin with more than four starts / ends) go to a per-warp queue in shared memory and are counted 32 at a time by
// the full warp, so the long path never runs with one or two active lanes.
constexpr int LN_THREADS = 256;
constexpr int LN_TILES = 4;                        // consecutive 1024-row tiles per block (ws_grans() rounds to it)
constexpr int LN_QCAP = 64;                        // slow queue entries per warp

__device__ __forceinline__ bool rec_fast(const Rec& r) {
    // inline flag set, at most four starts and four ends
    return (r.w[3] & (IV_BIN_INLINE << 24)) && (r.w[3] & 0xffu) <= 4 && ((r.w[3] >> 8) & 0xffu) <= 4;
}

template <typename X, int MODE, bool SLACK>
__global__ void __launch_bounds__(LN_THREADS, 6)
k_count_lean(iv_index_t ix, iv_queries_t q, int64_t* __restrict__ out_count, unsigned long long* __restrict__ gran_sums,
             uint32_t* __restrict__ gran_hits, HitEntry<X>* __restrict__ list) {
    __shared__ X q_sc[LN_THREADS / 32][LN_QCAP], q_ec[LN_THREADS / 32][LN_QCAP];
    __shared__ uint32_t q_row[LN_THREADS / 32][LN_QCAP], q_ent[LN_THREADS / 32][LN_QCAP];
    constexpr uint32_t BLOCK_ROWS = LN_TILES * OV_TILE;
    const uint32_t warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int64_t bbase = (int64_t)blockIdx.x * BLOCK_ROWS;
    // everything below is indexed relative to the block's first row, in 32 bits
    const uint32_t nrem = (uint32_t)(q.n - bbase < (int64_t)BLOCK_ROWS ? q.n - bbase : (int64_t)BLOCK_ROWS);
    const int64_t* __restrict__ qs = q.start + bbase;
    const int64_t* __restrict__ qe = q.end + bbase;
    int64_t* __restrict__ oc = out_count ? out_count + bbase : nullptr;
    HitEntry<X>* __restrict__ lb = list ? list + bbase : nullptr;
    uint32_t* __restrict__ gh = gran_hits + bbase / OV_GRAN;
    unsigned long long* __restrict__ gs = gran_sums + bbase / OV_GRAN;
    const int shift = ix.shift;
    const uint32_t mask = (1u << shift) - 1u;
    const unsigned lt = (1u << lane) - 1u;
    uint32_t q_head = 0, q_tail = 0;  // warp-uniform ring positions of the slow queue
    auto count_slow = [&](uint32_t n_valid) {
        if (lane < n_valid) {
            const uint32_t at = (q_head + lane) % LN_QCAP;
            const uint32_t row = q_row[warp][at];
            int64_t c = count_all<X>(ix, q_sc[warp][at], q_ec[warp][at]);
            if (MODE == IV_MODE_ANY) c = c > 0 ? 1 : 0;
            if (oc) oc[row] = c;
            if (c && lb) {
                // the granule's sum was stored (plain) by lane 0 of this warp when the row was queued
                atomicAdd(gs + row / OV_GRAN, (unsigned long long)c);
                // the entry was reserved (count 0) when the row was queued: fill in the count
                lb[q_ent[warp][at]].cnt = (uint32_t)c;
            }
        }
        q_head += n_valid;
        __syncwarp();
    };
    // segment geometry: most blocks lie inside one key; otherwise it is re-derived per half-tile
    int seg = find_seg(q.seg_off, q.n_seg, bbase), f_seg = -1;
    const bool blk_uniform = __ldg(q.seg_off + seg + 1) >= bbase + (int64_t)BLOCK_ROWS;
    Flat<X> f0;
    uint32_t lo32 = 0, hi32 = 0, off32 = 0;
    bool geo32 = false;
    for (uint32_t hrel = 0; hrel < BLOCK_ROWS; hrel += OV_TILE / 2) {  // 512 rows: 2 per thread
        if (hrel >= nrem) break;  // ragged last block
        bool uniform = true;
        if (!blk_uniform) {
            const int64_t hbase = bbase + hrel;
            while (seg + 1 < q.n_seg && __ldg(q.seg_off + seg + 1) <= hbase) seg++;
            uniform = __ldg(q.seg_off + seg + 1) >= hbase + OV_TILE / 2;
        }
        if (seg != f_seg) {
            f0 = load_flat<X>(ix, q, seg);
            f_seg = seg;
            const uint64_t hi = (uint64_t)f0.lo + f0.width;
            geo32 = sizeof(X) == 4 && f0.lo >= 0 && hi <= 0xffffffffull;
            lo32 = (uint32_t)f0.lo;
            hi32 = (uint32_t)hi;
            off32 = (uint32_t)f0.base - lo32;
        }
        const uint32_t r0 = hrel + 2 * threadIdx.x;  // rows r0, r0 + 1 of the block
        int64_t s[2] = {0, 0}, e[2] = {0, 0};
        if (r0 + 1 < nrem) {
            const longlong2 s2 = __ldg(reinterpret_cast<const longlong2*>(qs + r0));
            const longlong2 e2 = __ldg(reinterpret_cast<const longlong2*>(qe + r0));
            s[0] = s2.x; s[1] = s2.y; e[0] = e2.x; e[1] = e2.y;
        } else if (r0 < nrem) {
            s[0] = __ldg(qs + r0); e[0] = __ldg(qe + r0);
        }
        X sc[2], ec[2];
        Rec r[2];
#pragma unroll
        for (int k = 0; k < 2; k++) {
            if (SLACK) apply_slack(q.slack, s[k], e[k]);  // join(slack=k): its own instantiation
            if (uniform && geo32 && (((uint64_t)s[k] | (uint64_t)e[k]) >> 32) == 0) {
                // x' = clamp(x, lo, hi) + (base - lo), all in 32 bits
                sc[k] = (X)(min(max((uint32_t)s[k], lo32), hi32) + off32);
                ec[k] = (X)(min(max((uint32_t)e[k], lo32), hi32) + off32);
            } else {
                Flat<X> f = f0;
                if (!uniform) {
                    int sg = find_seg_range(q.seg_off, seg, q.n_seg - 1, bbase + r0 + k);
                    if (sg != seg) f = load_flat<X>(ix, q, sg);
                }
                sc[k] = f.at(s[k]);
                ec[k] = f.at(e[k]);
            }
            r[k] = load_rec(ix, ec[k] >> shift);
        }
        int cc[2];
        bool fast[2];
#pragma unroll
        for (int k = 0; k < 2; k++) {
            const X xs = sc[k] + 1;
            const bool dead = r0 + k >= nrem;  // beyond the ragged end: count 0, never queued
            fast[k] = dead || ((xs >> shift) == (ec[k] >> shift) && rec_fast(r[k]));
            int c = (int)(r[k].w[0] - r[k].w[1]) + (int)swar_cnt4(r[k].w[4], r[k].w[5], (uint32_t)ec[k] & mask) -
                    (int)swar_cnt4(r[k].w[6], r[k].w[7], (uint32_t)xs & mask);
            c = c > 0 ? c : 0;
            if (MODE == IV_MODE_ANY) c = c > 0 ? 1 : 0;
            cc[k] = (fast[k] && !dead) ? c : 0;
        }
        // rows r0, r0+1: one 16-byte store; queued rows are rewritten by count_slow (same warp, later)
        if (oc) {
            if (r0 + 1 < nrem) *reinterpret_cast<longlong2*>(oc + r0) = make_longlong2(cc[0], cc[1]);
            else if (r0 < nrem) oc[r0] = cc[0];
        }
        // this warp's 64 rows are one granule of the hit list: an entry for every row with hits, and a
        // reserved one (count 0 for now) for every row that goes to the slow queue
        const unsigned slow0 = __ballot_sync(FULL, !fast[0]), slow1 = __ballot_sync(FULL, !fast[1]);
        uint32_t ent[2] = {0, 0};
        if (lb) {
            const bool in0 = cc[0] > 0 || !fast[0], in1 = cc[1] > 0 || !fast[1];
            const unsigned m0 = __ballot_sync(FULL, in0), m1 = __ballot_sync(FULL, in1);
            const uint32_t gran = hrel / OV_GRAN + warp;  // granule of the block
            ent[0] = gran * OV_GRAN + __popc(m0 & lt) + __popc(m1 & lt);
            ent[1] = ent[0] + (in0 ? 1u : 0u);
            const uint32_t trow = r0 & (OV_TILE - 1);  // row inside its 1024-row tile
            if (in0) put_entry<X>(lb + ent[0], trow, (uint32_t)cc[0], sc[0], ec[0]);
            if (in1) put_entry<X>(lb + ent[1], trow + 1, (uint32_t)cc[1], sc[1], ec[1]);
            // pair sum of the granule's fast rows (plain store: the warp owns the granule; queued rows add theirs
            // later, from this same warp, with atomics)
            const unsigned long long g = warp_sum((unsigned long long)(unsigned)(cc[0] + cc[1]));
            if (lane == 0) {
                gh[gran] = __popc(m0) + __popc(m1);
                gs[gran] = g;
            }
        }
#pragma unroll
        for (int k = 0; k < 2; k++) {
            const unsigned m = k ? slow1 : slow0;
            if (m) {  // queue the others (flattened coordinates + row); the whole warp counts them 32 at a time
                if (!fast[k]) {
                    const uint32_t at = (q_tail + __popc(m & lt)) % LN_QCAP;
                    q_sc[warp][at] = sc[k];
                    q_ec[warp][at] = ec[k];
                    q_row[warp][at] = r0 + k;
                    q_ent[warp][at] = ent[k];
                }
                q_tail += __popc(m);
                __syncwarp();
                if (q_tail - q_head >= 32) count_slow(32);
            }
        }
    }
    if (q_tail != q_head) count_slow(q_tail - q_head);
}

// ---- emit -----------------------------------------------------------------------------------------------------------
// One WARP per quarter tile (four 64-row granules), no block-wide barrier.  The warp reads its granules' hit lists
// (left by the count pass) as one dense sequence, so every lane owns a query WITH hits; its output offset is the
// tile offset (scan of the tile sums) + the pair sums of the granules before it in the tile + a shuffle scan of the
// entry counts.  Granules without hits cost two loads.
constexpr int EM_GW = 4;                          // granules per warp
constexpr int EM_WARPS = OV_GPT / EM_GW;          // one block = one tile
constexpr int EM_THREADS = EM_WARPS * 32;

template <typename X, int MODE>
__global__ void __launch_bounds__(EM_THREADS, MODE == IV_MODE_ALL ? 16 : 1)
k_emit(iv_index_t ix, iv_queries_t q, const HitEntry<X>* __restrict__ list, const uint32_t* __restrict__ gran_hits,
       const int64_t* __restrict__ gran_sums, const int64_t* __restrict__ tile_off, int64_t* __restrict__ out_q,
       int64_t* __restrict__ out_s, const int64_t* __restrict__ s_label, int64_t cap) {
    const int lane = threadIdx.x & 31, quarter = threadIdx.x >> 5;
    const int64_t tile = blockIdx.x;
    const int64_t tbase = tile * OV_TILE;
    // lanes 0..15: the tile's granules
    const int64_t gi = tile * OV_GPT + (lane & (OV_GPT - 1));
    const bool gvalid = lane < OV_GPT && gi * OV_GRAN < q.n;
    const uint32_t gh = gvalid ? __ldg(gran_hits + gi) : 0u;
    int64_t carry = __ldg(tile_off + tile);
    uint32_t h[EM_GW], nh = 0;
#pragma unroll
    for (int j = 0; j < EM_GW; j++) {
        h[j] = __shfl_sync(FULL, gh, quarter * EM_GW + j);
        nh += h[j];
    }
    if (nh == 0) return;
    carry += warp_sum((gvalid && lane < quarter * EM_GW) ? __ldg(gran_sums + gi) : (int64_t)0);
    const HitEntry<X>* base = list + (tile * OV_GPT + quarter * EM_GW) * OV_GRAN;
    const int shift = ix.shift;
    const uint32_t mask = (1u << shift) - 1u;
    for (uint32_t kbase = 0; kbase < nh; kbase += 32) {
        uint32_t k = kbase + lane;
        Hit<X> hit{0u, 0u, 0, 0};
        if (k < nh) {
            int j = 0;  // which of the warp's granule lists holds entry k
#pragma unroll
            for (int u = 0; u < EM_GW - 1; u++)
                if (j == u && k >= h[u]) { k -= h[u]; j = u + 1; }
            hit = get_entry<X>(base + j * OV_GRAN + k);
        }
        const int64_t inc = warp_incl_sum((int64_t)hit.cnt);
        int64_t o = carry + inc - hit.cnt;
        carry += __shfl_sync(FULL, inc, 31);
        // a query whose pairs do not all fit the caller's buffers is dropped as a whole (speculative emission: the
        // caller compares the total with its capacity and repeats the call)
        if (hit.cnt == 0 || o + hit.cnt > cap) continue;
        const int64_t row = tbase + hit.row;
        const X sc = hit.sc, ec = hit.ec;
        const uint64_t b = sc >> shift;
        const uint32_t rel = (uint32_t)sc & mask;
        const Rec r = load_rec(ix, b);
        const int64_t ql = q.label ? __ldg(q.label + row) : row;
        if (MODE == IV_MODE_ALL) {
            int64_t* __restrict__ qp = out_q ? out_q + o : nullptr;  // this query's slice of the outputs
            int64_t* __restrict__ sp = out_s ? out_s + o : nullptr;
            uint32_t j = 0;
            auto put = [&](uint32_t srow) {
                if (qp) qp[j] = ql;
                if (sp) sp[j] = s_label ? __ldg(s_label + srow) : (int64_t)srow;
                j++;
            };
            // candidates in start order: [pL, pS) started in this bin before s' (class B if still open),
            // [pS, pE) start inside the query (class A); then the bin's stabbing list (class B, open at the bin
            // origin).
            uint32_t pL = r.w[0], pS, pE;
            const bool same = (ec >> shift) == b;
            const uint32_t ns = r.w[3] & 0xffu;
            if (rec_inline_s(r)) {
                pS = r.w[0] + swar_cnt4(r.w[4], r.w[5], rel);
                pE = same ? r.w[0] + swar_cnt4(r.w[4], r.w[5], (uint32_t)ec & mask) : rank_s_lt<X>(ix, ec);
            } else if (ns <= 32) {
                // more than four starts in the bin: ONE pass over its (start, end) offsets does both rank searches
                // and the candidate tests
                const uint32_t erel = same ? ((uint32_t)ec & mask) : 0xffffffffu, hi = pL + ns;
                for (uint32_t p = pL; p < hi; p += 4) {
                    uint32_t so[4], se[4];
#pragma unroll
                    for (int u = 0; u < 4; u++) so[u] = p + u < hi ? __ldg(ix.soff + p + u) : 0xffffffffu;
#pragma unroll
                    for (int u = 0; u < 4; u++) se[u] = p + u < hi ? __ldg(ix.seoff + p + u) : 0u;
#pragma unroll
                    for (int u = 0; u < 4; u++)
                        if (so[u] < rel ? se[u] > rel : so[u] < erel) put(__ldg(ix.row_s + p + u));
                }
                pL = pS = hi;                                // nothing left in this bin
                pE = same ? hi : rank_s_lt<X>(ix, ec);       // class A continues in the following bins
            } else {
                pS = rec_rank_s(ix, r, b, rel);
                pE = same ? rec_rank_s(ix, r, b, (uint32_t)ec & mask) : rank_s_lt<X>(ix, ec);
                if (pS - pL > 8 && ix.groups.max_len > 0 && (int64_t)rel >= ix.groups.max_len)
                    pL = rank_in(ix.soff, pL, pS, (uint32_t)((int64_t)rel - ix.groups.max_len + 1));
            }
            // four at a time, and the first chunk of every list is requested before any is consumed: one gather
            // latency per query instead of one per list
            const uint32_t nc = (r.w[3] >> 16) & 0xffu;
            const uint32_t x0 = r.w[2], x1 = nc < 255 ? x0 + nc : __ldg(&ix.bins[b + 1].co);
            const uint2* __restrict__ cross = reinterpret_cast<const uint2*>(ix.cross);
            uint32_t so[4], rw[4];
            uint2 en[4];
#pragma unroll
            for (int u = 0; u < 4; u++) so[u] = pL + u < pS ? __ldg(ix.seoff + pL + u) : 0u;
#pragma unroll
            for (int u = 0; u < 4; u++) en[u] = x0 + u < x1 ? __ldg(cross + x0 + u) : make_uint2(0u, 0u);
#pragma unroll
            for (int u = 0; u < 4; u++) rw[u] = pS + u < pE ? __ldg(ix.row_s + pS + u) : 0u;
            for (uint32_t p = pL; p < pS; p += 4) {  // started before s': open iff its end offset is beyond s'
                if (p != pL) {
#pragma unroll
                    for (int u = 0; u < 4; u++) so[u] = p + u < pS ? __ldg(ix.seoff + p + u) : 0u;
                }
#pragma unroll
                for (int u = 0; u < 4; u++)
                    if (so[u] > rel) put(__ldg(ix.row_s + p + u));
            }
            for (uint32_t p = pS; p < pE; p += 4) {  // start inside the query: all hits
                if (p != pS) {
#pragma unroll
                    for (int u = 0; u < 4; u++) rw[u] = p + u < pE ? __ldg(ix.row_s + p + u) : 0u;
                }
#pragma unroll
                for (int u = 0; u < 4; u++)
                    if (p + u < pE) put(rw[u]);
            }
            for (uint32_t j = x0; j < x1; j += 4) {
                if (j != x0) {
#pragma unroll
                    for (int u = 0; u < 4; u++) en[u] = j + u < x1 ? __ldg(cross + j + u) : make_uint2(0u, 0u);
                }
#pragma unroll
                for (int u = 0; u < 4; u++)
                    if (en[u].y > rel) put(en[u].x);
            }
        } else if (MODE == IV_MODE_ANY || MODE == IV_MODE_LAST) {
            if (out_q) out_q[o] = ql;
            if (out_s) {
                const int64_t srow = MODE == IV_MODE_ANY ? first_hit_rec<X>(ix, sc, ec, r) : last_hit_rec<X>(ix, sc, ec, r);
                out_s[o] = s_label ? __ldg(s_label + srow) : srow;
            }
        } else {
            const uint64_t qs = sc, qe = ec;
            for_each_hit_rec<X, true>(ix, sc, ec, r, [&](uint32_t p) {
                if (__ldg(ix.key_s + p) <= qs && qe <= __ldg(ix.end_s + p)) {
                    if (out_q) out_q[o] = ql;
                    if (out_s) { int64_t srow = __ldg(ix.row_s + p); out_s[o] = s_label ? __ldg(s_label + srow) : srow; }
                    o++;
                }
            });
        }
    }
}

// ---- NCLS.coverage -----------------------------------------------------------------------------------------------
template <typename X>
__global__ void __launch_bounds__(OV_THREADS, 6)
k_coverage(iv_index_t ix, iv_queries_t q, int64_t* __restrict__ out_bp, double* __restrict__ out_frac) {
    __shared__ int span[2];
    const int64_t tbase = (int64_t)blockIdx.x * OV_THREADS;
    const int64_t last = tbase + OV_THREADS - 1 < q.n - 1 ? tbase + OV_THREADS - 1 : q.n - 1;
    const FlatCache<X> fc(ix, q, block_seg_span(q.seg_off, q.n_seg, tbase, last, span));
    const int64_t i = tbase + threadIdx.x;
    if (i >= q.n) return;
    const int64_t s0 = __ldg(q.start + i), e0 = __ldg(q.end + i);
    int64_t s = s0, e = e0;
    apply_slack(q.slack, s, e);
    const Flat<X> f = fc.of(ix, q, i);
    const X sc = f.at(s), ec = f.at(e);
    int64_t bp = 0;
    if (count_all<X>(ix, sc, ec) > 0) {
        const uint64_t qs = sc, qe = ec;
        for_each_hit<X, true>(ix, sc, ec, [&](uint32_t p) {
            uint64_t S = __ldg(ix.key_s + p), E = __ldg(ix.end_s + p);
            bp += (int64_t)((E < qe ? E : qe) - (S > qs ? S : qs));
        });
    }
    if (out_bp) out_bp[i] = bp;
    if (out_frac) {
        int64_t len = e0 - s0;
        out_frac[i] = len == 0 ? 0.0 : (double)bp / (double)len;
    }
}

// ---- nearest --------------------------------------------------------------------------------------------------------
template <typename X>
__global__ void __launch_bounds__(OV_THREADS, 6)
k_nearest(iv_index_t ix, iv_queries_t q, int overlap, int64_t* __restrict__ out_row, int64_t* __restrict__ out_dist) {
    __shared__ int span[2];
    const int64_t tbase = (int64_t)blockIdx.x * OV_THREADS;
    const int64_t last = tbase + OV_THREADS - 1 < q.n - 1 ? tbase + OV_THREADS - 1 : q.n - 1;
    const FlatCache<X> fc(ix, q, block_seg_span(q.seg_off, q.n_seg, tbase, last, span));
    const int64_t i = tbase + threadIdx.x;
    if (i >= q.n) return;
    int seg;
    const Flat<X> f = fc.of(ix, q, i, &seg);
    int64_t s = __ldg(q.start + i), e = __ldg(q.end + i);
    apply_slack(q.slack, s, e);
    int64_t row = -1, dist = -1;
    if (f.g >= 0) {
        const int mode = q.seg_mode ? __ldg(q.seg_mode + seg) : IV_NEAREST_BOTH;
        const X sc = f.at(s), ec = f.at(e);
        // the record of bin(e') serves the overlap count, the rank of the next start and, when s'+1 shares the
        // bin (the usual case), the rank of the previous end and the first-hit walk as well
        const uint64_t b1 = ec >> ix.shift;
        const Rec r1 = load_rec(ix, b1);
        bool done = false;
        if (overlap && count_rec<X>(ix, sc, ec, r1, b1) > 0) {
            row = first_hit_rec<X>(ix, sc, ec, (uint64_t)(sc >> ix.shift) == b1 ? r1 : load_rec(ix, sc >> ix.shift));
            dist = 0; done = true;
        }
        if (!done) {
            const int64_t g_first = __ldg(ix.groups.row_off + f.g), g_end = __ldg(ix.groups.row_off + f.g + 1);
            int64_t prow = -1, pdist = -1, nrow = -1, ndist = -1;
            if (mode != IV_NEAREST_NEXT && s >= f.lo) {
                // last subject end <= s ; ties: first of the equal run (earliest row, stable sort)
                const X xs = (X)(sc + 1);
                const int64_t r = (uint64_t)(xs >> ix.shift) == b1 ? rec_rank_e(ix, r1, b1, rel_of(ix, xs)) : rank_e_lt<X>(ix, xs);
                if (r > g_first) {
                    const uint64_t ev = __ldg(ix.key_e + (r - 1));
                    int64_t jf = r - 1;
                    if (r - 2 >= g_first && __ldg(ix.key_e + (r - 2)) == ev) jf = rank_e_lt<X>(ix, (X)ev);  // duplicate ends
                    prow = __ldg(ix.row_e + jf);
                    pdist = s - f.raw(ev) + 1;
                }
            }
            if (mode != IV_NEAREST_PREVIOUS && !(e > f.lo && (uint64_t)(e - f.lo) > f.width)) {
                const int64_t j = rec_rank_s(ix, r1, b1, rel_of(ix, ec));
                if (j < g_end) {
                    nrow = __ldg(ix.row_s + j);
                    ndist = f.raw(__ldg(ix.key_s + j)) - e + 1;
                }
            }
            if (prow >= 0 && (nrow < 0 || pdist <= ndist)) { row = prow; dist = pdist; }
            else if (nrow >= 0) { row = nrow; dist = ndist; }
        }
    }
    out_row[i] = row;
    out_dist[i] = dist;
}

// ---- k-nearest ------------------------------------------------------------------------------------------------------
// sorted_nearest.k_nearest_previous_nonoverlapping / k_nearest_next_nonoverlapping as pyranges calls them
// (pyranges/methods/k_nearest.py:6-48): for every query the subjects that end at or before its Start, walked from the
// nearest one leftwards through the end order, then those that start at or after its End, walked rightwards through the
// start order.  ties: 0 = the k nearest of each direction, 1 / 2 = one subject (the first / last met) per distinct
// distance, k distinct distances, 3 = every subject of the k nearest distinct distances.  Equal ends / starts are met
// in descending / ascending row order (both orders are stable sorts of the rows).  EMIT = false counts the entries of
// a query (previous + next), EMIT = true writes (query label, subject row, distance = gap + 1) at offsets[i]: previous
// entries first, nearest first in both directions.
template <bool EMIT, typename V, typename E>
__device__ __forceinline__ int64_t knear_walk(int64_t p, const int64_t end, const int step, int64_t k, int ties, V value,
                                              E emit) {
    int64_t n = 0;
    if (k <= 0) return 0;
    if (ties == 0) {
        for (; p != end && n < k; p += step, n++)
            if (EMIT) emit(p);
        return n;
    }
    int64_t distinct = 0, pend = -1;
    uint64_t last = 0;
    bool have = false;
    for (; p != end; p += step) {
        const uint64_t v = value(p);
        if (!have || v != last) {
            if (ties == 2 && pend >= 0) { if (EMIT) emit(pend); n++; }
            if (distinct == k) { pend = -1; break; }
            distinct++;
            last = v;
            have = true;
            if (ties == 1) { if (EMIT) emit(p); n++; }
        }
        if (ties == 3) { if (EMIT) emit(p); n++; }
        pend = p;
    }
    if (ties == 2 && pend >= 0) { if (EMIT) emit(pend); n++; }
    return n;
}

template <typename X, bool EMIT>
__global__ void __launch_bounds__(OV_THREADS, 4)
k_knear(iv_index_t ix, iv_queries_t q, const int64_t* __restrict__ kq, int ties, int64_t* __restrict__ counts,
        const int64_t* __restrict__ offsets, int64_t* __restrict__ out_q, int64_t* __restrict__ out_s,
        int64_t* __restrict__ out_d) {
    __shared__ int span[2];
    const int64_t tbase = (int64_t)blockIdx.x * OV_THREADS;
    const int64_t last = tbase + OV_THREADS - 1 < q.n - 1 ? tbase + OV_THREADS - 1 : q.n - 1;
    const FlatCache<X> fc(ix, q, block_seg_span(q.seg_off, q.n_seg, tbase, last, span));
    const int64_t i = tbase + threadIdx.x;
    if (i >= q.n) return;
    int seg;
    const Flat<X> f = fc.of(ix, q, i, &seg);
    const int64_t s = __ldg(q.start + i), e = __ldg(q.end + i);
    const int64_t k = __ldg(kq + i);
    int64_t n = 0, o = EMIT ? __ldg(offsets + i) : 0;
    const int64_t ql = (EMIT && q.label) ? __ldg(q.label + i) : i;
    if (f.g >= 0 && k > 0) {
        const int mode = q.seg_mode ? __ldg(q.seg_mode + seg) : IV_NEAREST_BOTH;
        const X sc = f.at(s), ec = f.at(e);
        const int64_t g_first = __ldg(ix.groups.row_off + f.g), g_end = __ldg(ix.groups.row_off + f.g + 1);
        if (mode != IV_NEAREST_NEXT && s >= f.lo) {
            const int64_t r = rank_e_lt<X>(ix, (X)(sc + 1));  // subject ends <= s
            n += knear_walk<EMIT>(r - 1, g_first - 1, -1, k, ties,
                                  [&](int64_t p) { return (uint64_t)__ldg(ix.key_e + p); },
                                  [&](int64_t p) {
                                      out_q[o] = ql;
                                      out_s[o] = __ldg(ix.row_e + p);
                                      out_d[o] = s - f.raw(__ldg(ix.key_e + p)) + 1;
                                      o++;
                                  });
        }
        if (mode != IV_NEAREST_PREVIOUS && !(e > f.lo && (uint64_t)(e - f.lo) > f.width)) {
            const int64_t j = rank_s_lt<X>(ix, ec);  // subject starts < e
            n += knear_walk<EMIT>(j > g_first ? j : g_first, g_end, 1, k, ties,
                                  [&](int64_t p) { return (uint64_t)__ldg(ix.key_s + p); },
                                  [&](int64_t p) {
                                      out_q[o] = ql;
                                      out_s[o] = __ldg(ix.row_s + p);
                                      out_d[o] = f.raw(__ldg(ix.key_s + p)) - e + 1;
                                      o++;
                                  });
        }
    }
    if (!EMIT) counts[i] = n;
}

// ---- subtract -------------------------------------------------------------------------------------------------------
// NCLS.set_difference_helper (pyranges/methods/subtraction.py:63-65): the parts of every query not covered by the
// subjects.  The subjects are the MERGED other intervals (pyranges/pyranges_main.py:4868), i.e. disjoint: the hits
// of a query are then the `count` consecutive entries of the start order that follow the #{E' <= s'} subjects lying
// entirely before it.  A query without hit survives as one piece, a completely covered one yields none.
template <typename X, bool EMIT>
__global__ void __launch_bounds__(OV_THREADS, 6)
k_subtract(iv_index_t ix, iv_queries_t q, int64_t* __restrict__ pieces, const int64_t* __restrict__ offsets,
           int64_t* __restrict__ out_q, int64_t* __restrict__ out_start, int64_t* __restrict__ out_end) {
    __shared__ int span[2];
    const int64_t tbase = (int64_t)blockIdx.x * OV_THREADS;
    const int64_t last = tbase + OV_THREADS - 1 < q.n - 1 ? tbase + OV_THREADS - 1 : q.n - 1;
    const FlatCache<X> fc(ix, q, block_seg_span(q.seg_off, q.n_seg, tbase, last, span));
    const int64_t i = tbase + threadIdx.x;
    if (i >= q.n) return;
    const Flat<X> f = fc.of(ix, q, i);
    const int64_t s = __ldg(q.start + i), e = __ldg(q.end + i);
    int64_t o = EMIT ? __ldg(offsets + i) : 0, np = 0;
    const int64_t ql = (EMIT && q.label) ? __ldg(q.label + i) : i;
    auto piece = [&](int64_t a, int64_t b) {
        if (EMIT) { out_q[o] = ql; out_start[o] = a; out_end[o] = b; o++; }
        np++;
    };
    int64_t cnt = 0;
    uint32_t p0 = 0;
    if (f.g >= 0) {
        const X sc = f.at(s), ec = f.at(e);
        cnt = count_all<X>(ix, sc, ec);
        if (cnt > 0) p0 = rank_e_lt<X>(ix, (X)(sc + 1));
    }
    if (cnt == 0) {
        piece(s, e);
    } else {
        int64_t cur = s;
        for (int64_t j = 0; j < cnt; j++) {
            const int64_t S = f.raw(__ldg(ix.key_s + p0 + j)), En = f.raw(__ldg(ix.end_s + p0 + j));
            if (S > cur) piece(cur, S);
            cur = En > cur ? En : cur;
        }
        if (cur < e) piece(cur, e);
    }
    if (!EMIT) pieces[i] = np;
}

// ---- pairs per query segment, straight from the count pass ----------------------------------------------------------
// pairs of the rows before row r = tile offset + pair sums of the tile's granules before r's granule + counts of the
// granule's hit-list entries whose row lies before r.  One warp per segment: out[k] = prefix(seg_off[k+1]) -
// prefix(seg_off[k]).  Lets a multi-GPU caller start exchanging the per-key result counts while the emit pass runs.
template <typename X>
__device__ __forceinline__ int64_t pairs_before(const iv_queries_t& q, const HitEntry<X>* __restrict__ list,
                                                const uint32_t* __restrict__ gran_hits,
                                                const int64_t* __restrict__ gran_sums,
                                                const int64_t* __restrict__ tile_off, const int64_t* __restrict__ total,
                                                int64_t r) {
    const int lane = lane_id();
    if (r >= q.n) return __ldg(total);
    const int64_t t = r / OV_TILE, g = r / OV_GRAN;
    const int gin = (int)(g - t * OV_GPT);
    const uint32_t trow = (uint32_t)(r - t * OV_TILE);
    int64_t v = lane < gin ? __ldg(gran_sums + t * OV_GPT + lane) : 0;
    const uint32_t nh = __ldg(gran_hits + g);
    for (uint32_t k = lane; k < nh; k += 32) {
        const Hit<X> h = get_entry<X>(list + g * OV_GRAN + k);
        if (h.row < trow) v += h.cnt;
    }
    return warp_sum(v) + __ldg(tile_off + t);
}
template <typename X>
__global__ void __launch_bounds__(256)
k_seg_pairs(iv_queries_t q, const HitEntry<X>* __restrict__ list, const uint32_t* __restrict__ gran_hits,
            const int64_t* __restrict__ gran_sums, const int64_t* __restrict__ tile_off,
            const int64_t* __restrict__ total, int64_t* __restrict__ out) {
    const int k = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (k >= q.n_seg) return;
    const int64_t a = __ldg(q.seg_off + k), b = __ldg(q.seg_off + k + 1);
    int64_t v = 0;
    if (b > a) v = pairs_before<X>(q, list, gran_hits, gran_sums, tile_off, total, b) -
                   pairs_before<X>(q, list, gran_hits, gran_sums, tile_off, total, a);
    if (lane_id() == 0) out[k] = v;
}

static bool aligned16(const void* p) { return (((uintptr_t)p) & 15) == 0; }
static bool narrow(const iv_index_t* ix) { return ix->total_span < (int64_t)0xfffffff0ll; }

static int check_query(const iv_index_t* ix, const iv_queries_t* q, const char* who) {
    IV_REQUIRE(ix && q, IV_ERR_INVALID, "%s: null argument", who);
    IV_REQUIRE(q->n >= 0 && (q->n == 0 || (q->start && q->end && q->seg_off && q->seg_group && q->n_seg > 0)),
               IV_ERR_INVALID, "%s: bad query descriptor", who);
    return IV_OK;
}

// workspace of count / emit:
//   [tile sums -> tile offsets, int64 (n/1024 + 2)] [pair sums per granule, int64 (n/64 + 2)]
//   [hits per granule, uint32 (n/64 + 2)] [scan scratch]
//   [hit list: one 16-byte (32-bit index) or 32-byte (64-bit index) slot per query row]
static size_t ws_tiles(int64_t n) { return (size_t)cdiv(n > 0 ? n : 1, OV_TILE) + 2; }
// granules of the queries rounded up to a whole block of the lean count kernel (4 x 1024 rows): its ragged last block
// writes the granule slots of every started half-tile
static_assert(LN_TILES == 4, "ws_grans() sizes the granule arrays for blocks of LN_TILES tiles");
static size_t ws_grans(int64_t n) { return (size_t)cdiv(n > 0 ? n : 1, 4 * OV_TILE) * (4 * OV_TILE / OV_GRAN) + 2; }
struct OverlapWs {
    int64_t* tile_sums;
    int64_t* gran_sums;
    uint32_t* gran_hits;
    int64_t* scan;
    void* list;
    size_t bytes;
};
static OverlapWs plan_overlap_ws(void* ws, int64_t n, size_t entry_bytes) {
    Bump b(ws, 0);
    OverlapWs w;
    w.tile_sums = b.take<int64_t>(ws_tiles(n));
    w.gran_sums = b.take<int64_t>(ws_grans(n));
    w.gran_hits = b.take<uint32_t>(ws_grans(n));
    w.scan = b.take<int64_t>(scan_i64_large_temp_elems((int64_t)ws_tiles(n)));
    w.list = b.take<char>((size_t)cdiv(n > 0 ? n : 1, OV_TILE) * OV_TILE * entry_bytes);
    w.bytes = b.off;
    return w;
}

template <typename X>
static void launch_count(const iv_index_t& ix, const iv_queries_t& q, int mode, bool vec, int64_t* out_count, void* ws,
                         unsigned ntiles, cudaStream_t st) {
    OverlapWs w = plan_overlap_ws(ws, q.n, sizeof(HitEntry<X>));
    int64_t* gran_sums = ws ? w.gran_sums : nullptr;
    uint32_t* gran_hits = ws ? w.gran_hits : nullptr;
    HitEntry<X>* list = ws ? (HitEntry<X>*)w.list : nullptr;
    int64_t first = 0;
    const int64_t nblk = cdiv(q.n, LN_TILES * OV_TILE);
    if (mode == IV_MODE_LAST) mode = IV_MODE_ANY;  // one entry per query with hits: the count pass is the same
    if (mode != IV_MODE_CONTAIN && vec && nblk >= 16) {
        // large aligned inputs go through the lean kernel (ragged last block included)
        unsigned long long* gs = reinterpret_cast<unsigned long long*>(gran_sums);
        const unsigned nbk = (unsigned)nblk;
        if (mode == IV_MODE_ALL && q.slack == 0)
            k_count_lean<X, IV_MODE_ALL, false><<<nbk, LN_THREADS, 0, st>>>(ix, q, out_count, gs, gran_hits, list);
        else if (mode == IV_MODE_ALL)
            k_count_lean<X, IV_MODE_ALL, true><<<nbk, LN_THREADS, 0, st>>>(ix, q, out_count, gs, gran_hits, list);
        else if (q.slack == 0)
            k_count_lean<X, IV_MODE_ANY, false><<<nbk, LN_THREADS, 0, st>>>(ix, q, out_count, gs, gran_hits, list);
        else
            k_count_lean<X, IV_MODE_ANY, true><<<nbk, LN_THREADS, 0, st>>>(ix, q, out_count, gs, gran_hits, list);
        __atomic_fetch_add(&iv::g_launches, 1ull, __ATOMIC_RELAXED);
        first = nblk * LN_TILES;
    }
    if (first < (int64_t)ntiles) {
        const unsigned nb = (unsigned)(ntiles - first);
        if (mode == IV_MODE_ALL)
            k_count<X, IV_MODE_ALL><<<nb, OV_THREADS, 0, st>>>(ix, q, vec, out_count, gran_sums, gran_hits, list, first);
        else if (mode == IV_MODE_ANY)
            k_count<X, IV_MODE_ANY><<<nb, OV_THREADS, 0, st>>>(ix, q, vec, out_count, gran_sums, gran_hits, list, first);
        else
            k_count<X, IV_MODE_CONTAIN><<<nb, OV_THREADS, 0, st>>>(ix, q, vec, out_count, gran_sums, gran_hits, list, first);
        __atomic_fetch_add(&iv::g_launches, 1ull, __ATOMIC_RELAXED);
    }
    if (ws) {
        k_tile_sums<<<(unsigned)cdiv((int64_t)ntiles * OV_GPT, 256), 256, 0, st>>>(gran_sums, q.n, ntiles, w.tile_sums);
        __atomic_fetch_add(&iv::g_launches, 1ull, __ATOMIC_RELAXED);
    }
}

template <typename X>
static void launch_emit(const iv_index_t& ix, const iv_queries_t& q, int mode, const void* ws, int64_t* out_q,
                        int64_t* out_s, const int64_t* s_label, int64_t cap, cudaStream_t st) {
    const OverlapWs w = plan_overlap_ws((void*)ws, q.n, sizeof(HitEntry<X>));
    const HitEntry<X>* list = (const HitEntry<X>*)w.list;
    const unsigned nb = (unsigned)cdiv(q.n, OV_TILE);
    if (mode == IV_MODE_ALL)
        k_emit<X, IV_MODE_ALL><<<nb, EM_THREADS, 0, st>>>(ix, q, list, w.gran_hits, w.gran_sums, w.tile_sums, out_q, out_s, s_label, cap);
    else if (mode == IV_MODE_ANY)
        k_emit<X, IV_MODE_ANY><<<nb, EM_THREADS, 0, st>>>(ix, q, list, w.gran_hits, w.gran_sums, w.tile_sums, out_q, out_s, s_label, cap);
    else if (mode == IV_MODE_LAST)
        k_emit<X, IV_MODE_LAST><<<nb, EM_THREADS, 0, st>>>(ix, q, list, w.gran_hits, w.gran_sums, w.tile_sums, out_q, out_s, s_label, cap);
    else
        k_emit<X, IV_MODE_CONTAIN><<<nb, EM_THREADS, 0, st>>>(ix, q, list, w.gran_hits, w.gran_sums, w.tile_sums, out_q, out_s, s_label, cap);
}

}  // namespace iv

using namespace iv;

extern "C" size_t iv_overlap_workspace_bytes(int64_t n_queries) {
    // sized for the 64-bit index (32-byte list slots); a 32-bit index uses half of the list area
    return plan_overlap_ws(nullptr, n_queries, sizeof(HitEntry<uint64_t>)).bytes + 256;
}

extern "C" int iv_count_overlaps(const iv_index_t* index, const iv_queries_t* q, int32_t mode, int64_t* out_count,
                                 void* ws, size_t ws_bytes, int64_t* out_total, void* stream) {
    cudaStream_t st = (cudaStream_t)stream;
    int rc = check_query(index, q, "iv_count_overlaps");
    if (rc != IV_OK) return rc;
    IV_REQUIRE(mode >= IV_MODE_ALL && mode <= IV_MODE_LAST, IV_ERR_INVALID, "iv_count_overlaps: mode");
    const bool nar = narrow(index);
    const size_t need = plan_overlap_ws(nullptr, q->n, nar ? sizeof(HitEntry<uint32_t>) : sizeof(HitEntry<uint64_t>)).bytes;
    IV_REQUIRE(!ws || ws_bytes >= need, IV_ERR_WORKSPACE, "iv_count_overlaps: workspace %zu < %zu", ws_bytes, need);
    if (q->n == 0) {
        if (out_total) IV_CUDA(cudaMemsetAsync(out_total, 0, 8, st));
        return IV_OK;
    }
    IV_REQUIRE(out_count || ws, IV_ERR_INVALID, "iv_count_overlaps: neither out_count nor a workspace");
    IV_REQUIRE(!ws || (((uintptr_t)ws) & 255) == 0, IV_ERR_INVALID, "iv_count_overlaps: workspace not 256-byte aligned");
    int64_t ntiles = cdiv(q->n, OV_TILE);
    bool vec = aligned16(q->start) && aligned16(q->end) && (!out_count || aligned16(out_count));
    if (nar) launch_count<uint32_t>(*index, *q, mode, vec, out_count, ws, (unsigned)ntiles, st);
    else launch_count<uint64_t>(*index, *q, mode, vec, out_count, ws, (unsigned)ntiles, st);
    IV_CUDA(cudaGetLastError());  // launch_count tallies its own launches
    if (ws) {
        const OverlapWs w = plan_overlap_ws(ws, q->n, nar ? sizeof(HitEntry<uint32_t>) : sizeof(HitEntry<uint64_t>));
        rc = scan_exclusive_i64_large(w.tile_sums, ntiles, w.scan, out_total, st);
        if (rc != IV_OK) return rc;
    }
    return IV_OK;
}

extern "C" int iv_emit_pairs(const iv_index_t* index, const iv_queries_t* q, int32_t mode, const int64_t* counts,
                             const void* ws, int64_t* out_q, int64_t* out_s, const int64_t* s_label,
                             int64_t out_capacity, void* stream) {
    cudaStream_t st = (cudaStream_t)stream;
    (void)counts;  // the hit lists in the workspace carry the counts
    int rc = check_query(index, q, "iv_emit_pairs");
    if (rc != IV_OK) return rc;
    IV_REQUIRE(mode >= IV_MODE_ALL && mode <= IV_MODE_LAST, IV_ERR_INVALID, "iv_emit_pairs: mode");
    if (q->n == 0) return IV_OK;
    IV_REQUIRE(ws, IV_ERR_INVALID, "iv_emit_pairs: ws is null");
    IV_REQUIRE(out_capacity >= 0, IV_ERR_INVALID, "iv_emit_pairs: negative out_capacity");
    if (narrow(index)) launch_emit<uint32_t>(*index, *q, mode, ws, out_q, out_s, s_label, out_capacity, st);
    else launch_emit<uint64_t>(*index, *q, mode, ws, out_q, out_s, s_label, out_capacity, st);
    IV_LAUNCH_CHECK();
    return IV_OK;
}

extern "C" int iv_segment_pair_counts(const iv_index_t* index, const iv_queries_t* q, const void* ws,
                                      const int64_t* total, int64_t* out_seg_pairs, void* stream) {
    cudaStream_t st = (cudaStream_t)stream;
    int rc = check_query(index, q, "iv_segment_pair_counts");
    if (rc != IV_OK) return rc;
    if (q->n_seg <= 0) return IV_OK;
    IV_REQUIRE(out_seg_pairs, IV_ERR_INVALID, "iv_segment_pair_counts: null output");
    if (q->n == 0) {
        IV_CUDA(cudaMemsetAsync(out_seg_pairs, 0, (size_t)q->n_seg * 8, st));
        return IV_OK;
    }
    IV_REQUIRE(ws && total, IV_ERR_INVALID, "iv_segment_pair_counts: ws / total is null");
    const bool nar = narrow(index);
    const OverlapWs w = plan_overlap_ws((void*)ws, q->n, nar ? sizeof(HitEntry<uint32_t>) : sizeof(HitEntry<uint64_t>));
    const unsigned nb = (unsigned)cdiv(q->n_seg, 8);
    if (nar)
        k_seg_pairs<uint32_t><<<nb, 256, 0, st>>>(*q, (const HitEntry<uint32_t>*)w.list, w.gran_hits, w.gran_sums, w.tile_sums,
                                                  total, out_seg_pairs);
    else
        k_seg_pairs<uint64_t><<<nb, 256, 0, st>>>(*q, (const HitEntry<uint64_t>*)w.list, w.gran_hits, w.gran_sums, w.tile_sums,
                                                  total, out_seg_pairs);
    IV_LAUNCH_CHECK();
    return IV_OK;
}

extern "C" int iv_coverage_bp(const iv_index_t* index, const iv_queries_t* q, int64_t* out_bp, double* out_fraction,
                              void* stream) {
    cudaStream_t st = (cudaStream_t)stream;
    int rc = check_query(index, q, "iv_coverage_bp");
    if (rc != IV_OK) return rc;
    if (q->n == 0) return IV_OK;
    unsigned nb = (unsigned)cdiv(q->n, OV_THREADS);
    if (narrow(index)) k_coverage<uint32_t><<<nb, OV_THREADS, 0, st>>>(*index, *q, out_bp, out_fraction);
    else k_coverage<uint64_t><<<nb, OV_THREADS, 0, st>>>(*index, *q, out_bp, out_fraction);
    IV_LAUNCH_CHECK();
    return IV_OK;
}

extern "C" size_t iv_subtract_workspace_bytes(int64_t n_queries) {
    return scan_i64_large_temp_elems(n_queries > 0 ? n_queries : 1) * sizeof(int64_t) + 256;
}

extern "C" int iv_subtract_count(const iv_index_t* index, const iv_queries_t* q, int64_t* out_offsets, void* ws,
                                 size_t ws_bytes, int64_t* out_total, void* stream) {
    cudaStream_t st = (cudaStream_t)stream;
    int rc = check_query(index, q, "iv_subtract_count");
    if (rc != IV_OK) return rc;
    IV_REQUIRE(out_total && ws && ws_bytes >= iv_subtract_workspace_bytes(q->n), IV_ERR_WORKSPACE, "iv_subtract_count: workspace");
    if (q->n == 0) {
        IV_CUDA(cudaMemsetAsync(out_total, 0, 8, st));
        return IV_OK;
    }
    IV_REQUIRE(out_offsets, IV_ERR_INVALID, "iv_subtract_count: out_offsets is null");
    unsigned nb = (unsigned)cdiv(q->n, OV_THREADS);
    if (narrow(index)) k_subtract<uint32_t, false><<<nb, OV_THREADS, 0, st>>>(*index, *q, out_offsets, nullptr, nullptr, nullptr, nullptr);
    else k_subtract<uint64_t, false><<<nb, OV_THREADS, 0, st>>>(*index, *q, out_offsets, nullptr, nullptr, nullptr, nullptr);
    IV_LAUNCH_CHECK();
    return scan_exclusive_i64_large(out_offsets, q->n, (int64_t*)ws, out_total, st);
}

extern "C" int iv_subtract_emit(const iv_index_t* index, const iv_queries_t* q, const int64_t* offsets, int64_t* out_q,
                                int64_t* out_start, int64_t* out_end, void* stream) {
    cudaStream_t st = (cudaStream_t)stream;
    int rc = check_query(index, q, "iv_subtract_emit");
    if (rc != IV_OK) return rc;
    if (q->n == 0) return IV_OK;
    IV_REQUIRE(offsets && out_q && out_start && out_end, IV_ERR_INVALID, "iv_subtract_emit: null argument");
    unsigned nb = (unsigned)cdiv(q->n, OV_THREADS);
    if (narrow(index)) k_subtract<uint32_t, true><<<nb, OV_THREADS, 0, st>>>(*index, *q, nullptr, offsets, out_q, out_start, out_end);
    else k_subtract<uint64_t, true><<<nb, OV_THREADS, 0, st>>>(*index, *q, nullptr, offsets, out_q, out_start, out_end);
    IV_LAUNCH_CHECK();
    return IV_OK;
}

extern "C" int iv_nearest(const iv_index_t* index, const iv_queries_t* q, int32_t overlap, int64_t* out_row,
                          int64_t* out_dist, void* stream) {
    cudaStream_t st = (cudaStream_t)stream;
    int rc = check_query(index, q, "iv_nearest");
    if (rc != IV_OK) return rc;
    if (q->n == 0) return IV_OK;
    IV_REQUIRE(out_row && out_dist, IV_ERR_INVALID, "iv_nearest: null output");
    IV_REQUIRE(index->n == 0 || index->row_e, IV_ERR_INVALID, "iv_nearest: index built without IV_INDEX_ENDS_PERM");
    unsigned nb = (unsigned)cdiv(q->n, OV_THREADS);
    if (narrow(index)) k_nearest<uint32_t><<<nb, OV_THREADS, 0, st>>>(*index, *q, overlap, out_row, out_dist);
    else k_nearest<uint64_t><<<nb, OV_THREADS, 0, st>>>(*index, *q, overlap, out_row, out_dist);
    IV_LAUNCH_CHECK();
    return IV_OK;
}

extern "C" size_t iv_k_nearest_workspace_bytes(int64_t n_queries) { return iv_subtract_workspace_bytes(n_queries); }

extern "C" int iv_k_nearest_count(const iv_index_t* index, const iv_queries_t* q, const int64_t* k, int32_t ties,
                                  int64_t* out_offsets, void* ws, size_t ws_bytes, int64_t* out_total, void* stream) {
    cudaStream_t st = (cudaStream_t)stream;
    int rc = check_query(index, q, "iv_k_nearest_count");
    if (rc != IV_OK) return rc;
    IV_REQUIRE(ties >= 0 && ties <= 3, IV_ERR_INVALID, "iv_k_nearest_count: ties must be 0 .. 3");
    IV_REQUIRE(out_total && ws && ws_bytes >= iv_k_nearest_workspace_bytes(q->n), IV_ERR_WORKSPACE, "iv_k_nearest_count: workspace");
    if (q->n == 0) {
        IV_CUDA(cudaMemsetAsync(out_total, 0, 8, st));
        return IV_OK;
    }
    IV_REQUIRE(out_offsets && k, IV_ERR_INVALID, "iv_k_nearest_count: null argument");
    IV_REQUIRE(index->n == 0 || index->row_e, IV_ERR_INVALID, "iv_k_nearest_count: index built without IV_INDEX_ENDS_PERM");
    unsigned nb = (unsigned)cdiv(q->n, OV_THREADS);
    if (narrow(index)) k_knear<uint32_t, false><<<nb, OV_THREADS, 0, st>>>(*index, *q, k, ties, out_offsets, nullptr, nullptr, nullptr, nullptr);
    else k_knear<uint64_t, false><<<nb, OV_THREADS, 0, st>>>(*index, *q, k, ties, out_offsets, nullptr, nullptr, nullptr, nullptr);
    IV_LAUNCH_CHECK();
    return scan_exclusive_i64_large(out_offsets, q->n, (int64_t*)ws, out_total, st);
}

extern "C" int iv_k_nearest_emit(const iv_index_t* index, const iv_queries_t* q, const int64_t* k, int32_t ties,
                                 const int64_t* offsets, int64_t* out_q, int64_t* out_s, int64_t* out_dist, void* stream) {
    cudaStream_t st = (cudaStream_t)stream;
    int rc = check_query(index, q, "iv_k_nearest_emit");
    if (rc != IV_OK) return rc;
    if (q->n == 0) return IV_OK;
    IV_REQUIRE(ties >= 0 && ties <= 3, IV_ERR_INVALID, "iv_k_nearest_emit: ties must be 0 .. 3");
    IV_REQUIRE(k && offsets && out_q && out_s && out_dist, IV_ERR_INVALID, "iv_k_nearest_emit: null argument");
    IV_REQUIRE(index->n == 0 || index->row_e, IV_ERR_INVALID, "iv_k_nearest_emit: index built without IV_INDEX_ENDS_PERM");
    unsigned nb = (unsigned)cdiv(q->n, OV_THREADS);
    if (narrow(index)) k_knear<uint32_t, true><<<nb, OV_THREADS, 0, st>>>(*index, *q, k, ties, nullptr, offsets, out_q, out_s, out_dist);
    else k_knear<uint64_t, true><<<nb, OV_THREADS, 0, st>>>(*index, *q, k, ties, nullptr, offsets, out_q, out_s, out_dist);
    IV_LAUNCH_CHECK();
    return IV_OK;
}
