// overlap.cu -- batched overlap queries against the flattened subject index.
//
// Replaces (reference call sites; the native code itself lives in the un-vendored ncls /
// sorted_nearest packages):
//   NCLS.all_overlaps_both / all_overlaps_self / has_overlaps / all_containments_both /
//   first_overlap_both          pyranges/methods/join.py:15-23, intersection.py:73-78, coverage.py:26
//   NCLS.coverage               pyranges/methods/coverage.py:64
//   nearest_previous/next/_nonoverlapping   pyranges/methods/nearest.py:59,69,113
//
// count   = #{S' < e'} - #{E' <= s'}         two rank look-ups: one 32-byte bin record each (usually the
//                                            same record) + the bin's own few 32-bit offsets; no scan
// hits    = class A: s' <= S' < e'  (contiguous in start order, all of them overlap)
//         + class B: S' < s' < E'   (stabbing set) = members of the bin's stabbing list that end after s'
//                                    + earlier starts of the same bin that end after s'
// emit    = count -> per-tile sums -> exclusive scan of tile sums -> emit kernel re-scans the counts of
//           its tile in shared memory (no global offsets array).
#include "common.cuh"

namespace iv {

constexpr int OV_THREADS = 256;
constexpr int OV_ITEMS = 4;
constexpr int OV_TILE = OV_THREADS * OV_ITEMS;  // 1024 queries per block

// first 16 bytes of a bin record: bs, be, bs1, be1
__device__ __forceinline__ uint4 bin_ranks(const iv_index_t& ix, uint64_t b) {
    return __ldg(reinterpret_cast<const uint4*>(ix.bins + b));
}
__device__ __forceinline__ uint2 bin_cross(const iv_index_t& ix, uint64_t b) {
    return __ldg(reinterpret_cast<const uint2*>(ix.bins + b) + 2);
}
// lo + #{off[p] < rel : p in [lo, hi)}   (off ascending inside a bin)
__device__ __forceinline__ uint32_t rank_in(const uint32_t* __restrict__ off, uint32_t lo, uint32_t hi, uint32_t rel) {
    if (hi - lo <= 8) {
        uint32_t r = lo;
        for (uint32_t p = lo; p < hi; p++) r += __ldg(off + p) < rel ? 1u : 0u;
        return r;
    }
    while (lo < hi) {
        uint32_t mid = (lo + hi) >> 1;
        if (__ldg(off + mid) < rel) lo = mid + 1; else hi = mid;
    }
    return lo;
}
// #starts < x, #ends < x  (x: flattened coordinate)
__device__ __forceinline__ uint32_t rank_s_lt(const iv_index_t& ix, uint64_t x) {
    uint64_t b = x >> ix.shift;
    uint4 r = bin_ranks(ix, b);
    return rank_in(ix.soff, r.x, r.z, (uint32_t)(x - (b << ix.shift)));
}
__device__ __forceinline__ uint32_t rank_e_lt(const iv_index_t& ix, uint64_t x) {
    uint64_t b = x >> ix.shift;
    uint4 r = bin_ranks(ix, b);
    return rank_in(ix.eoff, r.y, r.w, (uint32_t)(x - (b << ix.shift)));
}

struct QGeom {
    int g;           // partner group, -1 = none
    int64_t s, e;    // raw (slack-extended) coordinates
    int64_t lo, hi, base;
    uint64_t sc, ec; // flattened, clamped
    bool clamped;    // query sticks out of [lo, hi]
};

__device__ __forceinline__ QGeom make_geom(const iv_index_t& ix, const iv_queries_t& q, int seg, int64_t s, int64_t e) {
    QGeom r;
    r.g = __ldg(q.seg_group + seg);
    if (q.slack != 0) {
        s -= q.slack;
        if (s < 0) s = 0;
        e += q.slack;
    }
    r.s = s; r.e = e;
    if (r.g < 0) { r.lo = r.hi = r.base = 0; r.sc = r.ec = 0; r.clamped = true; return r; }
    r.lo = __ldg(ix.groups.lo + r.g);
    r.hi = __ldg(ix.groups.hi + r.g);
    r.base = __ldg(ix.groups.base + r.g);
    int64_t cs = s < r.lo ? r.lo : (s > r.hi ? r.hi : s);
    int64_t ce = e < r.lo ? r.lo : (e > r.hi ? r.hi : e);
    r.clamped = (cs != s) || (ce != e);
    r.sc = (uint64_t)(cs - r.lo + r.base);
    r.ec = (uint64_t)(ce - r.lo + r.base);
    return r;
}

__device__ __forceinline__ int64_t count_all(const iv_index_t& ix, const QGeom& g) {
    if (g.g < 0 || ix.n == 0) return 0;
    const uint64_t xe = g.ec, xs = g.sc + 1;
    const uint64_t b1 = xe >> ix.shift, b0 = xs >> ix.shift;
    const uint4 r1 = bin_ranks(ix, b1);
    const uint4 r0 = b0 == b1 ? r1 : bin_ranks(ix, b0);
    int64_t c = (int64_t)rank_in(ix.soff, r1.x, r1.z, (uint32_t)(xe - (b1 << ix.shift))) -
                (int64_t)rank_in(ix.eoff, r0.y, r0.w, (uint32_t)(xs - (b0 << ix.shift)));
    return c > 0 ? c : 0;
}

// Visit every subject overlapping [sc, ec).  WANT_P: f(position in start order), else f(subject row).
template <bool WANT_P, typename F>
__device__ __forceinline__ void for_each_hit(const iv_index_t& ix, uint64_t sc, uint64_t ec, F&& f) {
    const uint64_t b = sc >> ix.shift;
    const uint32_t rel = (uint32_t)(sc - (b << ix.shift));
    const uint4 r = bin_ranks(ix, b);
    const uint32_t pS = rank_in(ix.soff, r.x, r.z, rel);
    const uint64_t b1 = ec >> ix.shift;
    const uint32_t pE = b1 == b ? rank_in(ix.soff, r.x, r.z, (uint32_t)(ec - (b << ix.shift))) : rank_s_lt(ix, ec);
    for (uint32_t p = pS; p < pE; p++) f(WANT_P ? p : __ldg(ix.row_s + p));  // class A
    uint32_t pL = r.x;  // class B, started in this bin: only starts later than s' - max_len can still be open
    if (pS - pL > 8 && ix.groups.max_len > 0 && (int64_t)rel >= ix.groups.max_len)
        pL = rank_in(ix.soff, pL, pS, (uint32_t)((int64_t)rel - ix.groups.max_len + 1));
    for (uint32_t p = pL; p < pS; p++)
        if (__ldg(ix.seoff + p) > rel) f(WANT_P ? p : __ldg(ix.row_s + p));
    const uint2 c = bin_cross(ix, b);
    for (uint32_t j = c.x; j < c.y; j++) {                                    // class B, open at the bin origin
        const uint2 en = __ldg(reinterpret_cast<const uint2*>(ix.cross) + j);
        if (en.y > rel) f(WANT_P ? __ldg(ix.cross_p + j) : en.x);
    }
}

// first hit in nested-containment-list order: smallest start, then largest end, then smallest row
struct FirstHit {
    const iv_index_t& ix;
    uint64_t bs, be;
    uint32_t brow;
    bool any;
    __device__ FirstHit(const iv_index_t& i) : ix(i), bs(0), be(0), brow(0), any(false) {}
    __device__ __forceinline__ void operator()(uint32_t j) {
        uint64_t s = __ldg(ix.key_s + j), e = __ldg(ix.end_s + j);
        uint32_t r = __ldg(ix.row_s + j);
        bool better = !any || s < bs || (s == bs && (e > be || (e == be && r < brow)));
        if (better) { bs = s; be = e; brow = r; any = true; }
    }
};

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
template <int MODE>
__global__ void __launch_bounds__(OV_THREADS)
k_count(iv_index_t ix, iv_queries_t q, bool vec, int64_t* __restrict__ out_count, int64_t* __restrict__ tile_sums) {
    __shared__ int span[2];
    __shared__ int64_t red[OV_THREADS / 32];
    const int64_t tbase = (int64_t)blockIdx.x * OV_TILE;
    const int64_t last = tbase + OV_TILE - 1 < q.n - 1 ? tbase + OV_TILE - 1 : q.n - 1;
    SegSpan sp = block_seg_span(q.seg_off, q.n_seg, tbase, last, span);
    const int64_t i0 = tbase + (int64_t)threadIdx.x * OV_ITEMS;
    int64_t s[OV_ITEMS], e[OV_ITEMS], c[OV_ITEMS];
    load_items(q.start, i0, q.n, vec, s);
    load_items(q.end, i0, q.n, vec, e);
    int64_t tsum = 0;
#pragma unroll
    for (int j = 0; j < OV_ITEMS; j++) {
        c[j] = 0;
        if (i0 + j < q.n) {
            QGeom g = make_geom(ix, q, sp.of(q.seg_off, i0 + j), s[j], e[j]);
            int64_t ca = count_all(ix, g);
            if (MODE == IV_MODE_ALL) c[j] = ca;
            else if (MODE == IV_MODE_ANY) c[j] = ca > 0 ? 1 : 0;
            else {  // containment: subject.start <= q.start && q.end <= subject.end (raw coordinates)
                if (ca > 0 && !g.clamped) {
                    int64_t k = 0;
                    uint64_t sc = g.sc, ec = g.ec;
                    for_each_hit<true>(ix, sc, ec, [&](uint32_t p) {
                        k += (__ldg(ix.key_s + p) <= sc && ec <= __ldg(ix.end_s + p)) ? 1 : 0;
                    });
                    c[j] = k;
                }
            }
            tsum += c[j];
        }
    }
    store_items(out_count, i0, q.n, vec, c);
    if (tile_sums) {
        tsum = warp_sum(tsum);
        if (lane_id() == 0) red[threadIdx.x >> 5] = tsum;
        __syncthreads();
        if (threadIdx.x < 32) {
            int64_t x = threadIdx.x < OV_THREADS / 32 ? red[threadIdx.x] : 0;
            x = warp_sum(x);
            if (threadIdx.x == 0) tile_sums[blockIdx.x] = x;
        }
    }
}

// ---- emit -----------------------------------------------------------------------------------------------------------
template <int MODE>
__global__ void __launch_bounds__(OV_THREADS)
k_emit(iv_index_t ix, iv_queries_t q, bool vec, const int64_t* __restrict__ counts,
       const int64_t* __restrict__ tile_off, int64_t* __restrict__ out_q, int64_t* __restrict__ out_s,
       const int64_t* __restrict__ s_label) {
    __shared__ int span[2];
    __shared__ int64_t sm[OV_THREADS / 32 + 1];
    const int64_t tbase = (int64_t)blockIdx.x * OV_TILE;
    const int64_t last = tbase + OV_TILE - 1 < q.n - 1 ? tbase + OV_TILE - 1 : q.n - 1;
    SegSpan sp = block_seg_span(q.seg_off, q.n_seg, tbase, last, span);
    const int64_t i0 = tbase + (int64_t)threadIdx.x * OV_ITEMS;
    int64_t c[OV_ITEMS];
    load_items(counts, i0, q.n, vec, c);
    int64_t tsum = 0;
#pragma unroll
    for (int j = 0; j < OV_ITEMS; j++) tsum += c[j];
    int64_t tot;
    int64_t o = block_excl_sum<int64_t, OV_THREADS>(tsum, sm, tot) + __ldg(tile_off + blockIdx.x);
    if (tot == 0) return;
    if (tsum == 0) return;
    int64_t s[OV_ITEMS], e[OV_ITEMS];
    load_items(q.start, i0, q.n, vec, s);
    load_items(q.end, i0, q.n, vec, e);
#pragma unroll
    for (int j = 0; j < OV_ITEMS; j++) {
        if (c[j] == 0) continue;
        const int64_t i = i0 + j;
        const int64_t ql = q.label ? __ldg(q.label + i) : i;
        QGeom g = make_geom(ix, q, sp.of(q.seg_off, i), s[j], e[j]);
        if (MODE == IV_MODE_ALL) {
            for_each_hit<false>(ix, g.sc, g.ec, [&](uint32_t r) {
                if (out_q) out_q[o] = ql;
                if (out_s) out_s[o] = s_label ? __ldg(s_label + r) : (int64_t)r;
                o++;
            });
        } else if (MODE == IV_MODE_ANY) {
            if (out_q) out_q[o] = ql;
            if (out_s) {
                FirstHit fh(ix);
                for_each_hit<true>(ix, g.sc, g.ec, fh);
                int64_t r = fh.brow;
                out_s[o] = s_label ? __ldg(s_label + r) : r;
            }
            o++;
        } else {
            uint64_t sc = g.sc, ec = g.ec;
            for_each_hit<true>(ix, sc, ec, [&](uint32_t p) {
                if (__ldg(ix.key_s + p) <= sc && ec <= __ldg(ix.end_s + p)) {
                    if (out_q) out_q[o] = ql;
                    if (out_s) { int64_t r = __ldg(ix.row_s + p); out_s[o] = s_label ? __ldg(s_label + r) : r; }
                    o++;
                }
            });
        }
    }
}

// ---- NCLS.coverage -----------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(OV_THREADS)
k_coverage(iv_index_t ix, iv_queries_t q, int64_t* __restrict__ out_bp, double* __restrict__ out_frac) {
    __shared__ int span[2];
    const int64_t tbase = (int64_t)blockIdx.x * OV_THREADS;
    const int64_t last = tbase + OV_THREADS - 1 < q.n - 1 ? tbase + OV_THREADS - 1 : q.n - 1;
    SegSpan sp = block_seg_span(q.seg_off, q.n_seg, tbase, last, span);
    const int64_t i = tbase + threadIdx.x;
    if (i >= q.n) return;
    int64_t s = __ldg(q.start + i), e = __ldg(q.end + i);
    QGeom g = make_geom(ix, q, sp.of(q.seg_off, i), s, e);
    int64_t ca = count_all(ix, g);
    int64_t bp = 0;
    if (ca > 0) {
        uint64_t sc = g.sc, ec = g.ec;
        for_each_hit<true>(ix, sc, ec, [&](uint32_t p) {
            uint64_t S = __ldg(ix.key_s + p), E = __ldg(ix.end_s + p);
            bp += (int64_t)((E < ec ? E : ec) - (S > sc ? S : sc));
        });
    }
    if (out_bp) out_bp[i] = bp;
    if (out_frac) {
        int64_t len = e - s;
        out_frac[i] = len == 0 ? 0.0 : (double)bp / (double)len;
    }
}

// ---- nearest --------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(OV_THREADS)
k_nearest(iv_index_t ix, iv_queries_t q, int overlap, int64_t* __restrict__ out_row, int64_t* __restrict__ out_dist) {
    __shared__ int span[2];
    const int64_t tbase = (int64_t)blockIdx.x * OV_THREADS;
    const int64_t last = tbase + OV_THREADS - 1 < q.n - 1 ? tbase + OV_THREADS - 1 : q.n - 1;
    SegSpan sp = block_seg_span(q.seg_off, q.n_seg, tbase, last, span);
    const int64_t i = tbase + threadIdx.x;
    if (i >= q.n) return;
    const int seg = sp.of(q.seg_off, i);
    int64_t s = __ldg(q.start + i), e = __ldg(q.end + i);
    QGeom g = make_geom(ix, q, seg, s, e);
    int64_t row = -1, dist = -1;
    if (g.g >= 0 && ix.n > 0) {
        const int mode = q.seg_mode ? __ldg(q.seg_mode + seg) : IV_NEAREST_BOTH;
        bool done = false;
        if (overlap) {
            int64_t ca = count_all(ix, g);
            if (ca > 0) {
                FirstHit fh(ix);
                for_each_hit<true>(ix, g.sc, g.ec, fh);
                row = fh.brow; dist = 0; done = true;
            }
        }
        if (!done) {
            const int64_t g_first = __ldg(ix.groups.row_off + g.g), g_end = __ldg(ix.groups.row_off + g.g + 1);
            int64_t prow = -1, pdist = -1, nrow = -1, ndist = -1;
            if (mode != IV_NEAREST_NEXT && g.s >= g.lo) {
                // last subject end <= s ; ties: first of the equal run (earliest row, stable sort)
                uint64_t sc = (uint64_t)((g.s > g.hi ? g.hi : g.s) - g.lo + g.base);
                int64_t r = rank_e_lt(ix, sc + 1);
                if (r > g_first) {
                    uint64_t ev = __ldg(ix.key_e + (r - 1));
                    int64_t jf = rank_e_lt(ix, ev);
                    prow = __ldg(ix.row_e + jf);
                    pdist = g.s - ((int64_t)ev - g.base + g.lo) + 1;
                }
            }
            if (mode != IV_NEAREST_PREVIOUS && g.e <= g.hi) {
                uint64_t ec = (uint64_t)((g.e < g.lo ? g.lo : g.e) - g.lo + g.base);
                int64_t j = rank_s_lt(ix, ec);
                if (j < g_end) {
                    nrow = __ldg(ix.row_s + j);
                    ndist = ((int64_t)__ldg(ix.key_s + j) - g.base + g.lo) - g.e + 1;
                }
            }
            if (prow >= 0 && (nrow < 0 || pdist <= ndist)) { row = prow; dist = pdist; }
            else if (nrow >= 0) { row = nrow; dist = ndist; }
        }
    }
    out_row[i] = row;
    out_dist[i] = dist;
}

static bool aligned16(const void* p) { return (((uintptr_t)p) & 15) == 0; }

static int check_query(const iv_index_t* ix, const iv_queries_t* q, const char* who) {
    IV_REQUIRE(ix && q, IV_ERR_INVALID, "%s: null argument", who);
    IV_REQUIRE(q->n >= 0 && (q->n == 0 || (q->start && q->end && q->seg_off && q->seg_group && q->n_seg > 0)),
               IV_ERR_INVALID, "%s: bad query descriptor", who);
    return IV_OK;
}

}  // namespace iv

using namespace iv;

extern "C" size_t iv_overlap_workspace_bytes(int64_t n_queries) {
    return (size_t)(cdiv(n_queries > 0 ? n_queries : 1, OV_TILE) + 2) * sizeof(int64_t) + 256;
}

extern "C" int iv_count_overlaps(const iv_index_t* index, const iv_queries_t* q, int32_t mode, int64_t* out_count,
                                 void* ws, size_t ws_bytes, int64_t* out_total, void* stream) {
    cudaStream_t st = (cudaStream_t)stream;
    int rc = check_query(index, q, "iv_count_overlaps");
    if (rc != IV_OK) return rc;
    IV_REQUIRE(mode >= IV_MODE_ALL && mode <= IV_MODE_CONTAIN, IV_ERR_INVALID, "iv_count_overlaps: mode");
    IV_REQUIRE(!ws || ws_bytes >= iv_overlap_workspace_bytes(q->n), IV_ERR_WORKSPACE, "iv_count_overlaps: workspace");
    if (q->n == 0) {
        if (out_total) IV_CUDA(cudaMemsetAsync(out_total, 0, 8, st));
        return IV_OK;
    }
    IV_REQUIRE(out_count, IV_ERR_INVALID, "iv_count_overlaps: out_count is null");
    int64_t ntiles = cdiv(q->n, OV_TILE);
    int64_t* tile_sums = (int64_t*)ws;
    bool vec = aligned16(q->start) && aligned16(q->end) && aligned16(out_count);
    if (mode == IV_MODE_ALL)
        k_count<IV_MODE_ALL><<<(unsigned)ntiles, OV_THREADS, 0, st>>>(*index, *q, vec, out_count, tile_sums);
    else if (mode == IV_MODE_ANY)
        k_count<IV_MODE_ANY><<<(unsigned)ntiles, OV_THREADS, 0, st>>>(*index, *q, vec, out_count, tile_sums);
    else
        k_count<IV_MODE_CONTAIN><<<(unsigned)ntiles, OV_THREADS, 0, st>>>(*index, *q, vec, out_count, tile_sums);
    IV_LAUNCH_CHECK();
    if (tile_sums) {
        rc = scan_exclusive_i64_inplace(tile_sums, ntiles, out_total, st);
        if (rc != IV_OK) return rc;
    }
    return IV_OK;
}

extern "C" int iv_emit_pairs(const iv_index_t* index, const iv_queries_t* q, int32_t mode, const int64_t* counts,
                             const void* ws, int64_t* out_q, int64_t* out_s, const int64_t* s_label, void* stream) {
    cudaStream_t st = (cudaStream_t)stream;
    int rc = check_query(index, q, "iv_emit_pairs");
    if (rc != IV_OK) return rc;
    IV_REQUIRE(mode >= IV_MODE_ALL && mode <= IV_MODE_CONTAIN, IV_ERR_INVALID, "iv_emit_pairs: mode");
    if (q->n == 0) return IV_OK;
    IV_REQUIRE(counts && ws, IV_ERR_INVALID, "iv_emit_pairs: counts / ws is null");
    int64_t ntiles = cdiv(q->n, OV_TILE);
    const int64_t* tile_off = (const int64_t*)ws;
    bool vec = aligned16(q->start) && aligned16(q->end) && aligned16(counts);
    if (mode == IV_MODE_ALL)
        k_emit<IV_MODE_ALL><<<(unsigned)ntiles, OV_THREADS, 0, st>>>(*index, *q, vec, counts, tile_off, out_q, out_s, s_label);
    else if (mode == IV_MODE_ANY)
        k_emit<IV_MODE_ANY><<<(unsigned)ntiles, OV_THREADS, 0, st>>>(*index, *q, vec, counts, tile_off, out_q, out_s, s_label);
    else
        k_emit<IV_MODE_CONTAIN><<<(unsigned)ntiles, OV_THREADS, 0, st>>>(*index, *q, vec, counts, tile_off, out_q, out_s, s_label);
    IV_LAUNCH_CHECK();
    return IV_OK;
}

extern "C" int iv_coverage_bp(const iv_index_t* index, const iv_queries_t* q, int64_t* out_bp, double* out_fraction,
                              void* stream) {
    cudaStream_t st = (cudaStream_t)stream;
    int rc = check_query(index, q, "iv_coverage_bp");
    if (rc != IV_OK) return rc;
    if (q->n == 0) return IV_OK;
    k_coverage<<<(unsigned)cdiv(q->n, OV_THREADS), OV_THREADS, 0, st>>>(*index, *q, out_bp, out_fraction);
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
    k_nearest<<<(unsigned)cdiv(q->n, OV_THREADS), OV_THREADS, 0, st>>>(*index, *q, overlap, out_row, out_dist);
    IV_LAUNCH_CHECK();
    return IV_OK;
}
