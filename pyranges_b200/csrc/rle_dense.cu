// rle_dense.cu -- coverage run-length encoding of dense read sets without a sort of 64-bit keys.
//
// Same contract as the sort-based path in rle.cu (pyrle.methods.coverage, reference call site
// pyranges/methods/to_rle.py:18, semantics pinned by the doctests pyranges/pyranges_main.py:5800-5872), chosen by
// iv_rle_count when the reads are dense enough (at least one event per 64 coordinates of the flattened axis, span
// between 2^27 and 2^33: whole-genome read sets).  The flattened axis is cut into buckets of W = 8192 coordinates;
// an event (start: +1, end: -1) is then fully described by its bucket and a 14-bit word (offset << 1 | is_end):
//
//   prep    ONE read of Start / End: histogram of the events over the (d2, d1) = upper digits of the bucket id
//           (shared-memory table, <= 8192 counters) + exact first Start / last End of every group
//   pass A  partition by d2 (<= 64 bins): tiles of 4096 events generated straight from the rows, ranked in registers,
//           staged in shared memory, decoupled look-back for the tile offsets; writes 32-bit words (d1 d0 | 14 bits)
//   pass B  partition by d1 (128 bins) inside every d2 segment (tiles aligned to the segments, the look-back stops
//           at the segment's first tile); writes 32-bit words (d0 | 14 bits)
//   pass C  one block per (d2, d1) segment: histogram over d0 = the bucket starts, then the same tile body with a
//           running cursor; writes 16-bit words
//   dense   per bucket a difference array of W counters in shared memory (atomics), a scan over the W coordinates
//           = the coverage level at every coordinate; a coordinate starts a run iff the level changes there (plus
//           the two rules about coordinate 0 / the first Start of a group), the group's last End closes the last run.
//           Run once to count (per bucket: runs, net level change, last run start; per group: runs before its
//           origin), scanned over the buckets, and once to write: every run start writes its value and the LENGTH
//           OF THE RUN BEFORE IT (own coordinate - previous run start), so lengths need no pass of their own.
//
// Bytes per read at 200 M reads: 16 (prep) + 16 + 8 (A) + 8 + 8 (B) + 8 + 4 (C) + 4 (count) + 4 + 16 R/N (emit),
// against 16 + 16 + 5 x 48 (sort of 64-bit events) + 3 sweeps of the sorted events in the sort-based path.
#include "common.cuh"

namespace iv {

constexpr int RD_L = 13;                 // log2(coordinates per bucket)
constexpr int RD_W = 1 << RD_L;
constexpr int RD_LW = RD_L + 1;          // bits of an event word inside its bucket: offset << 1 | is_end
constexpr int RD_W0 = 7, RD_W1 = 7;      // digit widths of the bucket id, low to high; w2 = T - 14 (1 .. 6)
constexpr int RD_MIN_T = 15, RD_MAX_T = 20;
constexpr int RD_THREADS = 256;
constexpr int RD_ITEMS = 16;
constexpr int RD_TILE = RD_THREADS * RD_ITEMS;  // 4096 events per tile
constexpr int RD_BINS = 128;
constexpr int RD_WARPS = RD_THREADS / 32;
constexpr int RD_CHUNK = 32;             // consecutive buckets a block of the dense pass takes per ticket
constexpr int RD_NG = 16;                // groups of a chunk kept in shared memory (more: global lookups)
constexpr int RD_PF = 4;                 // events per thread of the NEXT bucket loaded ahead
constexpr int RD_CSTRIDE = 32;           // uint32 words per claim cursor: one cursor per 128-byte line

struct RdPlan {
    bool use;
    int T, w2;
    int64_t nb;  // buckets
};
static RdPlan rd_plan(int64_t n, int64_t total_span) {
    RdPlan p;
    p.use = false;
    p.T = p.w2 = 0;
    p.nb = 0;
    if (n <= 0 || total_span <= 0) return p;
    const int64_t m = 2 * n;
    p.nb = (total_span >> RD_L) + 1;
    p.T = bits_for((uint64_t)(p.nb - 1));
    p.w2 = p.T - RD_W0 - RD_W1;
    p.use = p.T >= RD_MIN_T && p.T <= RD_MAX_T && m < ((int64_t)1 << 30) && total_span / 64 <= m;
    return p;
}

struct RdLayout {
    uint32_t *wa, *wb;        // event words after pass A / pass B; pass C writes its 16-bit words over wa
    uint32_t *joint;          // [2^(w1+w2)] events per (d2, d1) segment -> exclusive starts (+1 entry)
    uint32_t *start2, *toff2; // [2^w2 + 1] first event / first pass-B tile of every d2 segment
    uint32_t *curA, *curB;    // claim cursors of pass A / pass B, one per 128-byte line
    uint32_t *tickets;        // [8]
    uint32_t *bstart;         // [2^T + 1] first event of every bucket
    int64_t *gF, *gZ;         // [n_groups] first Start / last End (flattened)
    int64_t *bcnt, *bsum, *blast;  // per bucket: runs, net level change, last run start -> their exclusive scans
    int64_t *tagg;            // per 2048-bucket tile: the three aggregates, then their exclusive scans
    int64_t *glocal;          // [n_groups] runs of the origin's bucket that start before the origin
    size_t zero_from, zero_to;  // joint .. tickets: cleared before every run
};
constexpr int RD_BSCAN = 2048;

static void rd_layout(int64_t n, int n_groups, Bump& b, RdLayout& L) {
    const size_t m = (size_t)(n > 0 ? 2 * n : 2);
    const size_t nbmax = ((size_t)1 << RD_MAX_T) + 64;
    L.wa = b.take<uint32_t>(m);
    L.wb = b.take<uint32_t>(m);
    L.zero_from = b.off;
    L.joint = b.take<uint32_t>((1u << (RD_W1 + 6)) + 64);
    L.curA = b.take<uint32_t>(64 * RD_CSTRIDE);
    L.curB = b.take<uint32_t>(((size_t)1 << (RD_W1 + 6)) * RD_CSTRIDE);
    L.tickets = b.take<uint32_t>(64);
    L.bcnt = b.take<int64_t>(nbmax);
    L.bsum = b.take<int64_t>(nbmax);
    L.zero_to = b.off;
    L.blast = b.take<int64_t>(nbmax);
    L.start2 = b.take<uint32_t>(128);
    L.toff2 = b.take<uint32_t>(128);
    L.bstart = b.take<uint32_t>(nbmax);
    L.gF = b.take<int64_t>((size_t)n_groups + 2);
    L.gZ = b.take<int64_t>((size_t)n_groups + 2);
    L.tagg = b.take<int64_t>(6 * (nbmax / RD_BSCAN + 2));
    L.glocal = b.take<int64_t>((size_t)n_groups + 2);
}

// ---- prep -------------------------------------------------------------------------------------------------------------
__global__ void k_rd_init(int64_t* gF, int64_t* gZ, int n_groups) {
    const int g = blockIdx.x * blockDim.x + threadIdx.x;
    if (g < n_groups) {
        gF[g] = INT64_MAX;
        gZ[g] = INT64_MIN;
    }
}

__global__ void __launch_bounds__(RD_THREADS)
k_rd_prep(const int64_t* __restrict__ s, const int64_t* __restrict__ e, int64_t n, iv_groups_t G, int nseg,
          uint32_t* __restrict__ joint, int64_t* __restrict__ gF, int64_t* __restrict__ gZ) {
    extern __shared__ uint32_t rd_h[];  // [nseg]
    __shared__ int span[2];
    __shared__ int64_t red[2 * RD_WARPS];
    for (int i = threadIdx.x; i < nseg; i += RD_THREADS) rd_h[i] = 0;
    const int64_t nchunks = (n + RD_THREADS * 4 - 1) / (RD_THREADS * 4);
    for (int64_t c = blockIdx.x; c < nchunks; c += gridDim.x) {
        const int64_t first = c * (RD_THREADS * 4);
        const int64_t last = first + RD_THREADS * 4 - 1 < n - 1 ? first + RD_THREADS * 4 - 1 : n - 1;
        __syncthreads();
        SegSpan sp = block_seg_span(G.row_off, G.n_groups, first, last, span);
        int64_t mn = INT64_MAX, mx = INT64_MIN;
        int64_t va[4], vb[4];
#pragma unroll
        for (int j = 0; j < 4; j++) {
            const int64_t i = first + j * RD_THREADS + threadIdx.x;
            va[j] = i <= last ? __ldg(s + i) : 0;
            vb[j] = i <= last ? __ldg(e + i) : 0;
        }
        const int64_t base0 = __ldg(G.base + sp.klo);
#pragma unroll
        for (int j = 0; j < 4; j++) {
            const int64_t i = first + j * RD_THREADS + threadIdx.x;
            if (i <= last) {
                const int g = sp.of(G.row_off, i);
                const int64_t off = sp.klo == sp.khi ? base0 : __ldg(G.base + g);
                const int64_t a = va[j] + off, b = vb[j] + off;
                // (coordinates are >= 0 by contract; the clamp keeps a violation from leaving the table)
                const uint32_t ia = (uint32_t)(a >> (RD_L + RD_W0)), ib = (uint32_t)(b >> (RD_L + RD_W0));
                atomicAdd(&rd_h[ia < (uint32_t)nseg ? ia : (uint32_t)nseg - 1u], 1u);
                atomicAdd(&rd_h[ib < (uint32_t)nseg ? ib : (uint32_t)nseg - 1u], 1u);
                if (sp.klo == sp.khi) {
                    mn = a < mn ? a : mn;
                    mx = b > mx ? b : mx;
                } else {
                    atomicMin((long long*)gF + g, (long long)a);
                    atomicMax((long long*)gZ + g, (long long)b);
                }
            }
        }
        if (sp.klo == sp.khi) {
            mn = warp_min(mn);
            mx = warp_max(mx);
            if (lane_id() == 0) {
                red[threadIdx.x >> 5] = mn;
                red[RD_WARPS + (threadIdx.x >> 5)] = mx;
            }
            __syncthreads();
            if (threadIdx.x == 0) {
                for (int k = 1; k < RD_WARPS; k++) {
                    mn = red[k] < mn ? red[k] : mn;
                    mx = red[RD_WARPS + k] > mx ? red[RD_WARPS + k] : mx;
                }
                atomicMin((long long*)gF + sp.klo, (long long)mn);
                atomicMax((long long*)gZ + sp.klo, (long long)mx);
            }
        }
    }
    __syncthreads();
    for (int i = threadIdx.x; i < nseg; i += RD_THREADS) {
        const uint32_t v = rd_h[i];
        if (v) atomicAdd(joint + i, v);
    }
}

// single block: exclusive scan of the (d2, d1) table in place (+ the total behind it); first event and first pass-B
// tile of every d2 segment; the bucket-start entry behind the last bucket
__global__ void __launch_bounds__(1024)
k_rd_scan(uint32_t* __restrict__ joint, int nseg, int w1, uint32_t* __restrict__ start2, uint32_t* __restrict__ toff2,
          uint32_t* __restrict__ bstart_end) {
    __shared__ uint32_t sm[33];
    uint32_t carry = 0;
    for (int b0 = 0; b0 < nseg; b0 += 1024) {
        const int i = b0 + threadIdx.x;
        const uint32_t v = i < nseg ? joint[i] : 0u;
        uint32_t tot;
        const uint32_t ex = block_excl_sum<uint32_t, 1024>(v, sm, tot);
        if (i < nseg) joint[i] = carry + ex;
        carry += tot;
    }
    if (threadIdx.x == 0) {
        joint[nseg] = carry;
        *bstart_end = carry;
    }
    __syncthreads();
    const int n2 = nseg >> w1;
    if (threadIdx.x == 0) {
        uint32_t t = 0;
        for (int d = 0; d < n2; d++) {
            const uint32_t a = joint[d << w1], b = joint[(d + 1) << w1];
            start2[d] = a;
            toff2[d] = t;
            t += (b - a + RD_TILE - 1) / RD_TILE;
        }
        start2[n2] = carry;
        toff2[n2] = t;
    }
}

// ---- the tile body of the partition passes --------------------------------------------------------------------------
struct PtSmem {
    uint32_t warp_cnt[RD_WARPS][RD_BINS + 1];
    uint32_t digit_base[RD_BINS + 1];
    uint32_t gbase[RD_BINS];
    uint32_t sbuf[RD_TILE];
    uint8_t sdig[RD_TILE];
    uint32_t wsum[RD_WARPS + 1];
    uint32_t s_tile;
};

// word[it] / dig[it] of item li = warp * 512 + it * 32 + lane; items >= nvalid carry digit RD_BINS (a bin of their own
// behind every real digit), so they land behind every real item of the tile.  global_base(run, ex): called by thread d < RD_BINS for digit d, returns the output
// position of the tile's first item of that digit minus ex.  S.warp_cnt must be zero on entry.
template <typename OutT, typename GB>
__device__ __forceinline__ void pt_tile(PtSmem& S, const uint32_t (&word)[RD_ITEMS], const uint32_t (&dig)[RD_ITEMS],
                                        int nvalid, OutT* __restrict__ out, GB global_base) {
    const int tid = threadIdx.x, w = tid >> 5, lane = tid & 31;
    // rank inside (warp, digit) from a shared-memory atomic: the order of the events inside a bucket is irrelevant
    // (they are summed into a difference array), so the partition need not be stable
    uint32_t rank[RD_ITEMS];
#pragma unroll
    for (int it = 0; it < RD_ITEMS; it++) rank[it] = atomicAdd(&S.warp_cnt[w][dig[it]], 1u);
    __syncthreads();
    {
        uint32_t run = 0;
        if (tid <= RD_BINS) {
#pragma unroll
            for (int ww = 0; ww < RD_WARPS; ww++) {
                const uint32_t c = S.warp_cnt[ww][tid];
                S.warp_cnt[ww][tid] = run;
                run += c;
            }
        }
        uint32_t tot;
        const uint32_t ex = block_excl_sum<uint32_t, RD_THREADS>(run, S.wsum, tot);
        if (tid <= RD_BINS) S.digit_base[tid] = ex;
        if (tid < RD_BINS) S.gbase[tid] = global_base(run, ex);
    }
    __syncthreads();
#pragma unroll
    for (int it = 0; it < RD_ITEMS; it++) {
        const uint32_t p = S.digit_base[dig[it]] + S.warp_cnt[w][dig[it]] + rank[it];
        S.sbuf[p] = word[it];
        S.sdig[p] = (uint8_t)dig[it];
    }
    __syncthreads();
    for (int k = tid; k < nvalid; k += RD_THREADS) out[S.gbase[S.sdig[k]] + (uint32_t)k] = (OutT)S.sbuf[k];
}

// Output space of a tile: one atomic claim per (tile, digit) on a cursor that owns a 128-byte line (the claims of the
// ~600 tiles in flight go to the same few digits; cursors sharing a line would serialise them).  The claims make the
// order of the tiles inside a digit -- not the result -- depend on timing: events are summed, never ordered.
__device__ __forceinline__ uint32_t pt_claim(uint32_t* cursors, uint32_t idx, uint32_t run) {
    return run ? atomicAdd(cursors + (size_t)idx * RD_CSTRIDE, run) : 0u;
}

// pass A: events straight from the rows, partitioned by d2
__global__ void __launch_bounds__(RD_THREADS, 4)
k_rd_pass_a(const int64_t* __restrict__ s, const int64_t* __restrict__ e, int64_t n, iv_groups_t G, int w2,
            const uint32_t* __restrict__ start2, uint32_t* cursors, uint32_t* __restrict__ out) {
    __shared__ PtSmem S;
    __shared__ int span[2];
    const int tid = threadIdx.x, w = tid >> 5, lane = tid & 31;
    for (int i = tid; i < RD_WARPS * (RD_BINS + 1); i += RD_THREADS) (&S.warp_cnt[0][0])[i] = 0;
    __syncthreads();
    const int64_t tile = blockIdx.x;
    const int64_t m = 2 * n, tbase = tile * RD_TILE;
    const int nvalid = (int)(m - tbase < (int64_t)RD_TILE ? m - tbase : (int64_t)RD_TILE);
    const int64_t r0 = tbase >> 1, r1 = (tbase + nvalid - 1) >> 1;
    SegSpan sp = block_seg_span(G.row_off, G.n_groups, r0, r1, span);
    const int nb2 = 1 << w2;
    uint32_t word[RD_ITEMS], dig[RD_ITEMS];
    int64_t xs[RD_ITEMS];
#pragma unroll
    for (int it = 0; it < RD_ITEMS; it++) {  // every load of the tile in flight before anything depends on one
        const int li = w * (RD_ITEMS * 32) + it * 32 + lane;
        const int64_t row = (tbase + li) >> 1;
        xs[it] = li < nvalid ? ((li & 1) ? __ldg(e + row) : __ldg(s + row)) : 0;
    }
    const int64_t base0 = __ldg(G.base + sp.klo);
#pragma unroll
    for (int it = 0; it < RD_ITEMS; it++) {
        const int li = w * (RD_ITEMS * 32) + it * 32 + lane;
        word[it] = 0;
        dig[it] = RD_BINS;
        if (li < nvalid) {
            const int64_t x = xs[it] + (sp.klo == sp.khi ? base0 : __ldg(G.base + sp.of(G.row_off, (tbase + li) >> 1)));
            const uint32_t bucket = (uint32_t)(x >> RD_L);
            word[it] = ((bucket & ((1u << (RD_W0 + RD_W1)) - 1u)) << RD_LW) | (((uint32_t)x & (RD_W - 1)) << 1) | (li & 1);
            dig[it] = bucket >> (RD_W0 + RD_W1);
        }
    }
    pt_tile<uint32_t>(S, word, dig, nvalid, out, [&](uint32_t run, uint32_t ex) {
        uint32_t excl = 0;
        if (tid < nb2) excl = pt_claim(cursors, tid, run) + __ldg(start2 + tid);
        return excl - ex;
    });
}

// pass B: partition by d1 inside every d2 segment (tiles aligned to the segments)
__global__ void __launch_bounds__(RD_THREADS, 4)
k_rd_pass_b(const uint32_t* __restrict__ in, int w2, const uint32_t* __restrict__ start2,
            const uint32_t* __restrict__ toff2, const uint32_t* __restrict__ joint, uint32_t* cursors,
            uint32_t* __restrict__ out) {
    __shared__ PtSmem S;
    __shared__ uint32_t s_seg;
    const int tid = threadIdx.x, w = tid >> 5, lane = tid & 31;
    const uint32_t t = blockIdx.x;
    if (tid == 0) {
        const int n2 = 1 << w2;
        s_seg = 0xffffffffu;
        if (t < __ldg(toff2 + n2)) {
            int lo = 0, hi = n2;  // largest d with toff2[d] <= t (segments without events share their offset with the next)
            while (hi - lo > 1) {
                const int mid = (lo + hi) >> 1;
                if (__ldg(toff2 + mid) <= t) lo = mid; else hi = mid;
            }
            s_seg = (uint32_t)lo;
        }
    }
    for (int i = tid; i < RD_WARPS * (RD_BINS + 1); i += RD_THREADS) (&S.warp_cnt[0][0])[i] = 0;
    __syncthreads();
    if (s_seg == 0xffffffffu) return;
    const uint32_t d2 = s_seg;
    const uint32_t first_tile = __ldg(toff2 + d2);
    const uint32_t seg_lo = __ldg(start2 + d2), seg_hi = __ldg(start2 + d2 + 1);
    const uint32_t tbase = seg_lo + (t - first_tile) * RD_TILE;
    const int nvalid = (int)(seg_hi - tbase < (uint32_t)RD_TILE ? seg_hi - tbase : (uint32_t)RD_TILE);
    uint32_t word[RD_ITEMS], dig[RD_ITEMS];
#pragma unroll
    for (int it = 0; it < RD_ITEMS; it++) {
        const int li = w * (RD_ITEMS * 32) + it * 32 + lane;
        const uint32_t v = li < nvalid ? __ldg(in + tbase + li) : 0u;
        word[it] = v & ((1u << (RD_LW + RD_W0)) - 1u);
        dig[it] = li < nvalid ? (v >> (RD_LW + RD_W0)) : (uint32_t)RD_BINS;
    }
    pt_tile<uint32_t>(S, word, dig, nvalid, out, [&](uint32_t run, uint32_t ex) {
        const uint32_t seg = (d2 << RD_W1) | (uint32_t)tid;
        return __ldg(joint + seg) + pt_claim(cursors, seg, run) - ex;
    });
}

// pass C: one block per (d2, d1) segment: bucket starts + partition by d0 into 16-bit words
__global__ void __launch_bounds__(RD_THREADS, 4)
k_rd_pass_c(const uint32_t* __restrict__ in, const uint32_t* __restrict__ joint, uint32_t* __restrict__ bstart,
            uint16_t* __restrict__ out) {
    __shared__ PtSmem S;
    __shared__ uint32_t cursor[RD_BINS];
    const int tid = threadIdx.x, w = tid >> 5, lane = tid & 31;
    const uint32_t seg = blockIdx.x;
    const uint32_t lo = __ldg(joint + seg), hi = __ldg(joint + seg + 1);
    if (tid < RD_BINS) cursor[tid] = 0;
    __syncthreads();
    for (uint32_t i0 = lo; i0 < hi; i0 += 8 * RD_THREADS) {  // eight loads in flight per thread
        uint32_t v[8];
#pragma unroll
        for (int j = 0; j < 8; j++) {
            const uint32_t i = i0 + j * RD_THREADS + tid;
            v[j] = i < hi ? __ldg(in + i) : 0xffffffffu;
        }
#pragma unroll
        for (int j = 0; j < 8; j++)
            if (v[j] != 0xffffffffu) atomicAdd(&cursor[v[j] >> RD_LW], 1u);
    }
    __syncthreads();
    {
        const uint32_t v = tid < RD_BINS ? cursor[tid] : 0u;
        uint32_t tot;
        const uint32_t ex = block_excl_sum<uint32_t, RD_THREADS>(v, S.wsum, tot);
        if (tid < RD_BINS) {
            cursor[tid] = lo + ex;
            bstart[(seg << RD_W0) + tid] = lo + ex;
        }
    }
    __syncthreads();
    for (uint32_t tbase = lo; tbase < hi; tbase += RD_TILE) {
        for (int i = tid; i < RD_WARPS * (RD_BINS + 1); i += RD_THREADS) (&S.warp_cnt[0][0])[i] = 0;
        __syncthreads();
        const int nvalid = (int)(hi - tbase < (uint32_t)RD_TILE ? hi - tbase : (uint32_t)RD_TILE);
        uint32_t word[RD_ITEMS], dig[RD_ITEMS];
#pragma unroll
        for (int it = 0; it < RD_ITEMS; it++) {
            const int li = w * (RD_ITEMS * 32) + it * 32 + lane;
            const uint32_t v = li < nvalid ? __ldg(in + tbase + li) : 0u;
            word[it] = v & ((1u << RD_LW) - 1u);
            dig[it] = li < nvalid ? (v >> RD_LW) : (uint32_t)RD_BINS;
        }
        pt_tile<uint16_t>(S, word, dig, nvalid, out, [&](uint32_t run, uint32_t ex) {
            const uint32_t g = cursor[tid];
            cursor[tid] = g + run;
            return g - ex;
        });
        __syncthreads();
    }
}

// ---- the dense pass ---------------------------------------------------------------------------------------------------
// difference array with 4 words of padding per 32: thread t scans words [32 t, 32 t + 32) with conflict-free 128-bit
// loads (a quarter-warp touches 8 different bank quads)
__device__ __forceinline__ int rd_pad(int i) { return i + ((i >> 5) << 2); }
constexpr int RD_DELTA_WORDS = RD_W + (RD_W >> 5) * 4;

struct RdGroups {  // groups overlapping the current chunk of buckets
    int glo, ghi;  // ghi - glo + 1 <= RD_NG: tables below are valid; else global lookups
    int64_t base[RD_NG + 1], F[RD_NG], Z[RD_NG];
};

constexpr int RD_STAGE = 1024;  // runs of a bucket staged in shared memory for coalesced stores (more: direct stores)

template <bool EMIT>
__global__ void __launch_bounds__(RD_THREADS, 4)
k_rd_dense(const uint16_t* __restrict__ ev, const uint32_t* __restrict__ bstart, int64_t nb, iv_groups_t G,
           const int64_t* __restrict__ gF, const int64_t* __restrict__ gZ, uint32_t* ticket,
           int64_t* __restrict__ bcnt, int64_t* __restrict__ bsum, int64_t* __restrict__ blast,
           int64_t* __restrict__ glocal, int64_t* __restrict__ out_runs, double* __restrict__ out_values) {
    __shared__ __align__(16) int delta[RD_DELTA_WORDS];
    __shared__ uint32_t sstart[RD_CHUNK + 1];
    __shared__ RdGroups SG;
    __shared__ uint32_t s_chunk;
    __shared__ uint8_t sflag[RD_CHUNK];  // bit 0: a group origin inside the bucket, bit 1: a group end inside it
    __shared__ int64_t s_scan[EMIT ? 3 * RD_CHUNK : 1];  // the chunk's scanned bucket values: runs, level, last run start
    __shared__ int4 wred[RD_WARPS];
    __shared__ int st_len[EMIT ? RD_STAGE : 1];  // staged as 32-bit words (a longer run is stored directly)
    __shared__ int st_val[EMIT ? RD_STAGE : 1];
    const int tid = threadIdx.x, w = tid >> 5, lane = tid & 31;
    const int64_t nchunks = (nb + RD_CHUNK - 1) / RD_CHUNK;
    {   // every thread clears the words it scans after each bucket; the padding words are never written
        int4* d4 = reinterpret_cast<int4*>(delta);
        for (int i = tid; i < RD_DELTA_WORDS / 4; i += RD_THREADS) d4[i] = make_int4(0, 0, 0, 0);
    }
    while (true) {
        __syncthreads();
        if (tid == 0) s_chunk = atomicAdd(ticket, 1u);
        __syncthreads();
        const int64_t chunk = s_chunk;
        if (chunk >= nchunks) return;
        const int64_t b_lo = chunk * RD_CHUNK, b_hi = b_lo + RD_CHUNK < nb ? b_lo + RD_CHUNK : nb;
        if (tid <= (int)(b_hi - b_lo)) sstart[tid] = __ldg(bstart + b_lo + tid);
        if (EMIT && tid >= 64 && tid < 64 + (int)(b_hi - b_lo)) {
            s_scan[tid - 64] = bcnt[b_lo + tid - 64];
            s_scan[RD_CHUNK + tid - 64] = bsum[b_lo + tid - 64];
            s_scan[2 * RD_CHUNK + tid - 64] = blast[b_lo + tid - 64];
        }
        if (tid == 32) {
            const int64_t x0 = b_lo << RD_L, x1 = (b_hi << RD_L) - 1;
            const int glo = find_seg(G.base, G.n_groups, x0), ghi = find_seg(G.base, G.n_groups, x1);
            SG.glo = glo;
            SG.ghi = ghi;
            if (ghi - glo < RD_NG) {
                for (int g = glo; g <= ghi; g++) {
                    SG.base[g - glo] = __ldg(G.base + g);
                    SG.F[g - glo] = __ldg(gF + g);
                    SG.Z[g - glo] = __ldg(gZ + g);
                }
                SG.base[ghi - glo + 1] = __ldg(G.base + ghi + 1);
            }
        }
        __syncthreads();
        const bool gfast = SG.ghi - SG.glo < RD_NG;
        if (tid < (int)(b_hi - b_lo)) {  // one thread per bucket of the chunk: does it hold a group origin / a group end?
            const int64_t x0 = (b_lo + tid) << RD_L;
            int f = 0;
            if (gfast) {
                for (int k = 0; k <= SG.ghi - SG.glo; k++) {
                    if (SG.base[k] >= x0 && SG.base[k] < x0 + RD_W) f |= 1;
                    if (SG.Z[k] >= x0 && SG.Z[k] < x0 + RD_W) f |= 2;
                }
            } else {
                const int g0 = find_seg(G.base, G.n_groups, x0);
                for (int g = g0; g < G.n_groups && __ldg(G.base + g) < x0 + RD_W; g++) {
                    if (__ldg(G.base + g) >= x0) f |= 1;
                    const int64_t Z = __ldg(gZ + g);
                    if (Z >= x0 && Z < x0 + RD_W) f |= 2;
                }
            }
            sflag[tid] = (uint8_t)f;
        }
        __syncthreads();
        // events of the first bucket of the chunk, RD_PF per thread ahead
        uint32_t pf[RD_PF];
        {
            const uint32_t e0 = sstart[0], e1 = sstart[1];
#pragma unroll
            for (int j = 0; j < RD_PF; j++) {
                const uint32_t i = e0 + tid + j * RD_THREADS;
                pf[j] = i < e1 ? (uint32_t)__ldg(ev + i) : 0xffffffffu;
            }
        }
        for (int64_t b = b_lo; b < b_hi; b++) {
            const int bi = (int)(b - b_lo);
            const uint32_t e0 = sstart[bi], e1 = sstart[bi + 1];
            const int64_t x0 = b << RD_L;
            uint32_t cur[RD_PF];
#pragma unroll
            for (int j = 0; j < RD_PF; j++) cur[j] = pf[j];
            if (b + 1 < b_hi) {
                const uint32_t n0 = sstart[bi + 1], n1 = sstart[bi + 2];
#pragma unroll
                for (int j = 0; j < RD_PF; j++) {
                    const uint32_t i = n0 + tid + j * RD_THREADS;
                    pf[j] = i < n1 ? (uint32_t)__ldg(ev + i) : 0xffffffffu;
                }
            }
            // a bucket without events and without a group origin has nothing to do
            const int bflag = sflag[bi];
            if (e1 == e0 && !(bflag & 1)) continue;
            __syncthreads();  // delta[] is clear again, the staged runs of the previous bucket are out
#pragma unroll
            for (int j = 0; j < RD_PF; j++)
                if (cur[j] != 0xffffffffu) atomicAdd(&delta[rd_pad((int)(cur[j] >> 1))], (cur[j] & 1u) ? -1 : 1);
            for (uint32_t i = e0 + RD_PF * RD_THREADS + tid; i < e1; i += RD_THREADS) {
                const uint32_t v = __ldg(ev + i);
                atomicAdd(&delta[rd_pad((int)(v >> 1))], (v & 1u) ? -1 : 1);
            }
            __syncthreads();
            // thread t: coordinates [x0 + 32 t, x0 + 32 t + 32), all arithmetic relative to x0 in 32 bits
            const int xr0 = tid * 32;
            int* dmine = delta + tid * 36;
            const int4* d4 = reinterpret_cast<const int4*>(dmine);
            // group context: one group unless an origin falls inside the thread's coordinates
            int g = 0, gb = 0, gnext = 0, Zr = 0;
            int64_t gb64 = 0;
            bool nonempty = false, lead = false, zero0 = false;  // rows; first Start > 0; first Start == 0
            auto rel = [&](int64_t v) {
                const int64_t d = v - x0;
                return d < -(1 << 30) ? -(1 << 30) : (d > (1 << 30) ? (1 << 30) : (int)d);
            };
            auto load_group = [&](int xr) {
                const int64_t x = x0 + xr;
                int64_t nx, F, Z;
                if (gfast) {
                    int k = 0;
                    while (k < SG.ghi - SG.glo && SG.base[k + 1] <= x) k++;
                    g = SG.glo + k;
                    gb64 = SG.base[k];
                    nx = SG.base[k + 1];
                    F = SG.F[k];
                    Z = SG.Z[k];
                } else {
                    g = find_seg(G.base, G.n_groups, x);
                    gb64 = __ldg(G.base + g);
                    nx = __ldg(G.base + g + 1);
                    F = __ldg(gF + g);
                    Z = __ldg(gZ + g);
                }
                nonempty = F <= Z;
                lead = nonempty && F > gb64;
                zero0 = nonempty && F == gb64;
                gb = rel(gb64);
                gnext = x >= nx ? (1 << 30) : rel(nx);  // behind the last group: nothing more to look up
                Zr = nonempty ? rel(Z) : -(1 << 30);
            };
            // a bucket without origin and without group end lies inside one group whose first run starts before it:
            // no group context is needed at all
            const bool bplain = bflag == 0;
            if (!bplain) load_group(xr0);
            // pass 1 over the 32 coordinates: markers (run starts / group ends), level change
            uint32_t emit_mask = 0, term_mask = 0, org_mask = 0;
            int dsum = 0;
            const bool plain = bplain || (gnext > xr0 + 31 && gb < xr0 && (Zr < xr0 || Zr > xr0 + 31));
            if (plain) {  // no origin, no group end among the 32 coordinates: a run starts wherever the level changes
#pragma unroll
                for (int q = 0; q < 8; q++) {
                    const int4 v = d4[q];
                    dsum += v.x + v.y + v.z + v.w;
                    emit_mask |= ((v.x != 0 ? 1u : 0u) | (v.y != 0 ? 2u : 0u) | (v.z != 0 ? 4u : 0u) | (v.w != 0 ? 8u : 0u)) << (4 * q);
                }
            } else {
#pragma unroll 1
                for (int j = 0; j < 32; j++) {
                    const int xr = xr0 + j, dv = dmine[j];
                    if (xr >= gnext) load_group(xr);
                    dsum += dv;
                    const bool at0 = xr == gb;
                    if (at0) org_mask |= 1u << j;
                    if ((at0 && lead) || (nonempty && xr != Zr && (dv != 0 || (at0 && zero0)))) emit_mask |= 1u << j;
                    if (xr == Zr) term_mask |= 1u << j;
                }
            }
            const int cnt = __popc(emit_mask);
            const int my_last = emit_mask ? xr0 + (31 - __clz(emit_mask)) : -1;   // relative to x0
            const int my_term = term_mask ? xr0 + (31 - __clz(term_mask)) : -1;
            // block scans (32-bit): exclusive run count, exclusive level, last run start before the thread
            const int c_inc = warp_incl_sum(cnt), s_inc = warp_incl_sum(dsum), l_inc = warp_incl_max(my_last);
            const int t_w = bplain ? -1 : warp_max(my_term);
            if (lane == 31) wred[w] = make_int4(c_inc, s_inc, l_inc, t_w);
            __syncthreads();
            int c_ex32 = c_inc - cnt, s_ex32 = s_inc - dsum;
            int l_ex32 = __shfl_up_sync(FULL, l_inc, 1);
            if (lane == 0) l_ex32 = -1;
            int c_tot32 = 0, s_tot32 = 0, l_tot32 = -1, t_tot32 = -1;
#pragma unroll
            for (int k = 0; k < RD_WARPS; k++) {
                const int4 r = wred[k];
                if (k < w) {
                    c_ex32 += r.x;
                    s_ex32 += r.y;
                    l_ex32 = r.z > l_ex32 ? r.z : l_ex32;
                }
                c_tot32 += r.x;
                s_tot32 += r.y;
                l_tot32 = r.z > l_tot32 ? r.z : l_tot32;
                t_tot32 = r.w > t_tot32 ? r.w : t_tot32;
            }
            const int64_t c_ex = c_ex32, s_ex = s_ex32, c_tot = c_tot32, s_tot = s_tot32;
            const int64_t l_ex = l_ex32 >= 0 ? x0 + l_ex32 : -1, l_tot = l_tot32 >= 0 ? x0 + l_tot32 : -1;
            const int64_t t_tot = t_tot32 >= 0 ? x0 + t_tot32 : -1;
            if (!EMIT) {
                if (tid == 0) {
                    bcnt[b] = c_tot;
                    bsum[b] = s_tot;
                    blast[b] = l_tot;
                }
                // runs of this bucket that start before a group origin inside it
                while (org_mask) {
                    const int j = __ffs(org_mask) - 1;
                    org_mask &= org_mask - 1;
                    load_group(xr0 + j);
                    glocal[g] = c_ex + __popc(emit_mask & ((1u << j) - 1u));
                }
            } else {
                // every marker closes the run before it (length = own coordinate - previous run start); a run start
                // also opens the next one with the level behind it.  The bucket's runs are staged in shared memory
                // and written as coalesced rows; the run before the bucket's first one is closed with a direct store.
                const int64_t base_r = s_scan[bi];           // scanned: runs before this bucket
                const bool staged = c_tot <= RD_STAGE;
                if (bplain && staged) {
                    // the common case in 32-bit arithmetic relative to the bucket: every marker is a run start, its
                    // predecessor lies in the same group, and only the bucket's first run start looks outside the bucket
                    uint32_t marks = emit_mask;
                    int level = (int)s_scan[RD_CHUNK + bi] + s_ex32, prevr = l_ex32, k = c_ex32;
                    while (marks) {
                        const int j = __ffs(marks) - 1;
                        marks &= marks - 1;
                        const int xr = xr0 + j;
                        level += dmine[j];
                        if (k > 0) st_len[k - 1] = xr - prevr;
                        else out_runs[base_r - 1] = (x0 + xr) - s_scan[2 * RD_CHUNK + bi];
                        st_val[k] = level;
                        prevr = xr;
                        k++;
                    }
                } else {
                    uint32_t marks = emit_mask | term_mask;
                    if (marks) {
                        int level = 0;
                        const int64_t level0 = s_scan[RD_CHUNK + bi] + s_ex;           // scanned: level before this bucket
                        int64_t prev = l_ex >= 0 ? l_ex : s_scan[2 * RD_CHUNK + bi];  // scanned: last run start before it
                        int k = (int)c_ex;                                             // runs of this bucket before the marker
                        if (!bplain && !plain) load_group(xr0);
                        while (marks) {
                            const int j = __ffs(marks) - 1;
                            marks &= marks - 1;
                            const int xr = xr0 + j;
                            if (!bplain && xr >= gnext) load_group(xr);
                            level += dmine[j];
                            const int64_t x = x0 + xr;
                            if (bplain || prev >= gb64) {
                                const int64_t len = x - prev;
                                if (staged && k > 0) st_len[k - 1] = len < INT32_MAX ? (int)len : -1;
                                if (!staged || k == 0 || len >= INT32_MAX) out_runs[base_r + k - 1] = len;
                            }
                            if (emit_mask >> j & 1u) {
                                const int64_t v = (!bplain && xr == gb && lead) ? 0 : level0 + level;
                                if (staged) st_val[k] = (int)v;
                                else out_values[base_r + k] = (double)v;
                                prev = x;
                                k++;
                            }
                        }
                    }
                }
                if (staged && c_tot > 0) {
                    __syncthreads();
                    // the bucket's last run is closed inside the bucket only if a group ends behind it
                    const int nlen = (int)c_tot - (t_tot > l_tot ? 0 : 1);
                    for (int i = tid; i < (int)c_tot; i += RD_THREADS) {
                        out_values[base_r + i] = (double)st_val[i];
                        if (i < nlen && st_len[i] >= 0) out_runs[base_r + i] = (int64_t)st_len[i];
                    }
                }
            }
            // clear the words this thread scanned (all events of the bucket are inside them)
            {
                int4* z4 = reinterpret_cast<int4*>(dmine);
#pragma unroll
                for (int q = 0; q < 8; q++) z4[q] = make_int4(0, 0, 0, 0);
            }
        }
    }
}

// ---- scan over the buckets: (runs, level change) summed, last run start max-ed; all exclusive, in place ---------------
__global__ void __launch_bounds__(256)
k_rd_bscan_tiles(const int64_t* __restrict__ bcnt, const int64_t* __restrict__ bsum, const int64_t* __restrict__ blast,
                 int64_t nb, int64_t* __restrict__ tagg, int64_t ntiles) {
    __shared__ int64_t sm[3 * 8];
    const int64_t t0 = (int64_t)blockIdx.x * RD_BSCAN;
    int64_t c = 0, s = 0, l = -1;
    for (int j = 0; j < RD_BSCAN / 256; j++) {
        const int64_t i = t0 + j * 256 + threadIdx.x;
        if (i < nb) {
            c += bcnt[i];
            s += bsum[i];
            const int64_t v = blast[i];
            l = v > l ? v : l;
        }
    }
    c = warp_sum(c);
    s = warp_sum(s);
    l = warp_max(l);
    if (lane_id() == 0) {
        sm[threadIdx.x >> 5] = c;
        sm[8 + (threadIdx.x >> 5)] = s;
        sm[16 + (threadIdx.x >> 5)] = l;
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        for (int k = 1; k < 8; k++) {
            c += sm[k];
            s += sm[8 + k];
            l = sm[16 + k] > l ? sm[16 + k] : l;
        }
        tagg[blockIdx.x] = c;
        tagg[ntiles + blockIdx.x] = s;
        tagg[2 * ntiles + blockIdx.x] = l;
    }
}
__global__ void __launch_bounds__(1024)
k_rd_bscan_mid(int64_t* __restrict__ tagg, int64_t ntiles, int64_t* __restrict__ total_runs) {
    __shared__ int64_t sm[33];
    __shared__ int64_t lprev[1024];
    const int64_t t = threadIdx.x;  // ntiles <= 513
    const int64_t vc = t < ntiles ? tagg[t] : 0, vs = t < ntiles ? tagg[ntiles + t] : 0;
    const int64_t vl = t < ntiles ? tagg[2 * ntiles + t] : -1;
    int64_t tot_c, tot;
    const int64_t c_ex = block_excl_sum<int64_t, 1024>(vc, sm, tot_c);
    const int64_t s_ex = block_excl_sum<int64_t, 1024>(vs, sm, tot);
    const int64_t l_inc = block_incl_max<int64_t, 1024>(vl, sm, (int64_t)-1, tot);
    lprev[threadIdx.x] = l_inc;
    __syncthreads();
    if (t < ntiles) {
        tagg[t] = c_ex;
        tagg[ntiles + t] = s_ex;
        tagg[2 * ntiles + t] = threadIdx.x ? lprev[threadIdx.x - 1] : -1;
    }
    if (threadIdx.x == 0) *total_runs = tot_c;
}
__global__ void __launch_bounds__(256)
k_rd_bscan_apply(int64_t* __restrict__ bcnt, int64_t* __restrict__ bsum, int64_t* __restrict__ blast, int64_t nb,
                 const int64_t* __restrict__ tagg, int64_t ntiles) {
    __shared__ int64_t sm[9];
    __shared__ int64_t sl[9];
    const int64_t t0 = (int64_t)blockIdx.x * RD_BSCAN + (int64_t)threadIdx.x * (RD_BSCAN / 256);
    int64_t vc[RD_BSCAN / 256], vs[RD_BSCAN / 256], vl[RD_BSCAN / 256];
    int64_t c = 0, s = 0, l = -1;
#pragma unroll
    for (int j = 0; j < RD_BSCAN / 256; j++) {
        const int64_t i = t0 + j;
        vc[j] = i < nb ? bcnt[i] : 0;
        vs[j] = i < nb ? bsum[i] : 0;
        vl[j] = i < nb ? blast[i] : -1;
        c += vc[j];
        s += vs[j];
        l = vl[j] > l ? vl[j] : l;
    }
    int64_t tot;
    int64_t c_ex = block_excl_sum<int64_t, 256>(c, sm, tot) + tagg[blockIdx.x];
    int64_t s_ex = block_excl_sum<int64_t, 256>(s, sm, tot) + tagg[ntiles + blockIdx.x];
    int64_t l_inc = block_incl_max<int64_t, 256>(l, sl, (int64_t)-1, tot);
    // exclusive max: shift by one thread
    __shared__ int64_t lprev[256];
    lprev[threadIdx.x] = l_inc;
    __syncthreads();
    int64_t l_ex = threadIdx.x ? lprev[threadIdx.x - 1] : -1;
    const int64_t tl = tagg[2 * ntiles + blockIdx.x];
    l_ex = tl > l_ex ? tl : l_ex;
#pragma unroll
    for (int j = 0; j < RD_BSCAN / 256; j++) {
        const int64_t i = t0 + j;
        if (i < nb) {
            bcnt[i] = c_ex;
            bsum[i] = s_ex;
            blast[i] = l_ex;
        }
        c_ex += vc[j];
        s_ex += vs[j];
        l_ex = vl[j] > l_ex ? vl[j] : l_ex;
    }
}
__global__ void k_rd_run_off(iv_groups_t G, const int64_t* __restrict__ bcnt, const int64_t* __restrict__ glocal,
                             const int64_t* __restrict__ total_runs, int64_t* __restrict__ out_run_off) {
    const int g = blockIdx.x * blockDim.x + threadIdx.x;
    if (g < G.n_groups) out_run_off[g] = bcnt[__ldg(G.base + g) >> RD_L] + glocal[g];
    if (g == G.n_groups) out_run_off[g] = *total_runs;
}

// ---- host side --------------------------------------------------------------------------------------------------------
size_t rle_dense_workspace_bytes(int64_t n, int n_groups) {
    Bump b(nullptr, 0);
    RdLayout L;
    rd_layout(n, n_groups, b, L);
    return b.off + 256;
}
bool rle_dense_applies(int64_t n, int64_t total_span) { return rd_plan(n, total_span).use; }

static int rd_dense_launch(bool emit, const RdPlan& pl, const RdLayout& L, const iv_groups_t& G, int64_t* out_runs,
                           double* out_values, cudaStream_t st) {
    const uint16_t* ev = reinterpret_cast<const uint16_t*>(L.wa);
    const int64_t nchunks = cdiv(pl.nb, RD_CHUNK);
    const unsigned grid = (unsigned)(nchunks < 148 * 4 ? nchunks : 148 * 4);
    if (!emit)
        k_rd_dense<false><<<grid, RD_THREADS, 0, st>>>(ev, L.bstart, pl.nb, G, L.gF, L.gZ, L.tickets + 2, L.bcnt, L.bsum,
                                                      L.blast, L.glocal, nullptr, nullptr);
    else
        k_rd_dense<true><<<grid, RD_THREADS, 0, st>>>(ev, L.bstart, pl.nb, G, L.gF, L.gZ, L.tickets + 3, L.bcnt, L.bsum,
                                                     L.blast, L.glocal, out_runs, out_values);
    IV_LAUNCH_CHECK();
    return IV_OK;
}

int rle_dense_count(const int64_t* starts, const int64_t* ends, int64_t n, const iv_groups_t* groups, void* ws,
                    size_t ws_bytes, int64_t* out_run_off, cudaStream_t st) {
    const iv_groups_t& G = *groups;
    const RdPlan pl = rd_plan(n, G.total_span);
    IV_REQUIRE(pl.use, IV_ERR_INVALID, "rle_dense_count: plan does not apply");
    Bump b(ws, ws_bytes);
    RdLayout L;
    rd_layout(n, G.n_groups, b, L);
    IV_REQUIRE(b.fits(), IV_ERR_WORKSPACE, "iv_rle_count: workspace %zu < %zu", ws_bytes, b.off);
    const int64_t m = 2 * n;
    const int nseg = 1 << (RD_W1 + pl.w2);
    IV_CUDA(cudaMemsetAsync((char*)ws + L.zero_from, 0, L.zero_to - L.zero_from, st));
    IV_CUDA(cudaMemsetAsync(L.blast, 0xff, (size_t)pl.nb * 8, st));
    k_rd_init<<<(G.n_groups + 255) / 256, 256, 0, st>>>(L.gF, L.gZ, G.n_groups);
    IV_LAUNCH_CHECK();
    k_rd_prep<<<148 * 4, RD_THREADS, (size_t)nseg * 4, st>>>(starts, ends, n, G, nseg, L.joint, L.gF, L.gZ);
    IV_LAUNCH_CHECK();
    k_rd_scan<<<1, 1024, 0, st>>>(L.joint, nseg, RD_W1, L.start2, L.toff2, L.bstart + ((size_t)1 << pl.T));
    IV_LAUNCH_CHECK();
    const unsigned tiles_a = (unsigned)cdiv(m, RD_TILE);
    k_rd_pass_a<<<tiles_a, RD_THREADS, 0, st>>>(starts, ends, n, G, pl.w2, L.start2, L.curA, L.wa);
    IV_LAUNCH_CHECK();
    const unsigned tiles_b = tiles_a + (1u << pl.w2);  // >= the exact count in toff2[2^w2]; surplus blocks leave at once
    k_rd_pass_b<<<tiles_b, RD_THREADS, 0, st>>>(L.wa, pl.w2, L.start2, L.toff2, L.joint, L.curB, L.wb);
    IV_LAUNCH_CHECK();
    k_rd_pass_c<<<(unsigned)nseg, RD_THREADS, 0, st>>>(L.wb, L.joint, L.bstart, reinterpret_cast<uint16_t*>(L.wa));
    IV_LAUNCH_CHECK();
    int rc = rd_dense_launch(false, pl, L, G, nullptr, nullptr, st);
    if (rc != IV_OK) return rc;
    const int64_t ntiles = cdiv(pl.nb, RD_BSCAN);
    k_rd_bscan_tiles<<<(unsigned)ntiles, 256, 0, st>>>(L.bcnt, L.bsum, L.blast, pl.nb, L.tagg, ntiles);
    IV_LAUNCH_CHECK();
    k_rd_bscan_mid<<<1, 1024, 0, st>>>(L.tagg, ntiles, L.tagg + 3 * ntiles);
    IV_LAUNCH_CHECK();
    k_rd_bscan_apply<<<(unsigned)ntiles, 256, 0, st>>>(L.bcnt, L.bsum, L.blast, pl.nb, L.tagg, ntiles);
    IV_LAUNCH_CHECK();
    k_rd_run_off<<<(G.n_groups + 256) / 256, 256, 0, st>>>(G, L.bcnt, L.glocal, L.tagg + 3 * ntiles, out_run_off);
    IV_LAUNCH_CHECK();
    return IV_OK;
}

int rle_dense_emit(int64_t n, const iv_groups_t* groups, const void* ws, int64_t* out_runs, double* out_values,
                   cudaStream_t st) {
    const iv_groups_t& G = *groups;
    const RdPlan pl = rd_plan(n, G.total_span);
    IV_REQUIRE(pl.use, IV_ERR_INVALID, "rle_dense_emit: plan does not apply");
    Bump b((void*)ws, (size_t)-1);
    RdLayout L;
    rd_layout(n, G.n_groups, b, L);
    IV_CUDA(cudaMemsetAsync(L.tickets + 3, 0, 4, st));
    return rd_dense_launch(true, pl, L, G, out_runs, out_values, st);
}

}  // namespace iv
