// cluster.cu -- batched cluster / merge over all groups at once.
//
// Replaces df.sort_values("Start") + sorted_nearest.annotate_clusters / find_clusters
// (reference call sites pyranges/methods/cluster.py:11-15, merge.py:12-14) and the serial id-offset
// loop over keys (pyranges/pyranges_main.py:1210-1223).
//
//   key   = (flattened start << len_bits) | (end - start)       one 64-bit radix sort gives the
//           canonical (group, Start, End, row) order
//   flag  = start' - slack > max(end' of everything before)     new-cluster rule; the prefix max over ALL
//           earlier rows is equivalent to the running max of the open cluster (older clusters end
//           before start' - slack), and groups are separated by a gap > slack in the flattened axis
//   id    = inclusive sum of flags (1-based, global over groups in key order)
//   merge = one row per flag: start, max end (atomicMax per member), member count
#include "common.cuh"

namespace iv {

constexpr int CL_THREADS = 256;
constexpr int CL_ITEMS = 8;
constexpr int CL_TILE = CL_THREADS * CL_ITEMS;  // 2048 rows per block

__global__ void __launch_bounds__(CL_THREADS)
k_cluster_keys(const int64_t* __restrict__ s, const int64_t* __restrict__ e, int64_t n, iv_groups_t G, int lbits,
               uint64_t* __restrict__ keys, uint32_t* __restrict__ rows, int* __restrict__ unsorted) {
    __shared__ int span[2];
    int64_t first = (int64_t)blockIdx.x * CL_THREADS;
    int64_t last = first + CL_THREADS - 1 < n - 1 ? first + CL_THREADS - 1 : n - 1;
    SegSpan sp = block_seg_span(G.row_off, G.n_groups, first, last, span);
    int64_t i = first + threadIdx.x;
    if (i >= n) return;
    int g = sp.of(G.row_off, i);
    int64_t a = s[i], b = e[i];
    uint64_t len = b > a ? (uint64_t)(b - a) : 0;
    const uint64_t key = ((uint64_t)(a - __ldg(G.lo + g) + __ldg(G.base + g)) << lbits) | len;
    keys[i] = key;
    rows[i] = (uint32_t)i;
    // already in (group, Start, End) order?  then the stable sort is the identity and is skipped
    if (i > 0) {
        int64_t pa = s[i - 1], pb = e[i - 1];
        // (the row before the block's first one may belong to the previous group)
        int pg = i - 1 >= first ? sp.of(G.row_off, i - 1) : find_seg(G.row_off, G.n_groups, i - 1);
        uint64_t plen = pb > pa ? (uint64_t)(pb - pa) : 0;
        uint64_t pkey = ((uint64_t)(pa - __ldg(G.lo + pg) + __ldg(G.base + pg)) << lbits) | plen;
        if (pkey > key && *unsorted == 0) *unsorted = 1;
    }
}

// ---- wide case: (flattened start, length) does not fit one 64-bit key.  Two stable sorts instead -- by length,
// then by flattened start -- and the flattened ends kept in an array of their own.
__global__ void __launch_bounds__(CL_THREADS)
k_cluster_len_keys(const int64_t* __restrict__ s, const int64_t* __restrict__ e, int64_t n, uint64_t* __restrict__ keys,
                   uint32_t* __restrict__ rows) {
    int64_t i = (int64_t)blockIdx.x * CL_THREADS + threadIdx.x;
    if (i >= n) return;
    int64_t a = s[i], b = e[i];
    keys[i] = b > a ? (uint64_t)(b - a) : 0;
    rows[i] = (uint32_t)i;
}
// ENDS=false: keys[p] = flattened start of rows[p];  ENDS=true: ends[p] = keys[p] + length of rows[p]
template <bool ENDS>
__global__ void __launch_bounds__(CL_THREADS)
k_cluster_gather(const int64_t* __restrict__ s, const int64_t* __restrict__ e, int64_t n, iv_groups_t G,
                 const uint32_t* __restrict__ rows, uint64_t* __restrict__ keys, uint64_t* __restrict__ ends) {
    int64_t p = (int64_t)blockIdx.x * CL_THREADS + threadIdx.x;
    if (p >= n) return;
    const int64_t r = rows[p];
    const int64_t a = s[r], b = e[r];
    if (!ENDS) {
        const int g = find_seg(G.row_off, G.n_groups, r);
        keys[p] = (uint64_t)(a - __ldg(G.lo + g) + __ldg(G.base + g));
    } else {
        ends[p] = keys[p] + (b > a ? (uint64_t)(b - a) : 0);
    }
}

// single block exclusive max scan (identity INT64_MIN)
__global__ void __launch_bounds__(1024) k_scan_small_max(int64_t* __restrict__ data, int64_t m) {
    __shared__ int64_t sm[1024 / 32 + 1];
    __shared__ int64_t prev[1024];
    int64_t carry = INT64_MIN;
    for (int64_t base = 0; base < m; base += 1024) {
        int64_t i = base + threadIdx.x;
        int64_t v = i < m ? data[i] : INT64_MIN;
        int64_t tot;
        int64_t inc = block_incl_max<int64_t, 1024>(v, sm, (int64_t)INT64_MIN, tot);
        prev[threadIdx.x] = inc;
        __syncthreads();
        int64_t ex = threadIdx.x ? prev[threadIdx.x - 1] : INT64_MIN;
        ex = ex > carry ? ex : carry;
        if (i < m) data[i] = ex;
        carry = tot > carry ? tot : carry;
        __syncthreads();
    }
}

__global__ void __launch_bounds__(CL_THREADS)
k_cl_tilemax(const uint64_t* __restrict__ keys, const uint64_t* __restrict__ ends_opt, int64_t n, int lbits,
             int64_t* __restrict__ tile_max) {
    __shared__ int64_t red[CL_THREADS / 32];
    const uint64_t lmask = lbits ? ((1ull << lbits) - 1) : 0;
    int64_t base = (int64_t)blockIdx.x * CL_TILE;
    int64_t m = INT64_MIN;
#pragma unroll
    for (int it = 0; it < CL_ITEMS; it++) {
        int64_t i = base + it * CL_THREADS + threadIdx.x;
        if (i < n) {
            uint64_t k = __ldg(keys + i);
            int64_t E = ends_opt ? (int64_t)__ldg(ends_opt + i) : (int64_t)((k >> lbits) + (k & lmask));
            m = E > m ? E : m;
        }
    }
    m = warp_max(m);
    if (lane_id() == 0) red[threadIdx.x >> 5] = m;
    __syncthreads();
    if (threadIdx.x == 0) {
        for (int k = 1; k < CL_THREADS / 32; k++) m = red[k] > m ? red[k] : m;
        tile_max[blockIdx.x] = m;
    }
}

// flags[p] = 1 iff sorted row p opens a cluster; tile_cnt = flags per tile
__global__ void __launch_bounds__(CL_THREADS)
k_cl_flags(const uint64_t* __restrict__ keys, const uint64_t* __restrict__ ends_opt, int64_t n, int lbits, int64_t slack,
           const int64_t* __restrict__ tile_max_excl, uint8_t* __restrict__ flags, int64_t* __restrict__ tile_cnt) {
    __shared__ int64_t sm[CL_THREADS / 32 + 1];
    __shared__ int64_t thr_inc[CL_THREADS];
    const uint64_t lmask = lbits ? ((1ull << lbits) - 1) : 0;
    const int64_t p0 = (int64_t)blockIdx.x * CL_TILE + (int64_t)threadIdx.x * CL_ITEMS;
    int64_t S[CL_ITEMS], E[CL_ITEMS];
    int64_t tmax = INT64_MIN;
#pragma unroll
    for (int j = 0; j < CL_ITEMS; j++) {
        if (p0 + j < n) {
            uint64_t k = __ldg(keys + p0 + j);
            S[j] = (int64_t)(k >> lbits);
            E[j] = ends_opt ? (int64_t)__ldg(ends_opt + p0 + j) : S[j] + (int64_t)(k & lmask);
            tmax = E[j] > tmax ? E[j] : tmax;
        } else { S[j] = 0; E[j] = INT64_MIN; }
    }
    int64_t tot;
    int64_t inc = block_incl_max<int64_t, CL_THREADS>(tmax, sm, (int64_t)INT64_MIN, tot);
    thr_inc[threadIdx.x] = inc;
    __syncthreads();
    int64_t run = threadIdx.x ? thr_inc[threadIdx.x - 1] : INT64_MIN;
    int64_t tb = __ldg(tile_max_excl + blockIdx.x);
    run = tb > run ? tb : run;
    int cnt = 0;
    uint8_t f[CL_ITEMS];
#pragma unroll
    for (int j = 0; j < CL_ITEMS; j++) {
        f[j] = 0;
        if (p0 + j < n) {
            f[j] = (p0 + j == 0) || (run == INT64_MIN) || (S[j] - slack > run);
            run = E[j] > run ? E[j] : run;
            cnt += f[j];
        }
    }
    if (p0 + CL_ITEMS <= n) {
        uint64_t packed = 0;
#pragma unroll
        for (int j = 0; j < CL_ITEMS; j++) packed |= (uint64_t)f[j] << (8 * j);
        *reinterpret_cast<uint64_t*>(flags + p0) = packed;
    } else {
#pragma unroll
        for (int j = 0; j < CL_ITEMS; j++) if (p0 + j < n) flags[p0 + j] = f[j];
    }
    int64_t c64 = warp_sum((int64_t)cnt);
    __syncthreads();
    if (lane_id() == 0) sm[threadIdx.x >> 5] = c64;
    __syncthreads();
    if (threadIdx.x == 0) {
        int64_t t = 0;
        for (int k = 0; k < CL_THREADS / 32; k++) t += sm[k];
        tile_cnt[blockIdx.x] = t;
    }
}

// MERGE=false: out_id / out_row.  MERGE=true: per-cluster start / group / first index / max end.
template <bool MERGE>
__global__ void __launch_bounds__(CL_THREADS)
k_cl_ids(const uint64_t* __restrict__ keys, const uint64_t* __restrict__ ends_opt, const uint32_t* __restrict__ rows,
         const uint8_t* __restrict__ flags, int64_t n, int lbits, iv_groups_t G, const int64_t* __restrict__ tile_off, int64_t* __restrict__ out_row,
         int64_t* __restrict__ out_id, int64_t* __restrict__ c_start, int32_t* __restrict__ c_group,
         int64_t* __restrict__ c_first, int64_t* __restrict__ c_end) {
    __shared__ int64_t sm[CL_THREADS / 32 + 1];
    const uint64_t lmask = lbits ? ((1ull << lbits) - 1) : 0;
    const int64_t p0 = (int64_t)blockIdx.x * CL_TILE + (int64_t)threadIdx.x * CL_ITEMS;
    uint8_t f[CL_ITEMS];
    int cnt = 0;
#pragma unroll
    for (int j = 0; j < CL_ITEMS; j++) {
        f[j] = (p0 + j < n) ? flags[p0 + j] : 0;
        cnt += f[j];
    }
    int64_t tot;
    const int64_t toff = __ldg(tile_off + blockIdx.x);
    int64_t id = block_excl_sum<int64_t, CL_THREADS>((int64_t)cnt, sm, tot) + toff;
    if (!MERGE) {
        // the ids leave through shared memory (12 bits each, relative to the tile's first id) so that every store
        // instruction of a warp writes 256 consecutive bytes; the rows are a widening copy in the same order
        __shared__ uint16_t rel[CL_TILE];
        uint32_t r = (uint32_t)(id - toff);
#pragma unroll
        for (int j = 0; j < CL_ITEMS; j++) {
            r += f[j];
            rel[threadIdx.x * CL_ITEMS + j] = (uint16_t)r;
        }
        __syncthreads();
        const int64_t tb = (int64_t)blockIdx.x * CL_TILE;
#pragma unroll
        for (int k = 0; k < CL_ITEMS; k++) {
            const int64_t p = tb + k * CL_THREADS + threadIdx.x;
            if (p < n) {
                __stcs(out_id + p, toff + (int64_t)rel[k * CL_THREADS + threadIdx.x]);
                __stcs(out_row + p, (int64_t)__ldg(rows + p));
            }
        }
        return;
    }
    // the max end of a cluster: one atomic per run of rows of the same cluster inside the thread's eight rows (a
    // cluster holds nine rows on average at C4), not one per row
    int64_t run_id = 0, run_max = INT64_MIN;
#pragma unroll
    for (int j = 0; j < CL_ITEMS; j++) {
        const int64_t p = p0 + j;
        if (p >= n) break;
        id += f[j];
        uint64_t k = __ldg(keys + p);
        int64_t S = (int64_t)(k >> lbits);
        int64_t E = ends_opt ? (int64_t)__ldg(ends_opt + p) : S + (int64_t)(k & lmask);
        if (f[j]) {
            // group of a flattened coordinate: largest g with base[g] <= S
            int g = find_seg(G.base, G.n_groups, S);
            c_start[id - 1] = S - __ldg(G.base + g) + __ldg(G.lo + g);
            c_group[id - 1] = g;
            c_first[id - 1] = p;
        }
        if (id != run_id) {
            if (run_id > 0) atomicMax((long long*)c_end + (run_id - 1), (long long)run_max);
            run_id = id;
            run_max = E;
        } else {
            run_max = E > run_max ? E : run_max;
        }
    }
    if (run_id > 0) atomicMax((long long*)c_end + (run_id - 1), (long long)run_max);
}

__global__ void k_merge_finish(int64_t n, int64_t nc, iv_groups_t G,
                               const int64_t* __restrict__ c_first, const int32_t* __restrict__ c_group,
                               int64_t* __restrict__ c_end, int64_t* __restrict__ c_count) {
    for (int64_t c = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; c < nc; c += (int64_t)gridDim.x * blockDim.x) {
        int64_t nxt = c + 1 < nc ? c_first[c + 1] : n;
        if (c_count) c_count[c] = nxt - c_first[c];
        int g = c_group[c];
        c_end[c] = c_end[c] - __ldg(G.base + g) + __ldg(G.lo + g);
    }
}

struct ClusterLayout {
    uint64_t *k0, *k1;
    uint32_t *r0, *r1;
    uint64_t* e_sorted;  // wide case only: flattened ends in sorted order
    uint8_t* flags;
    int64_t *tile_max, *tile_cnt, *c_first;
    int* unsorted;  // [0] device flag written by k_cluster_keys (its host copy decides about the sort),
                    // [1] which ping-pong buffer holds the sorted keys, [2] wide case -- for iv_merge_emit
    void* sort_temp;
};

static void plan_cluster(int64_t n, Bump& b, ClusterLayout& L) {
    size_t nn = (size_t)(n > 0 ? n : 1);
    size_t nt = (size_t)cdiv(nn, CL_TILE) + 1;
    L.k0 = b.take<uint64_t>(nn);
    L.k1 = b.take<uint64_t>(nn);
    L.r0 = b.take<uint32_t>(nn);
    L.r1 = b.take<uint32_t>(nn);
    L.e_sorted = b.take<uint64_t>(nn);
    L.flags = b.take<uint8_t>(nn + 8);
    L.tile_max = b.take<int64_t>(nt);
    L.tile_cnt = b.take<int64_t>(nt);
    L.c_first = b.take<int64_t>(nn);
    L.unsorted = b.take<int>(64);
    L.sort_temp = b.take<char>(radix_sort_temp_bytes(n));
}

// sort + flags + scan; leaves tile offsets in L.tile_cnt and the total in *out_n_clusters
static int cluster_phase1(const int64_t* starts, const int64_t* ends, int64_t n, const iv_groups_t* G, int64_t slack,
                          int len_bits, void* ws, size_t ws_bytes, ClusterLayout& L, int* res, bool* wide_out,
                          int64_t* out_n_clusters, cudaStream_t st, const char* who) {
    IV_REQUIRE(G && n >= 0 && len_bits >= 0, IV_ERR_INVALID, "%s: bad arguments", who);
    IV_REQUIRE(n < ((int64_t)1 << 31), IV_ERR_RANGE, "%s: more than 2^31-1 rows", who);
    const int sbits = bits_for((uint64_t)G->total_span);
    const int tbits = sbits + len_bits;
    const bool wide = tbits > 63;
    IV_REQUIRE(sbits <= 62 && len_bits <= 62, IV_ERR_RANGE, "%s: span (%d bits) / max length (%d bits) too large", who,
               sbits, len_bits);
    Bump b(ws, ws_bytes);
    plan_cluster(n, b, L);
    IV_REQUIRE(b.fits(), IV_ERR_WORKSPACE, "%s: workspace %zu < %zu", who, ws_bytes, b.off);
    *wide_out = wide;
    if (n == 0) {
        IV_CUDA(cudaMemsetAsync(out_n_clusters, 0, 8, st));
        *res = 0;
        return IV_OK;
    }
    const unsigned nb = (unsigned)cdiv(n, CL_THREADS);
    int rc = IV_OK;
    *res = 0;
    if (!wide) {
        IV_CUDA(cudaMemsetAsync(L.unsorted, 0, sizeof(int), st));
        k_cluster_keys<<<nb, CL_THREADS, 0, st>>>(starts, ends, n, *G, len_bits, L.k0, L.r0, L.unsorted);
        IV_LAUNCH_CHECK();
        // one small host read (the only synchronisation of the call): sorted input skips all radix passes
        int h_unsorted = 1;
        IV_CUDA(cudaMemcpyAsync(&h_unsorted, L.unsorted, sizeof(int), cudaMemcpyDeviceToHost, st));
        IV_CUDA(cudaStreamSynchronize(st));
        if (h_unsorted) {
            rc = radix_sort_pingpong(L.k0, L.k1, L.r0, L.r1, 4, n, 0, tbits, L.sort_temp, radix_sort_temp_bytes(n), st, res);
            if (rc != IV_OK) return rc;
        }
    } else {
        k_cluster_len_keys<<<nb, CL_THREADS, 0, st>>>(starts, ends, n, L.k0, L.r0);
        IV_LAUNCH_CHECK();
        int r1 = 0, r2 = 0;
        rc = radix_sort_pingpong(L.k0, L.k1, L.r0, L.r1, 4, n, 0, len_bits, L.sort_temp, radix_sort_temp_bytes(n), st, &r1);
        if (rc != IV_OK) return rc;
        uint64_t *ka = r1 ? L.k1 : L.k0, *kb = r1 ? L.k0 : L.k1;
        uint32_t *ra = r1 ? L.r1 : L.r0, *rb = r1 ? L.r0 : L.r1;
        k_cluster_gather<false><<<nb, CL_THREADS, 0, st>>>(starts, ends, n, *G, ra, ka, nullptr);
        IV_LAUNCH_CHECK();
        rc = radix_sort_pingpong(ka, kb, ra, rb, 4, n, 0, sbits, L.sort_temp, radix_sort_temp_bytes(n), st, &r2);
        if (rc != IV_OK) return rc;
        *res = r2 ? (r1 ? 0 : 1) : r1;  // index of the buffer (k0/r0 = 0, k1/r1 = 1) holding the result
        k_cluster_gather<true><<<nb, CL_THREADS, 0, st>>>(starts, ends, n, *G, *res ? L.r1 : L.r0, *res ? L.k1 : L.k0,
                                                          L.e_sorted);
        IV_LAUNCH_CHECK();
    }
    int h_state[2] = {*res, wide ? 1 : 0};
    IV_CUDA(cudaMemcpyAsync(L.unsorted + 1, h_state, sizeof(h_state), cudaMemcpyHostToDevice, st));
    const uint64_t* keys = *res ? L.k1 : L.k0;
    const uint64_t* eopt = wide ? L.e_sorted : nullptr;
    const int lb = wide ? 0 : len_bits;
    unsigned nt = (unsigned)cdiv(n, CL_TILE);
    k_cl_tilemax<<<nt, CL_THREADS, 0, st>>>(keys, eopt, n, lb, L.tile_max);
    IV_LAUNCH_CHECK();
    k_scan_small_max<<<1, 1024, 0, st>>>(L.tile_max, nt);
    IV_LAUNCH_CHECK();
    k_cl_flags<<<nt, CL_THREADS, 0, st>>>(keys, eopt, n, lb, slack, L.tile_max, L.flags, L.tile_cnt);
    IV_LAUNCH_CHECK();
    return scan_exclusive_i64_inplace(L.tile_cnt, nt, out_n_clusters, st);
}

}  // namespace iv

using namespace iv;

extern "C" size_t iv_cluster_workspace_bytes(int64_t n) {
    Bump b(nullptr, 0);
    ClusterLayout L;
    plan_cluster(n, b, L);
    return b.off + 256;
}

extern "C" int iv_cluster(const int64_t* starts, const int64_t* ends, int64_t n, const iv_groups_t* groups, int64_t slack,
                          int32_t len_bits, int64_t* out_row, int64_t* out_id, int64_t* out_n_clusters, void* ws,
                          size_t ws_bytes, void* stream) {
    cudaStream_t st = (cudaStream_t)stream;
    ClusterLayout L;
    int res = 0;
    bool wide = false;
    int rc = cluster_phase1(starts, ends, n, groups, slack, len_bits, ws, ws_bytes, L, &res, &wide, out_n_clusters, st,
                            "iv_cluster");
    if (rc != IV_OK || n == 0) return rc;
    unsigned nt = (unsigned)cdiv(n, CL_TILE);
    k_cl_ids<false><<<nt, CL_THREADS, 0, st>>>(res ? L.k1 : L.k0, wide ? L.e_sorted : nullptr, res ? L.r1 : L.r0, L.flags, n,
                                                wide ? 0 : len_bits, *groups, L.tile_cnt, out_row, out_id, nullptr, nullptr,
                                                nullptr, nullptr);
    IV_LAUNCH_CHECK();
    return IV_OK;
}

extern "C" int iv_merge_count(const int64_t* starts, const int64_t* ends, int64_t n, const iv_groups_t* groups,
                              int64_t slack, int32_t len_bits, int64_t* out_n_clusters, void* ws, size_t ws_bytes,
                              void* stream) {
    ClusterLayout L;
    int res = 0;
    bool wide = false;
    return cluster_phase1(starts, ends, n, groups, slack, len_bits, ws, ws_bytes, L, &res, &wide, out_n_clusters,
                          (cudaStream_t)stream, "iv_merge_count");
}

extern "C" int iv_merge_emit(int64_t n, int64_t n_clusters, const iv_groups_t* groups, int32_t len_bits,
                             const void* ws, int64_t* out_start, int64_t* out_end, int64_t* out_count,
                             int32_t* out_group, void* stream) {
    cudaStream_t st = (cudaStream_t)stream;
    IV_REQUIRE(groups && n >= 0 && n_clusters >= 0 && n_clusters <= n, IV_ERR_INVALID, "iv_merge_emit: bad arguments");
    if (n == 0 || n_clusters == 0) return IV_OK;
    IV_REQUIRE(ws && out_start && out_end && out_group, IV_ERR_INVALID, "iv_merge_emit: null argument");
    Bump b((void*)ws, (size_t)-1);
    ClusterLayout L;
    plan_cluster(n, b, L);
    int h_state[2] = {0, 0};  // sorted-keys buffer and wide flag: left by iv_merge_count in the workspace
    IV_CUDA(cudaMemcpyAsync(h_state, L.unsorted + 1, sizeof(h_state), cudaMemcpyDeviceToHost, st));
    IV_CUDA(cudaStreamSynchronize(st));
    const int res = h_state[0];
    const bool wide = h_state[1] != 0;
    unsigned nt = (unsigned)cdiv(n, CL_TILE);
    // out_end doubles as the atomicMax accumulator of the flattened cluster end: 0x80.. = very negative
    IV_CUDA(cudaMemsetAsync(out_end, 0x80, (size_t)n_clusters * 8, st));
    k_cl_ids<true><<<nt, CL_THREADS, 0, st>>>(res ? L.k1 : L.k0, wide ? L.e_sorted : nullptr, res ? L.r1 : L.r0, L.flags, n,
                                               wide ? 0 : len_bits, *groups, L.tile_cnt, nullptr, nullptr, out_start,
                                               out_group, L.c_first, out_end);
    IV_LAUNCH_CHECK();
    unsigned nb = (unsigned)cdiv(n_clusters, 256);
    if (nb > 65535u * 16u) nb = 65535u * 16u;
    k_merge_finish<<<nb, 256, 0, st>>>(n, n_clusters, *groups, L.c_first, out_group, out_end, out_count);
    IV_LAUNCH_CHECK();
    return IV_OK;
}
