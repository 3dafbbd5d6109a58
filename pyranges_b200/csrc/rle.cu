// rle.cu -- batched coverage run-length encoding over all groups at once.
//
// Replaces pyrle.methods.coverage (reference call site pyranges/methods/to_rle.py:18; semantics
// pinned by the doctests pyranges/pyranges_main.py:5800-5872): per (chromosome[, strand]) key the
// step function  level(p) = #{Start <= p} - #{End <= p}  as runs of equal value from coordinate 0 up
// to the largest End of the key; leading zero run kept, no trailing zero run, equal neighbours
// coalesced.
//
//   events  key = (flattened position << 1) | is_end        2 per interval
//   sort    LSD radix sort on the position bits only (bit 0 rides along: the order of +1/-1 at one
//           position is irrelevant because only the per-position sum matters)
//   level   L[j] = inclusive sum of (+1|-1) over the sorted events.  Every group sums to zero, so ONE
//           device-wide scan serves all groups.
//   pieces  the tail event of each distinct position p carries  A = level after p  and
//           B = level before p (= L just before the head of the equal-position run, carried along the
//           run; across tiles through a "last head" tile scan).
//   runs    a piece starts a run iff A != B (or it sits on coordinate 0) and it is not the last
//           position of its group; the first piece of a group at p > 0 also emits the leading zero run.
//   length  next run start of the same group (or the group's largest End) minus own start.
#include "common.cuh"
#include <stdlib.h>

namespace iv {

constexpr int RL_THREADS = 256;
constexpr int RL_ITEMS = 8;
constexpr int RL_TILE = RL_THREADS * RL_ITEMS;  // 2048 events per block
constexpr int64_t RL_NONE = INT64_MIN;

__global__ void __launch_bounds__(RL_THREADS)
k_rle_keys(const int64_t* __restrict__ s, const int64_t* __restrict__ e, int64_t n, iv_groups_t G,
           uint64_t* __restrict__ keys) {
    __shared__ int span[2];
    int64_t first = (int64_t)blockIdx.x * RL_THREADS;
    int64_t last = first + RL_THREADS - 1 < n - 1 ? first + RL_THREADS - 1 : n - 1;
    SegSpan sp = block_seg_span(G.row_off, G.n_groups, first, last, span);
    int64_t i = first + threadIdx.x;
    if (i >= n) return;
    int g = sp.of(G.row_off, i);
    int64_t off = __ldg(G.base + g) - __ldg(G.lo + g);
    ulonglong2 v;
    v.x = (uint64_t)(s[i] + off) << 1;
    v.y = ((uint64_t)(e[i] + off) << 1) | 1ull;
    reinterpret_cast<ulonglong2*>(keys)[i] = v;
}

// shared per-tile state of the three sweeps over the sorted events
struct RleTile {
    uint64_t k[RL_ITEMS];
    uint64_t kprev, knext;  // neighbours of the thread's slice (kprev invalid when j0 == 0)
    int64_t j0;
    int nv;  // valid items of this thread
};

__device__ __forceinline__ void rle_load(const uint64_t* __restrict__ keys, int64_t m, RleTile& T) {
    T.j0 = (int64_t)blockIdx.x * RL_TILE + (int64_t)threadIdx.x * RL_ITEMS;
    int64_t rem = m - T.j0;
    T.nv = rem >= RL_ITEMS ? RL_ITEMS : (rem > 0 ? (int)rem : 0);
    if (T.nv == RL_ITEMS) {
        const ulonglong2* p = reinterpret_cast<const ulonglong2*>(keys + T.j0);
#pragma unroll
        for (int j = 0; j < RL_ITEMS / 2; j++) {
            ulonglong2 v = __ldg(p + j);
            T.k[2 * j] = v.x;
            T.k[2 * j + 1] = v.y;
        }
    } else {
#pragma unroll
        for (int j = 0; j < RL_ITEMS; j++) T.k[j] = j < T.nv ? __ldg(keys + T.j0 + j) : 0;
    }
    T.kprev = (T.nv > 0 && T.j0 > 0) ? __ldg(keys + T.j0 - 1) : 0;
    T.knext = (T.nv > 0 && T.j0 + T.nv < m) ? __ldg(keys + T.j0 + T.nv) : 0;
}
__device__ __forceinline__ int64_t rle_delta(uint64_t k) { return 1 - 2 * (int64_t)(k & 1ull); }

// block-wide "value of the last thread <= me that has one" (exclusive: strictly before me).
// idx = my thread id if I have a value else -1.  sm: int[RL_THREADS/32].  returns thread id or -1.
__device__ __forceinline__ int block_last_before(int idx, int* sm) {
    int inc = warp_incl_max(idx);
    int w = threadIdx.x >> 5;
    if (lane_id() == 31) sm[w] = inc;
    __syncthreads();
    int ex = __shfl_up_sync(FULL, inc, 1);
    if (lane_id() == 0) ex = -1;
    int carry = -1;
    for (int k = 0; k < w; k++) carry = sm[k] > carry ? sm[k] : carry;
    __syncthreads();
    return ex > carry ? ex : carry;
}

// sweep 1: tile sums of the deltas and the (packed) level before the last head of the tile
__global__ void __launch_bounds__(RL_THREADS)
k_rle_tiles(const uint64_t* __restrict__ keys, int64_t m, const int64_t* __restrict__ ev_off, int n_groups,
            int64_t* __restrict__ tile_sum, int64_t* __restrict__ tile_head) {
    __shared__ int64_t sm[RL_THREADS / 32 + 1];
    __shared__ int span[2];
    __shared__ int last_thr;
    __shared__ int red[RL_THREADS / 32];
    RleTile T;
    rle_load(keys, m, T);
    const int64_t tb = (int64_t)blockIdx.x * RL_TILE;
    const int64_t tl = tb + RL_TILE - 1 < m - 1 ? tb + RL_TILE - 1 : m - 1;
    SegSpan sp = block_seg_span(ev_off, n_groups, tb, tl, span);
    int64_t tsum = 0;
    int64_t head_val = RL_NONE;  // (exclusive in-thread prefix at my last head) * 2 + gfirst
#pragma unroll
    for (int j = 0; j < RL_ITEMS; j++) {
        if (j < T.nv) {
            uint64_t prev = j ? T.k[j - 1] : T.kprev;
            bool head = (T.j0 + j == 0) || ((prev >> 1) != (T.k[j] >> 1));
            if (head) {
                int g = sp.of(ev_off, T.j0 + j);
                bool gfirst = (T.j0 + j) == __ldg(ev_off + g);
                head_val = tsum * 2 + (gfirst ? 1 : 0);
            }
            tsum += rle_delta(T.k[j]);
        }
    }
    int64_t tot;
    int64_t ex = block_excl_sum<int64_t, RL_THREADS>(tsum, sm, tot);
    // last thread with a head
    int idx = head_val != RL_NONE ? (int)threadIdx.x : -1;
    int wm = warp_max(idx);
    if (lane_id() == 0) red[threadIdx.x >> 5] = wm;
    __syncthreads();
    if (threadIdx.x == 0) {
        int b = -1;
        for (int k = 0; k < RL_THREADS / 32; k++) b = red[k] > b ? red[k] : b;
        last_thr = b;
        tile_sum[blockIdx.x] = tot;
        if (b < 0) tile_head[blockIdx.x] = RL_NONE;
    }
    __syncthreads();
    if ((int)threadIdx.x == last_thr) tile_head[blockIdx.x] = head_val + 2 * ex;
}

// single block: carry[t] = packed (level before, gfirst) of the run that is open at the start of tile t
__global__ void __launch_bounds__(1024)
k_rle_carry(const int64_t* __restrict__ tile_excl, const int64_t* __restrict__ tile_head, int64_t nt,
            int64_t* __restrict__ carry_out) {
    __shared__ long long sm[1024 / 32 + 1];
    long long carry = -1;
    for (int64_t base = 0; base < nt; base += 1024) {
        int64_t t = base + threadIdx.x;
        long long idx = (t < nt && tile_head[t] != RL_NONE) ? (long long)t : -1;
        long long tot;
        long long inc = block_incl_max<long long, 1024>(idx, sm, (long long)-1, tot);
        inc = inc > carry ? inc : carry;
        // exclusive: last head tile strictly before t
        __shared__ long long prev[1024];
        prev[threadIdx.x] = inc;
        __syncthreads();
        long long ex = threadIdx.x ? prev[threadIdx.x - 1] : carry;
        if (t < nt) carry_out[t] = ex >= 0 ? (2 * tile_excl[ex] + tile_head[ex]) : 0;
        carry = tot > carry ? tot : carry;
        __syncthreads();
    }
}

// sweeps 2 and 3 share the per-event classification
//   EMIT=false: flags[j] (bit0 real run, bit1 leading zero run) + tile_cnt
//   EMIT=true : run starts (group-relative) and values
template <bool EMIT>
__global__ void __launch_bounds__(RL_THREADS)
k_rle_sweep(const uint64_t* __restrict__ keys, int64_t m, iv_groups_t G, const int64_t* __restrict__ ev_off,
            const int64_t* __restrict__ tile_excl, const int64_t* __restrict__ tile_carry,
            uint8_t* __restrict__ flags, int64_t* __restrict__ tile_cnt, const int64_t* __restrict__ tile_off,
            int64_t* __restrict__ run_start, double* __restrict__ run_value) {
    __shared__ int64_t sm[RL_THREADS / 32 + 1];
    __shared__ int span[2];
    __shared__ int smi[RL_THREADS / 32];
    __shared__ int64_t hv[RL_THREADS];
    RleTile T;
    rle_load(keys, m, T);
    const int64_t tb = (int64_t)blockIdx.x * RL_TILE;
    const int64_t tl = tb + RL_TILE - 1 < m - 1 ? tb + RL_TILE - 1 : m - 1;
    SegSpan sp = block_seg_span(ev_off, G.n_groups, tb, tl, span);

    int64_t tsum = 0;
    if (!EMIT) {
#pragma unroll
        for (int j = 0; j < RL_ITEMS; j++) if (j < T.nv) tsum += rle_delta(T.k[j]);
    } else {
#pragma unroll
        for (int j = 0; j < RL_ITEMS; j++) if (j < T.nv) tsum += rle_delta(T.k[j]);
    }
    int64_t tot;
    const int64_t ex = block_excl_sum<int64_t, RL_THREADS>(tsum, sm, tot) + __ldg(tile_excl + blockIdx.x);

    uint8_t f[RL_ITEMS];
    int cnt = 0;
    if (!EMIT) {
        // packed (level before, gfirst) at my last head -> neighbours to the right
        int64_t run = ex, my_head = RL_NONE;
#pragma unroll
        for (int j = 0; j < RL_ITEMS; j++) {
            if (j < T.nv) {
                uint64_t prev = j ? T.k[j - 1] : T.kprev;
                bool head = (T.j0 + j == 0) || ((prev >> 1) != (T.k[j] >> 1));
                if (head) {
                    int g = sp.of(ev_off, T.j0 + j);
                    my_head = run * 2 + ((T.j0 + j) == __ldg(ev_off + g) ? 1 : 0);
                }
                run += rle_delta(T.k[j]);
            }
        }
        hv[threadIdx.x] = my_head;
        int lt = block_last_before(my_head != RL_NONE ? (int)threadIdx.x : -1, smi);
        int64_t open = lt >= 0 ? hv[lt] : __ldg(tile_carry + blockIdx.x);
        run = ex;
#pragma unroll
        for (int j = 0; j < RL_ITEMS; j++) {
            f[j] = 0;
            if (j < T.nv) {
                const int64_t jj = T.j0 + j;
                uint64_t prev = j ? T.k[j - 1] : T.kprev;
                uint64_t next = j + 1 < T.nv ? T.k[j + 1] : T.knext;
                const int g = sp.of(ev_off, jj);
                bool head = (jj == 0) || ((prev >> 1) != (T.k[j] >> 1));
                if (head) open = run * 2 + (jj == __ldg(ev_off + g) ? 1 : 0);
                run += rle_delta(T.k[j]);
                bool tail = (jj == m - 1) || ((next >> 1) != (T.k[j] >> 1));
                if (tail) {
                    const int64_t B = open >> 1;  // arithmetic shift: floor, gfirst in bit 0
                    const bool gfirst = open & 1;
                    const bool glast = (jj + 1) == __ldg(ev_off + g + 1);
                    const int64_t pos = (int64_t)(T.k[j] >> 1) - __ldg(G.base + g);
                    uint8_t ff = 0;
                    if (gfirst && pos > 0) ff |= 2;
                    if (!glast && (run != B || (gfirst && pos == 0))) ff |= 1;
                    f[j] = ff;
                    cnt += (ff & 1) + (ff >> 1);
                }
            }
        }
        if (T.nv == RL_ITEMS) {
            uint64_t packed = 0;
#pragma unroll
            for (int j = 0; j < RL_ITEMS; j++) packed |= (uint64_t)f[j] << (8 * j);
            *reinterpret_cast<uint64_t*>(flags + T.j0) = packed;
        } else {
#pragma unroll
            for (int j = 0; j < RL_ITEMS; j++) if (j < T.nv) flags[T.j0 + j] = f[j];
        }
        int64_t c64 = warp_sum((int64_t)cnt);
        __syncthreads();
        if (lane_id() == 0) sm[threadIdx.x >> 5] = c64;
        __syncthreads();
        if (threadIdx.x == 0) {
            int64_t t = 0;
            for (int k = 0; k < RL_THREADS / 32; k++) t += sm[k];
            tile_cnt[blockIdx.x] = t;
        }
    } else {
        if (T.nv == RL_ITEMS) {
            uint64_t packed = *reinterpret_cast<const uint64_t*>(flags + T.j0);
#pragma unroll
            for (int j = 0; j < RL_ITEMS; j++) f[j] = (uint8_t)(packed >> (8 * j));
        } else {
#pragma unroll
            for (int j = 0; j < RL_ITEMS; j++) f[j] = j < T.nv ? flags[T.j0 + j] : 0;
        }
#pragma unroll
        for (int j = 0; j < RL_ITEMS; j++) cnt += (f[j] & 1) + (f[j] >> 1);
        int64_t tot2;
        __syncthreads();
        int64_t r = block_excl_sum<int64_t, RL_THREADS>((int64_t)cnt, sm, tot2) + __ldg(tile_off + blockIdx.x);
        int64_t run = ex;
#pragma unroll
        for (int j = 0; j < RL_ITEMS; j++) {
            if (j < T.nv) {
                run += rle_delta(T.k[j]);
                if (f[j]) {
                    const int g = sp.of(ev_off, T.j0 + j);
                    const int64_t pos = (int64_t)(T.k[j] >> 1) - __ldg(G.base + g);
                    if (f[j] & 2) { run_start[r] = 0; run_value[r] = 0.0; r++; }
                    if (f[j] & 1) { run_start[r] = pos; run_value[r] = (double)run; r++; }
                }
            }
        }
    }
}

// out_run_off[g] = runs emitted before the first event of group g
__global__ void k_rle_group_offsets(const uint8_t* __restrict__ flags, int64_t m, const int64_t* __restrict__ ev_off,
                                    int n_groups, const int64_t* __restrict__ tile_off, int64_t* __restrict__ out_run_off) {
    int g = blockIdx.x;  // one warp per group, g in [0, n_groups]
    int64_t e0 = g < n_groups ? ev_off[g] : m;
    int64_t t = e0 / RL_TILE;
    int64_t c = 0;
    for (int64_t j = t * RL_TILE + threadIdx.x; j < e0; j += 32) {
        uint8_t f = flags[j];
        c += (f & 1) + (f >> 1);
    }
    c = warp_sum(c);
    if (threadIdx.x == 0) out_run_off[g] = tile_off[t] + c;
}

// run length = next start in the same group (or the group's largest End) - own start
__global__ void __launch_bounds__(RL_THREADS)
k_rle_finish(const int64_t* __restrict__ run_start, int64_t n_runs, const int64_t* __restrict__ run_off,
             const uint64_t* __restrict__ keys, const int64_t* __restrict__ ev_off, iv_groups_t G,
             int64_t* __restrict__ out_runs) {
    __shared__ int span[2];
    int64_t first = (int64_t)blockIdx.x * RL_THREADS;
    int64_t last = first + RL_THREADS - 1 < n_runs - 1 ? first + RL_THREADS - 1 : n_runs - 1;
    SegSpan sp = block_seg_span(run_off, G.n_groups, first, last, span);
    int64_t r = first + threadIdx.x;
    if (r >= n_runs) return;
    int g = sp.of(run_off, r);
    // skip empty groups that share the same offset: find_seg returns the LAST g with run_off[g] <= r
    int64_t s = run_start[r];
    int64_t nxt;
    if (r + 1 < __ldg(run_off + g + 1)) nxt = run_start[r + 1];
    else nxt = (int64_t)(__ldg(keys + __ldg(ev_off + g + 1) - 1) >> 1) - __ldg(G.base + g);
    out_runs[r] = nxt - s;
}

__global__ void k_rle_ev_off(const int64_t* __restrict__ row_off, int n, int64_t* __restrict__ ev_off) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i <= n) ev_off[i] = 2 * row_off[i];
}

struct RleLayout {
    uint64_t *k0, *k1;
    uint8_t* flags;
    int64_t *tile_sum, *tile_head, *tile_carry, *tile_cnt, *ev_off;
    void* sort_temp;
};

static void plan_rle(int64_t n, int n_groups, Bump& b, RleLayout& L) {
    size_t m = (size_t)(n > 0 ? 2 * n : 2);
    size_t nt = (size_t)cdiv((int64_t)m, RL_TILE) + 2;
    L.k0 = b.take<uint64_t>(m);
    L.k1 = b.take<uint64_t>(m);
    L.flags = b.take<uint8_t>(m + 8);
    L.tile_sum = b.take<int64_t>(nt);
    L.tile_head = b.take<int64_t>(nt);
    L.tile_carry = b.take<int64_t>(nt);
    L.tile_cnt = b.take<int64_t>(nt);
    L.ev_off = b.take<int64_t>((size_t)n_groups + 2);
    L.sort_temp = b.take<char>(radix_sort_temp_bytes((int64_t)m));
}

// the dense path (rle_dense.cu): whole-genome read sets
size_t rle_dense_workspace_bytes(int64_t n, int n_groups);
bool rle_dense_applies(int64_t n, int64_t total_span);
int rle_dense_count(const int64_t* starts, const int64_t* ends, int64_t n, const iv_groups_t* groups, void* ws,
                    size_t ws_bytes, int64_t* out_run_off, cudaStream_t st);
int rle_dense_emit(int64_t n, const iv_groups_t* groups, const void* ws, int64_t* out_runs, double* out_values,
                   cudaStream_t st);
static bool use_dense(int64_t n, const iv_groups_t& G) {
    static const bool off = getenv("IV_RLE_NO_DENSE") != nullptr;  // A/B timing and tests of the sort-based path
    return !off && rle_dense_applies(n, G.total_span);
}

}  // namespace iv

using namespace iv;

extern "C" size_t iv_rle_workspace_bytes(int64_t n, int32_t n_groups) {
    Bump b(nullptr, 0);
    RleLayout L;
    plan_rle(n, n_groups > 0 ? n_groups : 0, b, L);
    const size_t dense = rle_dense_workspace_bytes(n, n_groups > 0 ? n_groups : 0);
    return (b.off > dense ? b.off : dense) + 256;
}

extern "C" int iv_rle_count(const int64_t* starts, const int64_t* ends, int64_t n, const iv_groups_t* groups, void* ws,
                            size_t ws_bytes, int64_t* out_run_off, void* stream) {
    cudaStream_t st = (cudaStream_t)stream;
    IV_REQUIRE(groups && n >= 0 && out_run_off, IV_ERR_INVALID, "iv_rle_count: bad arguments");
    IV_REQUIRE(n < ((int64_t)1 << 30), IV_ERR_RANGE, "iv_rle_count: more than 2^30-1 rows");
    const iv_groups_t& G = *groups;
    int bits = bits_for((uint64_t)(G.total_span > 0 ? G.total_span : 1));
    IV_REQUIRE(bits <= 62, IV_ERR_RANGE, "iv_rle_count: flattened span too large");
    if (use_dense(n, G)) return rle_dense_count(starts, ends, n, groups, ws, ws_bytes, out_run_off, st);
    Bump b(ws, ws_bytes);
    RleLayout L;
    plan_rle(n, G.n_groups, b, L);
    IV_REQUIRE(b.fits(), IV_ERR_WORKSPACE, "iv_rle_count: workspace %zu < %zu", ws_bytes, b.off);
    if (n == 0) {
        IV_CUDA(cudaMemsetAsync(out_run_off, 0, ((size_t)G.n_groups + 1) * 8, st));
        return IV_OK;
    }
    const int64_t m = 2 * n;
    k_rle_ev_off<<<(G.n_groups + 256) / 256, 256, 0, st>>>(G.row_off, G.n_groups, L.ev_off);
    IV_LAUNCH_CHECK();
    k_rle_keys<<<(unsigned)cdiv(n, RL_THREADS), RL_THREADS, 0, st>>>(starts, ends, n, G, L.k0);
    IV_LAUNCH_CHECK();
    int res = 0;
    int rc = radix_sort_pingpong(L.k0, L.k1, nullptr, nullptr, 0, m, 1, bits + 1, L.sort_temp,
                                 radix_sort_temp_bytes(m), st, &res);
    if (rc != IV_OK) return rc;
    const uint64_t* keys = res ? L.k1 : L.k0;
    const unsigned nt = (unsigned)cdiv(m, RL_TILE);
    k_rle_tiles<<<nt, RL_THREADS, 0, st>>>(keys, m, L.ev_off, G.n_groups, L.tile_sum, L.tile_head);
    IV_LAUNCH_CHECK();
    rc = scan_exclusive_i64_inplace(L.tile_sum, nt, nullptr, st);
    if (rc != IV_OK) return rc;
    k_rle_carry<<<1, 1024, 0, st>>>(L.tile_sum, L.tile_head, nt, L.tile_carry);
    IV_LAUNCH_CHECK();
    k_rle_sweep<false><<<nt, RL_THREADS, 0, st>>>(keys, m, G, L.ev_off, L.tile_sum, L.tile_carry, L.flags, L.tile_cnt,
                                                  nullptr, nullptr, nullptr);
    IV_LAUNCH_CHECK();
    rc = scan_exclusive_i64_inplace(L.tile_cnt, nt, L.tile_cnt + nt, st);  // tile_cnt[nt] = total
    if (rc != IV_OK) return rc;
    k_rle_group_offsets<<<G.n_groups + 1, 32, 0, st>>>(L.flags, m, L.ev_off, G.n_groups, L.tile_cnt, out_run_off);
    IV_LAUNCH_CHECK();
    return IV_OK;
}

extern "C" int iv_rle_emit(int64_t n, const iv_groups_t* groups, const void* ws, int64_t n_runs,
                           const int64_t* run_off, int64_t* out_runs, double* out_values, void* stream) {
    cudaStream_t st = (cudaStream_t)stream;
    IV_REQUIRE(groups && n >= 0 && n_runs >= 0, IV_ERR_INVALID, "iv_rle_emit: bad arguments");
    if (n == 0 || n_runs == 0) return IV_OK;
    IV_REQUIRE(ws && run_off && out_runs && out_values, IV_ERR_INVALID, "iv_rle_emit: null argument");
    IV_REQUIRE(n_runs <= 2 * n, IV_ERR_INVALID, "iv_rle_emit: n_runs > 2n");
    const iv_groups_t& G = *groups;
    if (use_dense(n, G)) return rle_dense_emit(n, groups, ws, out_runs, out_values, st);
    Bump b((void*)ws, (size_t)-1);
    RleLayout L;
    plan_rle(n, G.n_groups, b, L);
    const int64_t m = 2 * n;
    int bits = bits_for((uint64_t)(G.total_span > 0 ? G.total_span : 1));
    int res = radix_sort_passes(m, bits) & 1;  // the buffer iv_rle*_count's sort ended in
    const uint64_t* keys = res ? L.k1 : L.k0;
    int64_t* run_start = (int64_t*)(res ? L.k0 : L.k1);  // the idle sort buffer holds the run starts
    const unsigned nt = (unsigned)cdiv(m, RL_TILE);
    k_rle_sweep<true><<<nt, RL_THREADS, 0, st>>>(keys, m, G, L.ev_off, L.tile_sum, nullptr, L.flags, nullptr, L.tile_cnt,
                                                 run_start, out_values);
    IV_LAUNCH_CHECK();
    k_rle_finish<<<(unsigned)cdiv(n_runs, RL_THREADS), RL_THREADS, 0, st>>>(run_start, n_runs, run_off, keys, L.ev_off, G,
                                                                             out_runs);
    IV_LAUNCH_CHECK();
    return IV_OK;
}

// =====================================================================================================================
// Weighted coverage: to_rle(value_col=...) -- pyrle.methods.coverage with float64 weights (pyranges/methods/to_rle.py:18,
// doctest pyranges/pyranges_main.py:5838-5856).  The level is a float64 sum, so a parallel scan can differ from
// pyrle's sequential cumulative sum in the last bits (exact for integer-valued weights); the sums restart at every
// group so no rounding residue leaks from one key into the next.  Straightforward multi-pass formulation (this is
// not the hot configuration): sort events with their row, signed weights, segmented inclusive sum of the weights,
// inclusive max of the head indices, flags, scan, emit.
// =====================================================================================================================
namespace iv {

struct SegD {  // segmented-sum element: flag = a group starts here
    double v;
    int flag;
};
struct SegSumOp {
    __device__ static SegD identity() { return SegD{0.0, 0}; }
    __device__ SegD operator()(const SegD& a, const SegD& b) const { return SegD{b.flag ? b.v : a.v + b.v, a.flag | b.flag}; }
};
struct MaxI64Op {
    __device__ static long long identity() { return -1; }
    __device__ long long operator()(long long a, long long b) const { return a > b ? a : b; }
};
__device__ __forceinline__ SegD shfl_up_any(SegD x, int o) {
    SegD r;
    r.v = __shfl_up_sync(FULL, x.v, o);
    r.flag = __shfl_up_sync(FULL, x.flag, o);
    return r;
}
__device__ __forceinline__ long long shfl_up_any(long long x, int o) { return __shfl_up_sync(FULL, x, o); }

constexpr int GS_THREADS = 256;
constexpr int GS_ITEMS = 8;
constexpr int GS_TILE = GS_THREADS * GS_ITEMS;

// in-block ordered inclusive scan of GS_TILE elements (blocked: thread t owns items [8t, 8t+8)); returns the
// exclusive prefix of the thread's slice and the block total.  sm: T[GS_THREADS/32]
template <typename T, typename Op>
__device__ __forceinline__ T block_scan_excl(T mine, T* sm, T& total, Op op) {
    T inc = mine;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        T t = shfl_up_any(inc, o);
        if (lane_id() >= o) inc = op(t, inc);
    }
    const int w = threadIdx.x >> 5;
    if (lane_id() == 31) sm[w] = inc;
    __syncthreads();
    T carry = Op::identity();
    for (int k = 0; k < w; k++) carry = op(carry, sm[k]);
    T tot = Op::identity();
    for (int k = 0; k < GS_THREADS / 32; k++) tot = op(tot, sm[k]);
    total = tot;
    T ex = shfl_up_any(inc, 1);
    if (lane_id() == 0) ex = Op::identity();
    __syncthreads();
    return op(carry, ex);
}

// pass 1: tile totals; pass 3 (APPLY): in-place inclusive scan with the scanned tile totals as carry
template <typename T, typename Op, bool APPLY>
__global__ void __launch_bounds__(GS_THREADS) k_gscan(T* __restrict__ data, int64_t n, T* __restrict__ partials) {
    __shared__ T sm[GS_THREADS / 32];
    Op op;
    const int64_t base = (int64_t)blockIdx.x * GS_TILE + (int64_t)threadIdx.x * GS_ITEMS;
    T v[GS_ITEMS];
    T mine = Op::identity();
#pragma unroll
    for (int j = 0; j < GS_ITEMS; j++) {
        v[j] = base + j < n ? data[base + j] : Op::identity();
        mine = op(mine, v[j]);
    }
    T tot;
    T ex = block_scan_excl<T, Op>(mine, sm, tot, op);
    if (!APPLY) {
        if (threadIdx.x == 0) partials[blockIdx.x] = tot;
    } else {
        T run = op(partials[blockIdx.x], ex);
#pragma unroll
        for (int j = 0; j < GS_ITEMS; j++) {
            run = op(run, v[j]);
            if (base + j < n) data[base + j] = run;
        }
    }
}
// pass 2: single block, exclusive scan of the tile totals in place
template <typename T, typename Op>
__global__ void __launch_bounds__(GS_THREADS) k_gscan_partials(T* __restrict__ partials, int64_t nt) {
    __shared__ T sm[GS_THREADS / 32];
    Op op;
    T carry = Op::identity();
    for (int64_t b0 = 0; b0 < nt; b0 += GS_TILE) {
        const int64_t base = b0 + (int64_t)threadIdx.x * GS_ITEMS;
        T v[GS_ITEMS];
        T mine = Op::identity();
#pragma unroll
        for (int j = 0; j < GS_ITEMS; j++) {
            v[j] = base + j < nt ? partials[base + j] : Op::identity();
            mine = op(mine, v[j]);
        }
        T tot;
        T ex = block_scan_excl<T, Op>(mine, sm, tot, op);
        T run = op(carry, ex);
#pragma unroll
        for (int j = 0; j < GS_ITEMS; j++) {
            if (base + j < nt) partials[base + j] = run;
            run = op(run, v[j]);
        }
        carry = op(carry, tot);
    }
}
template <typename T, typename Op>
static int device_scan_inclusive(T* data, int64_t n, T* partials, cudaStream_t st) {
    const unsigned nt = (unsigned)cdiv(n, GS_TILE);
    k_gscan<T, Op, false><<<nt, GS_THREADS, 0, st>>>(data, n, partials);
    IV_LAUNCH_CHECK();
    k_gscan_partials<T, Op><<<1, GS_THREADS, 0, st>>>(partials, nt);
    IV_LAUNCH_CHECK();
    k_gscan<T, Op, true><<<nt, GS_THREADS, 0, st>>>(data, n, partials);
    IV_LAUNCH_CHECK();
    return IV_OK;
}

__global__ void __launch_bounds__(RL_THREADS)
k_rlew_keys(const int64_t* __restrict__ s, const int64_t* __restrict__ e, int64_t n, iv_groups_t G,
            uint64_t* __restrict__ keys, uint32_t* __restrict__ rows) {
    __shared__ int span[2];
    int64_t first = (int64_t)blockIdx.x * RL_THREADS;
    int64_t last = first + RL_THREADS - 1 < n - 1 ? first + RL_THREADS - 1 : n - 1;
    SegSpan sp = block_seg_span(G.row_off, G.n_groups, first, last, span);
    int64_t i = first + threadIdx.x;
    if (i >= n) return;
    int g = sp.of(G.row_off, i);
    int64_t off = __ldg(G.base + g) - __ldg(G.lo + g);
    keys[2 * i] = (uint64_t)(s[i] + off) << 1;
    keys[2 * i + 1] = ((uint64_t)(e[i] + off) << 1) | 1ull;
    rows[2 * i] = rows[2 * i + 1] = (uint32_t)i;
}

// signed weight of every sorted event (+ group-start flag) and the head index of its equal-position run
__global__ void __launch_bounds__(RL_THREADS)
k_rlew_delta(const uint64_t* __restrict__ keys, const uint32_t* __restrict__ rows, const double* __restrict__ w,
             int64_t m, const int64_t* __restrict__ ev_off, int n_groups, SegD* __restrict__ level,
             long long* __restrict__ head) {
    __shared__ int span[2];
    int64_t first = (int64_t)blockIdx.x * RL_THREADS;
    int64_t last = first + RL_THREADS - 1 < m - 1 ? first + RL_THREADS - 1 : m - 1;
    SegSpan sp = block_seg_span(ev_off, n_groups, first, last, span);
    int64_t j = first + threadIdx.x;
    if (j >= m) return;
    const uint64_t k = keys[j];
    const double x = __ldg(w + __ldg(rows + j));
    const int g = sp.of(ev_off, j);
    level[j] = SegD{(k & 1ull) ? -x : x, j == __ldg(ev_off + g) ? 1 : 0};
    head[j] = (j == 0 || (keys[j - 1] >> 1) != (k >> 1)) ? j : -1;
}

// classification of every tail event (same rules as k_rle_sweep); EMIT writes run starts / values
template <bool EMIT>
__global__ void __launch_bounds__(RL_THREADS)
k_rlew_sweep(const uint64_t* __restrict__ keys, int64_t m, iv_groups_t G, const int64_t* __restrict__ ev_off,
             const SegD* __restrict__ level, const long long* __restrict__ head, uint8_t* __restrict__ flags,
             int64_t* __restrict__ tile_cnt, const int64_t* __restrict__ tile_off, int64_t* __restrict__ run_start,
             double* __restrict__ run_value) {
    __shared__ int span[2];
    __shared__ int64_t sm[RL_THREADS / 32 + 1];
    const int64_t first = (int64_t)blockIdx.x * RL_THREADS;
    const int64_t last = first + RL_THREADS - 1 < m - 1 ? first + RL_THREADS - 1 : m - 1;
    SegSpan sp = block_seg_span(ev_off, G.n_groups, first, last, span);
    const int64_t j = first + threadIdx.x;
    uint8_t ff = 0;
    double A = 0.0;
    int64_t pos = 0;
    if (j < m) {
        const uint64_t k = keys[j];
        const bool tail = (j == m - 1) || ((keys[j + 1] >> 1) != (k >> 1));
        if (tail) {
            const int g = sp.of(ev_off, j);
            const long long h = head[j];
            const int64_t g0 = __ldg(ev_off + g);
            const bool gfirst = h == g0;
            const bool glast = (j + 1) == __ldg(ev_off + g + 1);
            A = level[j].v;
            const double B = gfirst ? 0.0 : level[h - 1].v;
            pos = (int64_t)(k >> 1) - __ldg(G.base + g);
            if (gfirst && pos > 0) ff |= 2;
            if (!glast && (A != B || (gfirst && pos == 0))) ff |= 1;
        }
    }
    const int64_t cnt = (ff & 1) + (ff >> 1);
    if (!EMIT) {
        if (j < m) flags[j] = ff;
        int64_t tot;
        block_excl_sum<int64_t, RL_THREADS>(cnt, sm, tot);
        if (threadIdx.x == 0) tile_cnt[blockIdx.x] = tot;
    } else {
        int64_t tot;
        int64_t r = block_excl_sum<int64_t, RL_THREADS>(cnt, sm, tot) + __ldg(tile_off + blockIdx.x);
        if (ff & 2) { run_start[r] = 0; run_value[r] = 0.0; r++; }
        if (ff & 1) { run_start[r] = pos; run_value[r] = A; }
    }
}

// run offsets of the groups from per-256-event tile offsets + flags (weighted layout)
__global__ void k_rlew_group_offsets(const uint8_t* __restrict__ flags, int64_t m, const int64_t* __restrict__ ev_off,
                                     int n_groups, const int64_t* __restrict__ tile_off, int64_t* __restrict__ out_run_off) {
    int g = blockIdx.x;
    int64_t e0 = g < n_groups ? ev_off[g] : m;
    int64_t t = e0 / RL_THREADS;
    int64_t c = 0;
    for (int64_t j = t * RL_THREADS + threadIdx.x; j < e0; j += 32) {
        uint8_t f = flags[j];
        c += (f & 1) + (f >> 1);
    }
    c = warp_sum(c);
    if (threadIdx.x == 0) out_run_off[g] = tile_off[t] + c;
}

struct RlewLayout {
    uint64_t *k0, *k1;
    uint32_t *r0, *r1;
    SegD *level, *lpart;
    long long *head, *hpart;
    uint8_t* flags;
    int64_t *tile_cnt, *ev_off;
    void* sort_temp;
};
static void plan_rlew(int64_t n, int n_groups, Bump& b, RlewLayout& L) {
    size_t m = (size_t)(n > 0 ? 2 * n : 2);
    size_t ngs = (size_t)cdiv((int64_t)m, GS_TILE) + 2, nt = (size_t)cdiv((int64_t)m, RL_THREADS) + 2;
    L.k0 = b.take<uint64_t>(m);
    L.k1 = b.take<uint64_t>(m);
    L.r0 = b.take<uint32_t>(m);
    L.r1 = b.take<uint32_t>(m);
    L.level = b.take<SegD>(m);
    L.lpart = b.take<SegD>(ngs);
    L.head = b.take<long long>(m);
    L.hpart = b.take<long long>(ngs);
    L.flags = b.take<uint8_t>(m + 8);
    L.tile_cnt = b.take<int64_t>(nt);
    L.ev_off = b.take<int64_t>((size_t)n_groups + 2);
    L.sort_temp = b.take<char>(radix_sort_temp_bytes((int64_t)m));
}
}  // namespace iv

extern "C" size_t iv_rle_weighted_workspace_bytes(int64_t n, int32_t n_groups) {
    Bump b(nullptr, 0);
    RlewLayout L;
    plan_rlew(n, n_groups > 0 ? n_groups : 0, b, L);
    return b.off + 256;
}

extern "C" int iv_rle_weighted_count(const int64_t* starts, const int64_t* ends, const double* weights, int64_t n,
                                     const iv_groups_t* groups, void* ws, size_t ws_bytes, int64_t* out_run_off,
                                     void* stream) {
    cudaStream_t st = (cudaStream_t)stream;
    IV_REQUIRE(groups && n >= 0 && out_run_off && (weights || n == 0), IV_ERR_INVALID, "iv_rle_weighted_count: bad arguments");
    IV_REQUIRE(n < ((int64_t)1 << 30), IV_ERR_RANGE, "iv_rle_weighted_count: more than 2^30-1 rows");
    const iv_groups_t& G = *groups;
    int bits = bits_for((uint64_t)(G.total_span > 0 ? G.total_span : 1));
    IV_REQUIRE(bits <= 62, IV_ERR_RANGE, "iv_rle_weighted_count: flattened span too large");
    Bump b(ws, ws_bytes);
    RlewLayout L;
    plan_rlew(n, G.n_groups, b, L);
    IV_REQUIRE(b.fits(), IV_ERR_WORKSPACE, "iv_rle_weighted_count: workspace %zu < %zu", ws_bytes, b.off);
    if (n == 0) {
        IV_CUDA(cudaMemsetAsync(out_run_off, 0, ((size_t)G.n_groups + 1) * 8, st));
        return IV_OK;
    }
    const int64_t m = 2 * n;
    k_rle_ev_off<<<(G.n_groups + 256) / 256, 256, 0, st>>>(G.row_off, G.n_groups, L.ev_off);
    IV_LAUNCH_CHECK();
    k_rlew_keys<<<(unsigned)cdiv(n, RL_THREADS), RL_THREADS, 0, st>>>(starts, ends, n, G, L.k0, L.r0);
    IV_LAUNCH_CHECK();
    int res = 0;
    int rc = radix_sort_pingpong(L.k0, L.k1, L.r0, L.r1, 4, m, 1, bits + 1, L.sort_temp, radix_sort_temp_bytes(m), st, &res);
    if (rc != IV_OK) return rc;
    const uint64_t* keys = res ? L.k1 : L.k0;
    const uint32_t* rows = res ? L.r1 : L.r0;
    const unsigned nb = (unsigned)cdiv(m, RL_THREADS);
    k_rlew_delta<<<nb, RL_THREADS, 0, st>>>(keys, rows, weights, m, L.ev_off, G.n_groups, L.level, L.head);
    IV_LAUNCH_CHECK();
    rc = device_scan_inclusive<SegD, SegSumOp>(L.level, m, L.lpart, st);
    if (rc != IV_OK) return rc;
    rc = device_scan_inclusive<long long, MaxI64Op>(L.head, m, L.hpart, st);
    if (rc != IV_OK) return rc;
    k_rlew_sweep<false><<<nb, RL_THREADS, 0, st>>>(keys, m, G, L.ev_off, L.level, L.head, L.flags, L.tile_cnt, nullptr,
                                                   nullptr, nullptr);
    IV_LAUNCH_CHECK();
    rc = scan_exclusive_i64_inplace(L.tile_cnt, nb, L.tile_cnt + nb, st);  // tile_cnt[nb] = total
    if (rc != IV_OK) return rc;
    k_rlew_group_offsets<<<G.n_groups + 1, 32, 0, st>>>(L.flags, m, L.ev_off, G.n_groups, L.tile_cnt, out_run_off);
    IV_LAUNCH_CHECK();
    return IV_OK;
}

extern "C" int iv_rle_weighted_emit(int64_t n, const iv_groups_t* groups, const void* ws, int64_t n_runs,
                                    const int64_t* run_off, int64_t* out_runs, double* out_values, void* stream) {
    cudaStream_t st = (cudaStream_t)stream;
    IV_REQUIRE(groups && n >= 0 && n_runs >= 0, IV_ERR_INVALID, "iv_rle_weighted_emit: bad arguments");
    if (n == 0 || n_runs == 0) return IV_OK;
    IV_REQUIRE(ws && run_off && out_runs && out_values, IV_ERR_INVALID, "iv_rle_weighted_emit: null argument");
    IV_REQUIRE(n_runs <= 2 * n, IV_ERR_INVALID, "iv_rle_weighted_emit: n_runs > 2n");
    const iv_groups_t& G = *groups;
    Bump b((void*)ws, (size_t)-1);
    RlewLayout L;
    plan_rlew(n, G.n_groups, b, L);
    const int64_t m = 2 * n;
    int bits = bits_for((uint64_t)(G.total_span > 0 ? G.total_span : 1));
    int res = radix_sort_passes(m, bits) & 1;  // the buffer iv_rle*_count's sort ended in
    const uint64_t* keys = res ? L.k1 : L.k0;
    int64_t* run_start = (int64_t*)(res ? L.k0 : L.k1);  // the idle sort buffer holds the run starts
    const unsigned nb = (unsigned)cdiv(m, RL_THREADS);
    k_rlew_sweep<true><<<nb, RL_THREADS, 0, st>>>(keys, m, G, L.ev_off, L.level, L.head, L.flags, nullptr, L.tile_cnt,
                                                  run_start, out_values);
    IV_LAUNCH_CHECK();
    k_rle_finish<<<(unsigned)cdiv(n_runs, RL_THREADS), RL_THREADS, 0, st>>>(run_start, n_runs, run_off, keys, L.ev_off, G,
                                                                             out_runs);
    IV_LAUNCH_CHECK();
    return IV_OK;
}
