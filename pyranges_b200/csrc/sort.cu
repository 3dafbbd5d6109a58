// sort.cu -- stable LSD radix sort of 64-bit keys (+ optional 4/8-byte payload) and the device scans
// it needs.  Replaces libc qsort inside NCLS construction and pandas sort_values in
// pyranges/methods/{cluster,merge,nearest}.py and the mergesort of pyrle's event list.
//
// Per 8-bit digit, three launches whatever the size: (1) per-tile digit histogram, digit-major; (2) one warp
// per digit row: exclusive prefix along the tiles + row total; (3) scatter: every block scans the 256 row
// totals itself, warp-level multi-split ranks (__match_any_sync), keys staged through shared memory in digit
// order so that global writes are coalesced runs.  Several sorts of equal length share the launches
// (blockIdx.y).  Small inputs use 1024-key tiles so that they still spread over the 148 SMs.
// No inter-block dependencies (no spin waits), so it cannot hang.
#include "common.cuh"
#include <cooperative_groups.h>

namespace iv {

constexpr int RS_THREADS = 256;
constexpr int RS_ITEMS_BIG = 16;   // 4096 keys per block: large inputs
constexpr int RS_ITEMS_SMALL = 4;  // 1024 keys per block: small inputs still fill the 148 SMs
constexpr int RS_RADIX = 256;
constexpr int RS_WARPS = RS_THREADS / 32;

struct SortJobs {
    const uint64_t* kin[RS_MAX_JOBS];
    uint64_t* kout[RS_MAX_JOBS];
    const void* vin[RS_MAX_JOBS];
    void* vout[RS_MAX_JOBS];
    int vb[RS_MAX_JOBS];
};

// hist[job][digit][tile]
template <int RS_ITEMS>
__global__ void __launch_bounds__(RS_THREADS)
k_rs_hist(SortJobs J, int64_t n, int shift, uint32_t mask, uint32_t* __restrict__ hist, int64_t ntiles) {
    __shared__ uint32_t h[RS_RADIX];
    const int tid = threadIdx.x, job = blockIdx.y;
    const uint64_t* __restrict__ keys = J.kin[job];
    constexpr int RS_TILE = RS_THREADS * RS_ITEMS;
    h[tid] = 0;
    __syncthreads();
    const int64_t tbase = (int64_t)blockIdx.x * RS_TILE;
#pragma unroll
    for (int it = 0; it < RS_ITEMS; it++) {
        int64_t idx = tbase + it * RS_THREADS + tid;
        if (idx < n) atomicAdd(&h[(uint32_t)((__ldg(keys + idx) >> shift) & mask)], 1u);
    }
    __syncthreads();
    hist[((int64_t)job * RS_RADIX + tid) * ntiles + blockIdx.x] = h[tid];
}

// One warp per (job, digit) row of the histogram table: exclusive prefix along the tiles (in place) and the
// row total.  With the exclusive scan of the 256 totals (done by every scatter block for itself) this replaces a
// device-wide scan of the table: three launches per digit whatever the input size.
__global__ void __launch_bounds__(RS_THREADS)
k_rs_prefix(uint32_t* __restrict__ hist, int64_t ntiles, uint32_t* __restrict__ totals, int nrows) {
    const int row = blockIdx.x * RS_WARPS + (threadIdx.x >> 5);
    if (row >= nrows) return;
    const int lane = threadIdx.x & 31;
    uint32_t* h = hist + (int64_t)row * ntiles;
    uint32_t run = 0;
    for (int64_t base = 0; base < ntiles; base += 32) {
        const int64_t t = base + lane;
        const uint32_t v = t < ntiles ? h[t] : 0u;
        const uint32_t inc = warp_incl_sum(v);
        if (t < ntiles) h[t] = run + inc - v;
        run += __shfl_sync(FULL, inc, 31);
    }
    if (lane == 0) totals[row] = run;
}
// Long rows (large inputs: tens of thousands of tiles): one BLOCK per row, every thread scans a contiguous piece --
// 256 x 8 resident warps instead of 256 warps walking 32 entries at a time.
__global__ void __launch_bounds__(RS_THREADS)
k_rs_prefix_long(uint32_t* __restrict__ hist, int64_t ntiles, uint32_t* __restrict__ totals) {
    __shared__ uint32_t wsum[RS_WARPS + 1];
    uint32_t* h = hist + (int64_t)blockIdx.x * ntiles;
    const int64_t per = (ntiles + RS_THREADS - 1) / RS_THREADS;
    const int64_t a = (int64_t)threadIdx.x * per, b = a + per < ntiles ? a + per : ntiles;
    uint32_t sum = 0;
    for (int64_t t = a; t < b; t++) sum += h[t];
    uint32_t tot;
    uint32_t run = block_excl_sum<uint32_t, RS_THREADS>(sum, wsum, tot);
    for (int64_t t = a; t < b; t++) {
        const uint32_t v = h[t];
        h[t] = run;
        run += v;
    }
    if (threadIdx.x == 0) totals[blockIdx.x] = tot;
}

template <int RS_ITEMS>
__global__ void __launch_bounds__(RS_THREADS, RS_ITEMS > 4 ? 4 : 1)  // 16 keys per thread would take 112 registers: cap at 64 for 4 blocks/SM
k_rs_scatter(SortJobs J, int64_t n, int shift, uint32_t mask, const uint32_t* __restrict__ hist,
             const uint32_t* __restrict__ totals, int64_t ntiles) {
    constexpr int RS_TILE = RS_THREADS * RS_ITEMS;
    __shared__ uint32_t warp_cnt[RS_WARPS][RS_RADIX];
    __shared__ uint32_t digit_base[RS_RADIX];
    __shared__ uint32_t gbase[RS_RADIX];
    __shared__ uint64_t sbuf[RS_TILE];
    __shared__ uint8_t sdig[RS_TILE];
    __shared__ uint32_t wsum[RS_WARPS + 1];

    const int tid = threadIdx.x, w = tid >> 5, lane = tid & 31, job = blockIdx.y;
    const uint64_t* __restrict__ kin = J.kin[job];
    uint64_t* __restrict__ kout = J.kout[job];
    const int VB = J.vb[job];
    const int64_t tile = blockIdx.x;
    const int64_t tbase = tile * RS_TILE;
    const int nvalid = (int)((n - tbase) < (int64_t)RS_TILE ? (n - tbase) : (int64_t)RS_TILE);

    for (int i = tid; i < RS_WARPS * RS_RADIX; i += RS_THREADS) ((uint32_t*)warp_cnt)[i] = 0;
    __syncthreads();

    uint64_t key[RS_ITEMS];
    uint32_t rank[RS_ITEMS];
#pragma unroll
    for (int it = 0; it < RS_ITEMS; it++) {
        int li = w * (RS_ITEMS * 32) + it * 32 + lane;
        key[it] = li < nvalid ? __ldg(kin + tbase + li) : ~0ull;
    }
#pragma unroll
    for (int it = 0; it < RS_ITEMS; it++) {
        uint32_t d = (uint32_t)((key[it] >> shift) & mask);
        unsigned m = __match_any_sync(FULL, d);
        int leader = __ffs(m) - 1;
        uint32_t c = 0;
        if (lane == leader) {
            c = warp_cnt[w][d];
            warp_cnt[w][d] = c + (uint32_t)__popc(m);
        }
        c = __shfl_sync(FULL, c, leader);
        rank[it] = c + (uint32_t)__popc(m & ((1u << lane) - 1u));
        __syncwarp();
    }
    __syncthreads();
    {
        // thread tid owns digit tid: exclusive scan over warps, then over digits
        uint32_t run = 0;
#pragma unroll
        for (int ww = 0; ww < RS_WARPS; ww++) {
            uint32_t c = warp_cnt[ww][tid];
            warp_cnt[ww][tid] = run;
            run += c;
        }
        uint32_t tot;
        uint32_t ex = block_excl_sum<uint32_t, RS_THREADS>(run, wsum, tot);
        digit_base[tid] = ex;
        // global offset of (digit, tile) = digits before (scan of the row totals) + same digit in earlier tiles
        uint32_t tt;
        const uint32_t dstart = block_excl_sum<uint32_t, RS_THREADS>(__ldg(totals + job * RS_RADIX + tid), wsum, tt);
        gbase[tid] = dstart + __ldg(hist + ((int64_t)job * RS_RADIX + tid) * ntiles + tile) - ex;
    }
    __syncthreads();
    uint32_t pos[RS_ITEMS];
#pragma unroll
    for (int it = 0; it < RS_ITEMS; it++) {
        uint32_t d = (uint32_t)((key[it] >> shift) & mask);
        pos[it] = digit_base[d] + warp_cnt[w][d] + rank[it];
        sbuf[pos[it]] = key[it];
        sdig[pos[it]] = (uint8_t)d;
    }
    __syncthreads();
    for (int k = tid; k < nvalid; k += RS_THREADS) kout[gbase[sdig[k]] + (uint32_t)k] = sbuf[k];
    if (VB == 8) {
        __syncthreads();
        const uint64_t* vi = (const uint64_t*)J.vin[job];
        uint64_t* vo = (uint64_t*)J.vout[job];
#pragma unroll
        for (int it = 0; it < RS_ITEMS; it++) {
            int li = w * (RS_ITEMS * 32) + it * 32 + lane;
            if (li < nvalid) sbuf[pos[it]] = __ldg(vi + tbase + li);
        }
        __syncthreads();
        for (int k = tid; k < nvalid; k += RS_THREADS) vo[gbase[sdig[k]] + (uint32_t)k] = sbuf[k];
    } else if (VB == 4) {
        __syncthreads();
        const uint32_t* vi = (const uint32_t*)J.vin[job];
        uint32_t* vo = (uint32_t*)J.vout[job];
        uint32_t* sb32 = (uint32_t*)sbuf;
#pragma unroll
        for (int it = 0; it < RS_ITEMS; it++) {
            int li = w * (RS_ITEMS * 32) + it * 32 + lane;
            if (li < nvalid) sb32[pos[it]] = __ldg(vi + tbase + li);
        }
        __syncthreads();
        for (int k = tid; k < nvalid; k += RS_THREADS) vo[gbase[sdig[k]] + (uint32_t)k] = sb32[k];
    }
}

// ---- device-wide exclusive sum of uint32 (histogram table) ---------------------------------------------
constexpr int SC_THREADS = 256;
constexpr int SC_ITEMS = 16;
constexpr int SC_TILE = SC_THREADS * SC_ITEMS;

__global__ void __launch_bounds__(SC_THREADS)
k_u32_tile_sums(const uint32_t* __restrict__ in, int64_t n, uint32_t* __restrict__ partials) {
    __shared__ uint32_t sm[SC_THREADS / 32];
    int64_t base = (int64_t)blockIdx.x * SC_TILE;
    uint32_t s = 0;
#pragma unroll
    for (int it = 0; it < SC_ITEMS; it++) {
        int64_t i = base + it * SC_THREADS + threadIdx.x;
        if (i < n) s += __ldg(in + i);
    }
    s = warp_sum(s);
    if (lane_id() == 0) sm[threadIdx.x >> 5] = s;
    __syncthreads();
    if (threadIdx.x < 32) {
        uint32_t x = threadIdx.x < SC_THREADS / 32 ? sm[threadIdx.x] : 0;
        x = warp_sum(x);
        if (threadIdx.x == 0) partials[blockIdx.x] = x;
    }
}

// single block: in-place exclusive scan of m values.  Every warp owns a contiguous chunk of 32 * ITEMS values per
// round and reads it with coalesced loads issued together; shuffle scans inside the warp, one pass over the 32 warp
// totals, two barriers per round.
template <typename T, int ITEMS>
__global__ void __launch_bounds__(1024) k_scan_small_t(T* __restrict__ data, int64_t m, T* __restrict__ out_total) {
    __shared__ T sm[32];
    const int w = threadIdx.x >> 5, lane = threadIdx.x & 31;
    T carry = 0;
    for (int64_t base = 0; base < m; base += 1024 * ITEMS) {
        const int64_t wbase = base + (int64_t)w * (32 * ITEMS);
        T v[ITEMS];
#pragma unroll
        for (int j = 0; j < ITEMS; j++) {
            const int64_t i = wbase + j * 32 + lane;
            v[j] = i < m ? data[i] : T(0);
        }
        T run = 0;
#pragma unroll
        for (int j = 0; j < ITEMS; j++) {
            const T inc = warp_incl_sum(v[j]);
            v[j] = run + inc - v[j];  // exclusive inside the warp's chunk
            run += __shfl_sync(FULL, inc, 31);
        }
        if (lane == 0) sm[w] = run;
        __syncthreads();
        T x = sm[lane];                       // 32 warps
        const T xi = warp_incl_sum(x);
        const T before = carry + __shfl_sync(FULL, xi - x, w);
        const T tot = __shfl_sync(FULL, xi, 31);
#pragma unroll
        for (int j = 0; j < ITEMS; j++) {
            const int64_t i = wbase + j * 32 + lane;
            if (i < m) data[i] = before + v[j];
        }
        carry += tot;
        __syncthreads();
    }
    if (threadIdx.x == 0 && out_total) *out_total = carry;
}
template <typename T>
static void scan_small(T* data, int64_t m, T* out_total, cudaStream_t stream) {
    if (m > 4096) k_scan_small_t<T, 16><<<1, 1024, 0, stream>>>(data, m, out_total);
    else k_scan_small_t<T, 4><<<1, 1024, 0, stream>>>(data, m, out_total);
}

// Tile-local exclusive scan with coalesced accesses: warp w of the block owns values [w * 32 * SC_ITEMS, ...) of the
// tile, lane l reads / writes elements j * 32 + l of that chunk; shuffle scans inside the chunk, one pass over the
// warp totals.  `first` = exclusive prefix of the tile (its scanned partial).
template <typename T>
__device__ __forceinline__ void tile_scan_apply(T* __restrict__ data, int64_t n, T first, T* sm /* [SC_THREADS/32] */) {
    constexpr int W = SC_THREADS / 32;
    const int w = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int64_t wbase = (int64_t)blockIdx.x * SC_TILE + (int64_t)w * (32 * SC_ITEMS);
    T v[SC_ITEMS];
#pragma unroll
    for (int j = 0; j < SC_ITEMS; j++) {
        const int64_t i = wbase + j * 32 + lane;
        v[j] = i < n ? data[i] : T(0);
    }
    T run = 0;
#pragma unroll
    for (int j = 0; j < SC_ITEMS; j++) {
        const T inc = warp_incl_sum(v[j]);
        v[j] = run + inc - v[j];
        run += __shfl_sync(FULL, inc, 31);
    }
    if (lane == 0) sm[w] = run;
    __syncthreads();
    T x = lane < W ? sm[lane] : T(0);
    const T xi = warp_incl_sum(x);
    const T before = first + __shfl_sync(FULL, xi - x, w);
#pragma unroll
    for (int j = 0; j < SC_ITEMS; j++) {
        const int64_t i = wbase + j * 32 + lane;
        if (i < n) data[i] = before + v[j];
    }
}

__global__ void __launch_bounds__(SC_THREADS)
k_u32_apply(uint32_t* __restrict__ data, int64_t n, const uint32_t* __restrict__ partials) {
    __shared__ uint32_t sm[SC_THREADS / 32];
    tile_scan_apply<uint32_t>(data, n, __ldg(partials + blockIdx.x), sm);
}

size_t scan_u32_large_temp_elems(int64_t n) { return (size_t)cdiv(n > 0 ? n : 1, SC_TILE) + 1; }

int scan_exclusive_u32_large(uint32_t* data, int64_t n, uint32_t* temp_partials, cudaStream_t stream) {
    if (n <= 0) return IV_OK;
    int64_t nt = cdiv(n, SC_TILE);
    if (nt == 1) {
        scan_small<uint32_t>(data, n, nullptr, stream);
        IV_LAUNCH_CHECK();
        return IV_OK;
    }
    k_u32_tile_sums<<<(unsigned)nt, SC_THREADS, 0, stream>>>(data, n, temp_partials);
    IV_LAUNCH_CHECK();
    scan_small<uint32_t>(temp_partials, nt, nullptr, stream);
    IV_LAUNCH_CHECK();
    k_u32_apply<<<(unsigned)nt, SC_THREADS, 0, stream>>>(data, n, temp_partials);
    IV_LAUNCH_CHECK();
    return IV_OK;
}

// device-wide exclusive sum of int64 for long arrays: tile sums -> single-block scan of the sums -> apply
__global__ void __launch_bounds__(SC_THREADS)
k_i64_tile_sums(const int64_t* __restrict__ in, int64_t n, int64_t* __restrict__ partials) {
    __shared__ int64_t sm[SC_THREADS / 32];
    int64_t base = (int64_t)blockIdx.x * SC_TILE;
    int64_t s = 0;
#pragma unroll
    for (int it = 0; it < SC_ITEMS; it++) {
        int64_t i = base + it * SC_THREADS + threadIdx.x;
        if (i < n) s += __ldg(in + i);
    }
    s = warp_sum(s);
    if (lane_id() == 0) sm[threadIdx.x >> 5] = s;
    __syncthreads();
    if (threadIdx.x < 32) {
        int64_t x = threadIdx.x < SC_THREADS / 32 ? sm[threadIdx.x] : 0;
        x = warp_sum(x);
        if (threadIdx.x == 0) partials[blockIdx.x] = x;
    }
}
__global__ void __launch_bounds__(SC_THREADS)
k_i64_apply(int64_t* __restrict__ data, int64_t n, const int64_t* __restrict__ partials) {
    __shared__ int64_t sm[SC_THREADS / 32];
    tile_scan_apply<int64_t>(data, n, __ldg(partials + blockIdx.x), sm);
}
size_t scan_i64_large_temp_elems(int64_t n) { return (size_t)cdiv(n > 0 ? n : 1, SC_TILE) + 2; }
int scan_exclusive_i64_large(int64_t* data, int64_t n, int64_t* partials, int64_t* out_total, cudaStream_t stream) {
    if (n <= 4 * SC_TILE) return scan_exclusive_i64_inplace(data, n, out_total, stream);
    int64_t nt = cdiv(n, SC_TILE);
    k_i64_tile_sums<<<(unsigned)nt, SC_THREADS, 0, stream>>>(data, n, partials);
    IV_LAUNCH_CHECK();
    scan_small<int64_t>(partials, nt, out_total, stream);
    IV_LAUNCH_CHECK();
    k_i64_apply<<<(unsigned)nt, SC_THREADS, 0, stream>>>(data, n, partials);
    IV_LAUNCH_CHECK();
    return IV_OK;
}

int scan_exclusive_i64_inplace(int64_t* data, int64_t n, int64_t* out_total, cudaStream_t stream) {
    scan_small<int64_t>(data, n > 0 ? n : 0, out_total, stream);
    IV_LAUNCH_CHECK();
    return IV_OK;
}

// ---- small inputs: one partition pass + one bucket pass ------------------------------------------------------
// A few hundred thousand keys are bound by the NUMBER of dependent launches (three per digit), not by bytes.
// Such inputs take one stable partition on their top 8 bits (the three kernels above) and then ONE launch in which
// every block sorts one of the 256 buckets on the remaining bits with the same stable LSD radix, all digits in one
// go: in shared memory (composite "low bits | position in the bucket" words, 1024-key chunks ranked with
// __match_any_sync) when the bucket fits, through global memory when it does not (heavily skewed keys: slower,
// still correct).  5 launches instead of 3 per digit; the sorted data ends in the first buffer.
constexpr int BK_THREADS = 256;
constexpr int BK_WARPS = BK_THREADS / 32;
constexpr int BK_ITEMS = 4;
constexpr int BK_CHUNK = BK_THREADS * BK_ITEMS;  // keys ranked per round
constexpr int BK_CAP = 4096;                     // keys a block sorts in shared memory
constexpr int BK_POS_BITS = 12;                  // log2(BK_CAP)
constexpr int BK_MAX_LOW_BITS = 64 - BK_POS_BITS;
constexpr int64_t BK_MAX_N = (int64_t)1 << 18;
constexpr int64_t BK_MIN_N = 4096;

// One stable counting pass of m keys on digit (key >> shift) & dmask: src -> dst (shared or global memory),
// payloads moved alongside when given.  base[256], warp_cnt[BK_WARPS][256]: shared scratch.
template <bool WITH_VALUES>
__device__ __forceinline__ void bk_pass(const uint64_t* src, uint64_t* dst, uint32_t m, int shift, uint32_t dmask,
                                        uint32_t* base, uint32_t* warp_cnt, const char* vsrc, char* vdst, int VB) {
    const int tid = threadIdx.x, w = tid >> 5, lane = tid & 31;
    if (tid < RS_RADIX) base[tid] = 0;
    __syncthreads();
    for (uint32_t i = tid; i < m; i += BK_THREADS) atomicAdd(&base[(uint32_t)(src[i] >> shift) & dmask], 1u);
    __syncthreads();
    if (tid < 32) {  // exclusive scan of the 256 digit counts
        uint32_t run = 0;
        for (int d0 = 0; d0 < RS_RADIX; d0 += 32) {
            const uint32_t v = base[d0 + tid];
            const uint32_t inc = warp_incl_sum(v);
            base[d0 + tid] = run + inc - v;
            run += __shfl_sync(FULL, inc, 31);
        }
    }
    __syncthreads();
    for (uint32_t c0 = 0; c0 < m; c0 += BK_CHUNK) {
        for (int i = tid; i < BK_WARPS * RS_RADIX; i += BK_THREADS) warp_cnt[i] = 0;
        __syncthreads();
        uint64_t key[BK_ITEMS];
        uint32_t rank[BK_ITEMS], dig[BK_ITEMS];
#pragma unroll
        for (int it = 0; it < BK_ITEMS; it++) {  // warp w owns 128 consecutive keys of the chunk, 32 per item
            const uint32_t i = c0 + w * (BK_ITEMS * 32) + it * 32 + lane;
            const bool valid = i < m;
            key[it] = valid ? src[i] : 0ull;
            const uint32_t d = valid ? ((uint32_t)(key[it] >> shift) & dmask) : (0x100u | lane);
            dig[it] = d;
            const unsigned mm = __match_any_sync(FULL, d);
            const int leader = __ffs(mm) - 1;
            uint32_t c = 0;
            if (valid && lane == leader) {
                c = warp_cnt[w * RS_RADIX + d];
                warp_cnt[w * RS_RADIX + d] = c + (uint32_t)__popc(mm);
            }
            c = __shfl_sync(FULL, c, leader);
            rank[it] = c + (uint32_t)__popc(mm & ((1u << lane) - 1u));
            __syncwarp();
        }
        __syncthreads();
        if (tid < RS_RADIX) {  // digit tid: offsets of the warps, in order; the running total carries into the next chunk
            uint32_t run = base[tid];
#pragma unroll
            for (int ww = 0; ww < BK_WARPS; ww++) {
                const uint32_t c = warp_cnt[ww * RS_RADIX + tid];
                warp_cnt[ww * RS_RADIX + tid] = run;
                run += c;
            }
            base[tid] = run;
        }
        __syncthreads();
#pragma unroll
        for (int it = 0; it < BK_ITEMS; it++) {
            const uint32_t i = c0 + w * (BK_ITEMS * 32) + it * 32 + lane;
            if (i < m) {
                const uint32_t pos = warp_cnt[w * RS_RADIX + dig[it]] + rank[it];
                dst[pos] = key[it];
                if (WITH_VALUES) {
                    if (VB == 4) ((uint32_t*)vdst)[pos] = ((const uint32_t*)vsrc)[i];
                    else if (VB == 8) ((uint64_t*)vdst)[pos] = ((const uint64_t*)vsrc)[i];
                }
            }
        }
        __syncthreads();
    }
}

__global__ void __launch_bounds__(BK_THREADS)
k_rs_buckets(SortJobs J, int64_t n, int begin_bit, int low_bits, const uint32_t* __restrict__ totals) {
    extern __shared__ __align__(16) unsigned char bk_smem[];
    uint64_t* sa = reinterpret_cast<uint64_t*>(bk_smem);                 // [BK_CAP]
    uint64_t* sb = sa + BK_CAP;                                          // [BK_CAP]
    uint32_t* warp_cnt = reinterpret_cast<uint32_t*>(sb + BK_CAP);       // [BK_WARPS][256]
    uint32_t* base = warp_cnt + BK_WARPS * RS_RADIX;                     // [256]
    __shared__ uint32_t s_b0, s_m;
    const int tid = threadIdx.x, job = blockIdx.y, bucket = blockIdx.x;
    // after the partition pass the data is in the second buffer
    const uint64_t* __restrict__ kin = J.kin[job];
    uint64_t* __restrict__ kout = J.kout[job];
    const int VB = J.vb[job];
    if (tid < 32) {  // bucket range = prefix of the digit totals
        uint32_t before = 0, mine = 0;
        for (int d = tid; d < RS_RADIX; d += 32) {
            const uint32_t t = __ldg(totals + job * RS_RADIX + d);
            if (d < bucket) before += t;
            if (d == bucket) mine = t;
        }
        before = warp_sum(before);
        mine = warp_sum(mine);
        if (tid == 0) { s_b0 = before; s_m = mine; }
    }
    __syncthreads();
    const uint32_t b0 = s_b0, m = s_m;
    if (m == 0) return;
    if (m <= BK_CAP) {
        const uint64_t low_mask = (1ull << low_bits) - 1ull;  // low_bits <= 52
        for (uint32_t i = tid; i < m; i += BK_THREADS)
            sa[i] = (((__ldg(kin + b0 + i) >> begin_bit) & low_mask) << BK_POS_BITS) | i;
        __syncthreads();
        uint64_t *a = sa, *b = sb;
        for (int sh = 0; sh < low_bits; sh += 8) {
            const int bits = low_bits - sh < 8 ? low_bits - sh : 8;
            bk_pass<false>(a, b, m, BK_POS_BITS + sh, (1u << bits) - 1u, base, warp_cnt, nullptr, nullptr, 0);
            uint64_t* t = a; a = b; b = t;
        }
        for (uint32_t i = tid; i < m; i += BK_THREADS) {
            const uint32_t src = b0 + (uint32_t)(a[i] & ((1u << BK_POS_BITS) - 1u));
            kout[b0 + i] = __ldg(kin + src);
            if (VB == 4) ((uint32_t*)J.vout[job])[b0 + i] = __ldg((const uint32_t*)J.vin[job] + src);
            else if (VB == 8) ((uint64_t*)J.vout[job])[b0 + i] = __ldg((const uint64_t*)J.vin[job] + src);
        }
        return;
    }
    // oversize bucket: the same passes through global memory, ping-ponging between the two buffers inside the
    // bucket's range
    uint64_t* ka = const_cast<uint64_t*>(kin) + b0;
    uint64_t* kb = kout + b0;
    char* va = VB ? (char*)J.vin[job] + (size_t)b0 * VB : nullptr;
    char* vb = VB ? (char*)J.vout[job] + (size_t)b0 * VB : nullptr;
    for (int sh = 0; sh < low_bits; sh += 8) {
        const int bits = low_bits - sh < 8 ? low_bits - sh : 8;
        bk_pass<true>(ka, kb, m, begin_bit + sh, (1u << bits) - 1u, base, warp_cnt, va, vb, VB);
        uint64_t* tk = ka; ka = kb; kb = tk;
        char* tv = va; va = vb; vb = tv;
    }
    if (ka != kout + b0) {  // even number of passes: the sorted bucket sits in the input buffer
        for (uint32_t i = tid; i < m; i += BK_THREADS) {
            kout[b0 + i] = ka[i];
            if (VB == 4) ((uint32_t*)J.vout[job])[b0 + i] = ((const uint32_t*)va)[i];
            else if (VB == 8) ((uint64_t*)J.vout[job])[b0 + i] = ((const uint64_t*)va)[i];
        }
    }
}
constexpr size_t BK_SMEM = (size_t)BK_CAP * 16 + (size_t)(BK_WARPS + 1) * RS_RADIX * 4;  // 73 KB

static bool bucket_path(int64_t n, int nbits) {
    return n >= BK_MIN_N && n <= BK_MAX_N && nbits > 8 && nbits - 8 <= BK_MAX_LOW_BITS;
}

// ---- large inputs: one-sweep partition passes on the TOP bits + one shared-memory pass per bucket ------------------
// Above BK_MAX_N keys the sort is bound by bytes moved per pass, so the number of passes over global memory is what
// counts.  (1) ONE read of the keys gives the global digit histograms of every partition pass (k_rs_ghist).
// (2) P stable partition passes (LSD order) over the top T bits only, T ~ log2(n / 1024): each tile takes a ticket,
// ranks its keys, publishes its 256 digit counts and obtains the counts of all earlier tiles by a decoupled look-back
// over the status words (2 flag bits + 30 value bits in one 32-bit word, so no fence is needed) -- keys are read once
// and written once per pass, no per-tile histogram pass.  (3) The 2^T buckets (~1000 keys each) are found by binary
// search (k_hy_starts) and every bucket is finished by one block in ONE pass whatever the number of remaining bits
// (k_hy_buckets): composite words "low bits | position in the bucket" are all distinct, so a counting pass on the 12
// bits below the highest bit that differs inside the bucket (atomics, unordered inside a bin) followed by a
// rank-by-counting inside every bin (bins hold ~1 key) IS the stable order.  42-bit keys of 50 M rows: 3 passes over
// global memory instead of 6 x (histogram + scatter).  A tile never waits for a tile that is not running (tickets).
#ifndef HY_CAP_LOG
#define HY_CAP_LOG 12
#endif
constexpr int HY_CAP = 1 << HY_CAP_LOG;   // keys a bucket block sorts in shared memory
constexpr int HY_POS_BITS = HY_CAP_LOG;
constexpr int HY_DIG_BITS = HY_CAP_LOG;   // counting digit of the bucket pass
constexpr int HY_BINS = 1 << HY_DIG_BITS;
constexpr int HY_THREADS = 256;
constexpr int HY_ITEMS = HY_CAP / HY_THREADS;
constexpr int HY_MAX_P = 4;
constexpr int HY_MAX_T = 20;
constexpr uint32_t HY_HUGE = 1u << 16;  // keys: a pile-up beyond this is sorted by all blocks together (k_hy_huge)
constexpr int64_t HY_MAX_N = ((int64_t)1 << 30) - 1;
constexpr uint32_t SW_AGG = 1u << 30, SW_INC = 2u << 30, SW_VAL = (1u << 30) - 1u;
constexpr int SW_ITEMS = 16;
constexpr int SW_TILE = RS_THREADS * SW_ITEMS;

struct HyPlan {
    bool use;
    int P;                 // partition passes
    int T;                 // bits they cover: [end_bit - T, end_bit)
    int low_bits;          // bits left to the bucket pass (0: the partition passes already sort everything)
    int shift[HY_MAX_P];   // LSD order
    uint32_t mask[HY_MAX_P];
};
static HyPlan hy_plan(int64_t n, int begin_bit, int end_bit) {
    HyPlan p;
    memset(&p, 0, sizeof(p));
    const int nbits = end_bit - begin_bit;
    if (n <= BK_MAX_N || n > HY_MAX_N || nbits <= 0) return p;
    int T = 1;
    while (T < HY_MAX_T && ((int64_t)(HY_CAP / 4) << T) < n) T++;  // <= CAP / 4 keys per bucket if all were used
    if (T >= nbits) T = nbits;
    else if (nbits - T > 64 - HY_POS_BITS) T = nbits - (64 - HY_POS_BITS);
    p.use = true;
    p.T = T;
    p.low_bits = nbits - T;
    p.P = (T + 7) / 8;
    int lo = end_bit - T;
    for (int i = 0; i < p.P; i++) {  // balanced digit widths
        const int w = (T - (lo - (end_bit - T)) + (p.P - i) - 1) / (p.P - i);
        p.shift[i] = lo;
        p.mask[i] = (1u << w) - 1u;
        lo += w;
    }
    return p;
}
struct HyDigits {
    int P;
    int shift[HY_MAX_P];
    uint32_t mask[HY_MAX_P];
};

__global__ void __launch_bounds__(RS_THREADS)
k_rs_ghist(const uint64_t* __restrict__ keys, int64_t n, HyDigits D, uint32_t* __restrict__ ghist, bool vec) {
    __shared__ uint32_t h[HY_MAX_P][RS_RADIX];
    for (int i = threadIdx.x; i < HY_MAX_P * RS_RADIX; i += RS_THREADS) (&h[0][0])[i] = 0;
    __syncthreads();
    const int64_t stride = (int64_t)gridDim.x * RS_THREADS;
    for (int64_t i = (int64_t)blockIdx.x * RS_THREADS + threadIdx.x; 2 * i < n; i += stride) {
        uint64_t a, b = 0;
        const bool two = 2 * i + 1 < n;
        if (two && vec) {
            const ulonglong2 v = __ldg(reinterpret_cast<const ulonglong2*>(keys) + i);
            a = v.x;
            b = v.y;
        } else if (two) {
            a = __ldg(keys + 2 * i);
            b = __ldg(keys + 2 * i + 1);
        } else {
            a = __ldg(keys + 2 * i);
        }
#pragma unroll
        for (int p = 0; p < HY_MAX_P; p++) {
            if (p < D.P) {
                atomicAdd(&h[p][(uint32_t)(a >> D.shift[p]) & D.mask[p]], 1u);
                if (two) atomicAdd(&h[p][(uint32_t)(b >> D.shift[p]) & D.mask[p]], 1u);
            }
        }
    }
    __syncthreads();
    for (int i = threadIdx.x; i < D.P * RS_RADIX; i += RS_THREADS) {
        const uint32_t v = (&h[0][0])[i];
        if (v) atomicAdd(ghist + i, v);
    }
}

// One stable partition pass on digit (key >> shift) & mask.  VB: payload bytes (0 / 4 / 8); iota: the payload of
// input position i is i itself (first pass of an argsort: nothing is read).
template <int VB> struct SwVal { typedef uint32_t T; };
template <> struct SwVal<8> { typedef uint64_t T; };

struct SwSmem {
    uint32_t warp_cnt[RS_WARPS][RS_RADIX];
    uint32_t digit_base[RS_RADIX];
    uint32_t gbase[RS_RADIX];
    uint64_t sbuf[SW_TILE];
    uint8_t sdig[SW_TILE];
    uint32_t wsum[RS_WARPS + 1];
    uint32_t s_tile;
};

// The tile body shared by the partition pass and by the pile-up path of the bucket pass: SW_TILE keys starting at
// kin + tbase are ranked by digit (warp-level multi-split), staged through shared memory in digit order and written
// as coalesced runs to kout / vout at gbase[digit] + rank.  `global_base(run, ex)` is called by thread d for digit d
// with the tile's count of that digit and the exclusive prefix of the counts inside the tile, and returns the
// position of the tile's first key of that digit minus ex.  S.warp_cnt must be zero on entry.
template <int VB, typename GB>
__device__ __forceinline__ void sw_tile(SwSmem& S, const uint64_t* __restrict__ kin, uint64_t* __restrict__ kout,
                                        const void* __restrict__ vin, void* __restrict__ vout, int64_t tbase,
                                        int nvalid, int shift, uint32_t mask, int iota, GB global_base) {
    typedef typename SwVal<VB>::T V;
    const int tid = threadIdx.x, w = tid >> 5, lane = tid & 31;
    uint64_t key[SW_ITEMS];
    V val[VB ? SW_ITEMS : 1];
    uint32_t rank[SW_ITEMS];
#pragma unroll
    for (int it = 0; it < SW_ITEMS; it++) {
        const int li = w * (SW_ITEMS * 32) + it * 32 + lane;
        key[it] = li < nvalid ? kin[tbase + li] : ~0ull;
    }
    if (VB) {  // the payloads travel in registers: every load of the tile is in flight before the ranking starts
#pragma unroll
        for (int it = 0; it < SW_ITEMS; it++) {
            const int li = w * (SW_ITEMS * 32) + it * 32 + lane;
            val[it] = iota ? (V)(tbase + li) : (li < nvalid ? ((const V*)vin)[tbase + li] : V(0));
        }
    }
    // warp-level multi-split, eight keys at a time: the counter updates (shared-memory atomics by the leader of
    // every group of equal digits) are issued back to back and their results are only needed afterwards
#pragma unroll
    for (int h = 0; h < SW_ITEMS; h += 8) {
        unsigned mm[8];
        uint32_t cc[8];
#pragma unroll
        for (int j = 0; j < 8; j++) {
            const uint32_t d = (uint32_t)(key[h + j] >> shift) & mask;
            mm[j] = __match_any_sync(FULL, d);
            cc[j] = 0;
            if (lane == __ffs(mm[j]) - 1) cc[j] = atomicAdd(&S.warp_cnt[w][d], (uint32_t)__popc(mm[j]));
            __syncwarp();
        }
#pragma unroll
        for (int j = 0; j < 8; j++)
            rank[h + j] = __shfl_sync(FULL, cc[j], __ffs(mm[j]) - 1) + (uint32_t)__popc(mm[j] & ((1u << lane) - 1u));
    }
    __syncthreads();
    {
        // thread tid owns digit tid.  (The padding keys of a ragged last tile count as digit `mask`: they are the
        // last keys of the tile, so they only inflate that one count behind every real key.)
        uint32_t run = 0;
#pragma unroll
        for (int ww = 0; ww < RS_WARPS; ww++) {
            const uint32_t c = S.warp_cnt[ww][tid];
            S.warp_cnt[ww][tid] = run;
            run += c;
        }
        uint32_t tot;
        const uint32_t ex = block_excl_sum<uint32_t, RS_THREADS>(run, S.wsum, tot);
        S.digit_base[tid] = ex;
        S.gbase[tid] = global_base(run, ex);
    }
    __syncthreads();
    uint32_t pos[SW_ITEMS];
#pragma unroll
    for (int it = 0; it < SW_ITEMS; it++) {
        const uint32_t d = (uint32_t)(key[it] >> shift) & mask;
        pos[it] = S.digit_base[d] + S.warp_cnt[w][d] + rank[it];
        S.sbuf[pos[it]] = key[it];
        S.sdig[pos[it]] = (uint8_t)d;
    }
    __syncthreads();
    for (int k = tid; k < nvalid; k += RS_THREADS) kout[S.gbase[S.sdig[k]] + (uint32_t)k] = S.sbuf[k];
    if (VB) {
        __syncthreads();
        V* sv = (V*)S.sbuf;
#pragma unroll
        for (int it = 0; it < SW_ITEMS; it++) sv[pos[it]] = val[it];
        __syncthreads();
        V* vo = (V*)vout;
        for (int k = tid; k < nvalid; k += RS_THREADS) vo[S.gbase[S.sdig[k]] + (uint32_t)k] = sv[k];
    }
}

template <int VB>
__global__ void __launch_bounds__(RS_THREADS, VB == 8 ? 2 : (VB ? 3 : 4))
k_rs_sweep(const uint64_t* __restrict__ kin, uint64_t* __restrict__ kout, const void* __restrict__ vin,
           void* __restrict__ vout, int64_t n, int shift, uint32_t mask, const uint32_t* __restrict__ ghist,
           uint32_t* status, uint32_t* ticket, int iota) {
    __shared__ SwSmem S;
    const int tid = threadIdx.x;
    if (tid == 0) S.s_tile = atomicAdd(ticket, 1u);
    for (int i = tid; i < RS_WARPS * RS_RADIX; i += RS_THREADS) (&S.warp_cnt[0][0])[i] = 0;
    __syncthreads();
    const int64_t tile = S.s_tile;
    const int64_t tbase = tile * SW_TILE;
    const int nvalid = (int)((n - tbase) < (int64_t)SW_TILE ? (n - tbase) : (int64_t)SW_TILE);
    sw_tile<VB>(S, kin, kout, vin, vout, tbase, nvalid, shift, mask, iota, [&](uint32_t run, uint32_t ex) {
        volatile uint32_t* st = status;
        st[tile * RS_RADIX + tid] = (tile == 0 ? SW_INC : SW_AGG) | run;
        uint32_t tt;
        const uint32_t dstart = block_excl_sum<uint32_t, RS_THREADS>(__ldg(ghist + tid), S.wsum, tt);
        uint32_t excl = 0;
        if (tile > 0) {
            // decoupled look-back, eight predecessors per round trip
            int64_t t = tile - 1;
            unsigned ns = 20;
            bool done = false;
            while (!done) {
                uint32_t v[8];
#pragma unroll
                for (int j = 0; j < 8; j++) v[j] = t - j >= 0 ? st[(t - j) * RS_RADIX + tid] : (2u << 30);
                int used = 0;
#pragma unroll
                for (int j = 0; j < 8; j++) {
                    if (!done && used == j && (v[j] >> 30) != 0) {
                        excl += v[j] & SW_VAL;
                        used++;
                        if (v[j] & SW_INC) done = true;
                    }
                }
                t -= used;
                if (used == 0) {
                    __nanosleep(ns);
                    if (ns < 200) ns += 20;
                }
            }
            st[tile * RS_RADIX + tid] = SW_INC | (excl + run);
        }
        return dstart + excl - ex;
    });
}

// starts[b] = first position whose top-T field is >= b (b = 0 .. 2^T); the keys are sorted on that field
__global__ void __launch_bounds__(256)
k_hy_starts(const uint64_t* __restrict__ keys, int64_t n, int tshift, uint32_t tmask, uint32_t nb,
            uint32_t* __restrict__ starts) {
    const uint32_t b = blockIdx.x * 256u + threadIdx.x;
    if (b > nb) return;
    int64_t lo = 0, hi = n;  // first i in [0, n] with field(i) >= b
    while (lo < hi) {
        const int64_t mid = (lo + hi) >> 1;
        if (((uint32_t)(__ldg(keys + mid) >> tshift) & tmask) >= b) hi = mid; else lo = mid + 1;
    }
    starts[b] = (uint32_t)lo;
}

// Pile-up path of the bucket pass (a bucket beyond HY_CAP keys), kept out of line so that its registers do not
// weigh on the common path.
template <int VB>
__device__ __noinline__ void hy_pileup(unsigned char* hy_smem, unsigned long long& s_diff, uint32_t* base, uint32_t* wsum,
                                       uint64_t* kin, uint64_t* kout, void* vin, void* vout, int begin_bit, int low_bits,
                                       uint32_t b0, uint32_t m) {
    // oversize bucket (a pile-up): stable LSD passes through global memory inside the bucket's range, on the bits
    // that differ inside the bucket only; every pass = histogram + the tile body of the partition pass
    SwSmem& S = *reinterpret_cast<SwSmem*>(hy_smem);
    const int tid = threadIdx.x;
    const uint64_t low_mask = low_bits >= 64 ? ~0ull : ((1ull << low_bits) - 1ull);
    if (tid == 0) s_diff = 0ull;
    __syncthreads();
    {
        const uint64_t first = kin[b0];
        uint64_t diff = 0;
        for (uint32_t i = tid; i < m; i += HY_THREADS) diff |= kin[b0 + i] ^ first;
        diff = (diff >> begin_bit) & low_mask;
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) diff |= __shfl_xor_sync(FULL, diff, o);
        if ((tid & 31) == 0 && diff) atomicOr(&s_diff, (unsigned long long)diff);
    }
    __syncthreads();
    const unsigned long long dall = s_diff;
    uint64_t *ka = kin + b0, *kb = kout + b0;
    char* va = VB ? (char*)vin + (size_t)b0 * VB : nullptr;
    char* vb = VB ? (char*)vout + (size_t)b0 * VB : nullptr;
    if (dall) {
        const int lo = __ffsll((long long)dall) - 1, hi = 64 - __clzll((long long)dall);  // bits [lo, hi) differ
        const int npass = (hi - lo + 7) / 8;
        int at = lo;
        for (int p = 0; p < npass; p++) {
            const int bits = (hi - at + (npass - p) - 1) / (npass - p);
            const int shift = begin_bit + at;
            const uint32_t dmask = (1u << bits) - 1u;
            base[tid] = 0;
            __syncthreads();
            for (uint32_t i = tid; i < m; i += HY_THREADS) atomicAdd(&base[(uint32_t)(ka[i] >> shift) & dmask], 1u);
            __syncthreads();
            {
                uint32_t tot;
                const uint32_t v = base[tid];
                const uint32_t ex = block_excl_sum<uint32_t, HY_THREADS>(v, wsum, tot);
                base[tid] = ex;
            }
            __syncthreads();
            for (uint32_t c0 = 0; c0 < m; c0 += SW_TILE) {
                for (int i = tid; i < RS_WARPS * RS_RADIX; i += RS_THREADS) (&S.warp_cnt[0][0])[i] = 0;
                __syncthreads();
                const int nvalid = (int)(m - c0 < (uint32_t)SW_TILE ? m - c0 : (uint32_t)SW_TILE);
                sw_tile<VB>(S, ka, kb, va, vb, (int64_t)c0, nvalid, shift, dmask, 0, [&](uint32_t run, uint32_t ex) {
                    const uint32_t g = base[tid];
                    base[tid] = g + run;
                    return g - ex;
                });
                __syncthreads();
            }
            uint64_t* tk = ka; ka = kb; kb = tk;
            char* tv = va; va = vb; vb = tv;
            at += bits;
        }
    }
    if (ka != kout + b0) {
        for (uint32_t i = tid; i < m; i += HY_THREADS) {
            kout[b0 + i] = ka[i];
            if (VB == 4) ((uint32_t*)vout)[b0 + i] = ((const uint32_t*)va)[i];
            else if (VB == 8) ((uint64_t*)vout)[b0 + i] = ((const uint64_t*)va)[i];
        }
    }
}

template <int VB>
__global__ void __launch_bounds__(HY_THREADS, HY_CAP_LOG >= 12 ? 4 : 6)
k_hy_buckets(uint64_t* kin, uint64_t* kout, void* vin, void* vout, int begin_bit, int low_bits,
             const uint32_t* __restrict__ starts, uint32_t* __restrict__ pile_count, uint32_t* __restrict__ pile_list) {
    extern __shared__ __align__(16) unsigned char hy_smem[];
    uint64_t* c2 = reinterpret_cast<uint64_t*>(hy_smem);                    // [HY_CAP] composites grouped by bin
    uint32_t* cur = reinterpret_cast<uint32_t*>(c2 + HY_CAP);               // [HY_BINS] counts -> cursors -> bin ends
    __shared__ unsigned long long s_diff;
    __shared__ uint32_t wsum[HY_THREADS / 32 + 1];
    const int tid = threadIdx.x;
    const uint32_t b0 = __ldg(starts + blockIdx.x), m = __ldg(starts + blockIdx.x + 1) - b0;
    if (m == 0) return;
    if (m == 1) {
        if (tid == 0) {
            kout[b0] = kin[b0];
            if (VB == 4) ((uint32_t*)vout)[b0] = ((const uint32_t*)vin)[b0];
            if (VB == 8) ((uint64_t*)vout)[b0] = ((const uint64_t*)vin)[b0];
        }
        return;
    }
    if (m > (uint32_t)HY_CAP) {  // a pile-up: left to k_hy_pileups (one block) or, beyond HY_HUGE keys, to the whole grid
        if (tid == 0) {
            if (m > HY_HUGE) pile_list[(gridDim.x + 32) + atomicAdd(pile_count + 1, 1u)] = blockIdx.x;
            else pile_list[atomicAdd(pile_count, 1u)] = blockIdx.x;
        }
        return;
    }
    const uint64_t low_mask = low_bits >= 64 ? ~0ull : ((1ull << low_bits) - 1ull);
    uint64_t low[HY_ITEMS];
    // as many bins as keys (a power of two, 256 .. 4096): clearing and scanning the bins is the fixed cost of a bucket
    const int dbits = m <= 256 ? 8 : 32 - __clz(m - 1);
    const uint32_t nbins = 1u << dbits, per = nbins / HY_THREADS;
    if (tid == 0) s_diff = 0ull;
    for (uint32_t i = tid; i < nbins; i += HY_THREADS) cur[i] = 0;
    const uint64_t first = (kin[b0] >> begin_bit) & low_mask;
    uint64_t diff = 0;
#pragma unroll
    for (int j = 0; j < HY_ITEMS; j++) {
        const uint32_t i = tid + j * HY_THREADS;
        low[j] = i < m ? ((kin[b0 + i] >> begin_bit) & low_mask) : first;
        diff |= low[j] ^ first;
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) diff |= __shfl_xor_sync(FULL, diff, o);
    __syncthreads();
    if ((tid & 31) == 0 && diff) atomicOr(&s_diff, (unsigned long long)diff);
    __syncthreads();
    const unsigned long long dall = s_diff;
    const int hb = dall ? 63 - __clzll((long long)dall) : -1;      // highest bit that differs inside the bucket
    const int sh = hb + 1 > dbits ? hb + 1 - dbits : 0;
#pragma unroll
    for (int j = 0; j < HY_ITEMS; j++)
        if (tid + j * HY_THREADS < m) atomicAdd(&cur[(uint32_t)(low[j] >> sh) & (nbins - 1)], 1u);
    __syncthreads();
    {   // exclusive scan of the bin counts: `per` consecutive bins per thread
        uint32_t sum = 0;
        for (uint32_t j = 0; j < per; j++) sum += cur[tid * per + j];
        uint32_t tot;
        uint32_t run = block_excl_sum<uint32_t, HY_THREADS>(sum, wsum, tot);
        for (uint32_t j = 0; j < per; j++) {
            const uint32_t v = cur[tid * per + j];
            cur[tid * per + j] = run;
            run += v;
        }
    }
    __syncthreads();
#pragma unroll
    for (int j = 0; j < HY_ITEMS; j++) {
        const uint32_t i = tid + j * HY_THREADS;
        if (i < m) {
            const uint32_t p = atomicAdd(&cur[(uint32_t)(low[j] >> sh) & (nbins - 1)], 1u);
            c2[p] = (low[j] << HY_POS_BITS) | i;
        }
    }
    __syncthreads();
    // rank inside the bin (bins hold about one key), then the gathers of four keys at a time
    for (uint32_t p0 = tid; p0 < m; p0 += 4 * HY_THREADS) {
        uint32_t r[4], src[4];
#pragma unroll
        for (int u = 0; u < 4; u++) {
            const uint32_t p = p0 + u * HY_THREADS;
            r[u] = 0xffffffffu;
            src[u] = b0;
            if (p < m) {
                const uint64_t c = c2[p];
                const uint32_t d = (uint32_t)((c >> HY_POS_BITS) >> sh) & (nbins - 1);
                const uint32_t bs = d ? cur[d - 1] : 0u, be = cur[d];
                uint32_t rr = bs;
                for (uint32_t q = bs; q < be; q++) rr += c2[q] < c ? 1u : 0u;
                r[u] = b0 + rr;
                src[u] = b0 + ((uint32_t)c & (HY_CAP - 1));
            }
        }
        uint64_t kk[4];
        typename SwVal<VB>::T vv[4];
#pragma unroll
        for (int u = 0; u < 4; u++) {
            kk[u] = __ldg(kin + src[u]);
            if (VB) vv[u] = __ldg((const typename SwVal<VB>::T*)vin + src[u]);
        }
#pragma unroll
        for (int u = 0; u < 4; u++) {
            if (r[u] != 0xffffffffu) {
                kout[r[u]] = kk[u];
                if (VB) ((typename SwVal<VB>::T*)vout)[r[u]] = vv[u];
            }
        }
    }
}
constexpr size_t HY_SMEM = (size_t)HY_CAP * 8 + (size_t)HY_BINS * 4;  // 48 KB

// the buckets beyond HY_CAP keys, one block each (a persistent grid over the list the bucket pass left)
template <int VB>
__global__ void __launch_bounds__(HY_THREADS)
k_hy_pileups(uint64_t* kin, uint64_t* kout, void* vin, void* vout, int begin_bit, int low_bits,
             const uint32_t* __restrict__ starts, const uint32_t* __restrict__ pile_count,
             const uint32_t* __restrict__ pile_list) {
    __shared__ __align__(16) unsigned char smem[sizeof(SwSmem)];
    __shared__ unsigned long long s_diff;
    __shared__ uint32_t wsum[HY_THREADS / 32 + 1];
    __shared__ uint32_t base[RS_RADIX];
    const uint32_t np = *pile_count;
    for (uint32_t i = blockIdx.x; i < np; i += gridDim.x) {
        const uint32_t b = pile_list[i];
        const uint32_t b0 = __ldg(starts + b), m = __ldg(starts + b + 1) - b0;
        __syncthreads();
        hy_pileup<VB>(smem, s_diff, base, wsum, kin, kout, vin, vout, begin_bit, low_bits, b0, m);
    }
}

// A pile-up of more than HY_HUGE keys: stable LSD passes over the bits that differ inside the bucket, every pass by ALL
// blocks of a cooperative launch -- per-tile digit histograms, a prefix along the tiles of every digit row, the tile
// body of the partition pass with the offsets from the table -- with a grid barrier between the phases.  The grid is
// as large as the device holds (the launch is cooperative); without such buckets it leaves at once.
template <int VB>
__global__ void __launch_bounds__(HY_THREADS)
k_hy_huge(uint64_t* kin, uint64_t* kout, void* vin, void* vout, int begin_bit, int low_bits,
          const uint32_t* __restrict__ starts, const uint32_t* __restrict__ pile_count,
          const uint32_t* __restrict__ pile_list, uint32_t list_off, uint32_t* table /* [256][tiles] + [256] + 2 */,
          uint32_t table_tiles) {
    namespace cg = cooperative_groups;
    const uint32_t np = pile_count[1];
    if (np == 0) return;
    cg::grid_group grid = cg::this_grid();
    __shared__ SwSmem S;
    __shared__ uint32_t h[RS_RADIX];
    const int tid = threadIdx.x;
    uint32_t* totals = table + (size_t)RS_RADIX * table_tiles;
    unsigned long long* gdiff = reinterpret_cast<unsigned long long*>(totals + RS_RADIX);
    const uint64_t low_mask = low_bits >= 64 ? ~0ull : ((1ull << low_bits) - 1ull);
    for (uint32_t pi = 0; pi < np; pi++) {
        const uint32_t b = pile_list[list_off + pi];
        const uint32_t b0 = __ldg(starts + b), m = __ldg(starts + b + 1) - b0;
        const uint32_t ntiles = (m + SW_TILE - 1) / SW_TILE;
        if (ntiles > table_tiles) continue;  // (cannot happen: the table is sized for all keys)
        if (blockIdx.x == 0 && tid == 0) *gdiff = 0ull;
        grid.sync();
        {   // bits that differ inside the bucket
            const uint64_t first = kin[b0];
            uint64_t diff = 0;
            for (uint64_t i = (uint64_t)blockIdx.x * HY_THREADS + tid; i < m; i += (uint64_t)gridDim.x * HY_THREADS)
                diff |= kin[b0 + i] ^ first;
            diff = (diff >> begin_bit) & low_mask;
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) diff |= __shfl_xor_sync(FULL, diff, o);
            if ((tid & 31) == 0 && diff) atomicOr(gdiff, (unsigned long long)diff);
        }
        grid.sync();
        const unsigned long long dall = *gdiff;
        uint64_t *ka = kin + b0, *kb = kout + b0;
        char* va = VB ? (char*)vin + (size_t)b0 * VB : nullptr;
        char* vb = VB ? (char*)vout + (size_t)b0 * VB : nullptr;
        if (dall) {
            const int lo = __ffsll((long long)dall) - 1, hi = 64 - __clzll((long long)dall);
            const int npass = (hi - lo + 7) / 8;
            int at = lo;
            for (int p = 0; p < npass; p++) {
                const int bits = (hi - at + (npass - p) - 1) / (npass - p);
                const int shift = begin_bit + at;
                const uint32_t dmask = (1u << bits) - 1u;
                // per-tile digit histograms, digit-major
                for (uint32_t t = blockIdx.x; t < ntiles; t += gridDim.x) {
                    h[tid] = 0;
                    __syncthreads();
                    const uint32_t t0 = t * SW_TILE, t1 = t0 + SW_TILE < m ? t0 + SW_TILE : m;
                    for (uint32_t i = t0 + tid; i < t1; i += HY_THREADS) atomicAdd(&h[(uint32_t)(ka[i] >> shift) & dmask], 1u);
                    __syncthreads();
                    table[(size_t)tid * table_tiles + t] = h[tid];
                    __syncthreads();
                }
                grid.sync();
                // exclusive prefix along the tiles of every digit row + the row total
                for (uint32_t d = blockIdx.x; d < RS_RADIX; d += gridDim.x) {
                    uint32_t* row = table + (size_t)d * table_tiles;
                    const uint32_t per = (ntiles + HY_THREADS - 1) / HY_THREADS;
                    const uint32_t a = tid * per, e = a + per < ntiles ? a + per : ntiles;
                    uint32_t sum = 0;
                    for (uint32_t t = a; t < e; t++) sum += row[t];
                    uint32_t tot;
                    uint32_t run = block_excl_sum<uint32_t, HY_THREADS>(sum, S.wsum, tot);
                    for (uint32_t t = a; t < e; t++) {
                        const uint32_t v = row[t];
                        row[t] = run;
                        run += v;
                    }
                    if (tid == 0) totals[d] = tot;
                    __syncthreads();
                }
                grid.sync();
                // scatter
                for (uint32_t t = blockIdx.x; t < ntiles; t += gridDim.x) {
                    for (int i = tid; i < RS_WARPS * RS_RADIX; i += RS_THREADS) (&S.warp_cnt[0][0])[i] = 0;
                    __syncthreads();
                    const uint32_t t0 = t * SW_TILE;
                    const int nvalid = (int)(m - t0 < (uint32_t)SW_TILE ? m - t0 : (uint32_t)SW_TILE);
                    sw_tile<VB>(S, ka, kb, va, vb, (int64_t)t0, nvalid, shift, dmask, 0, [&](uint32_t run, uint32_t ex) {
                        uint32_t tt;
                        const uint32_t dstart = block_excl_sum<uint32_t, RS_THREADS>(totals[tid], S.wsum, tt);
                        return dstart + table[(size_t)tid * table_tiles + t] - ex;
                    });
                    __syncthreads();
                }
                grid.sync();
                uint64_t* tk = ka; ka = kb; kb = tk;
                char* tv = va; va = vb; vb = tv;
                at += bits;
            }
        }
        if (ka != kout + b0) {
            for (uint64_t i = (uint64_t)blockIdx.x * HY_THREADS + tid; i < m; i += (uint64_t)gridDim.x * HY_THREADS) {
                kout[b0 + i] = ka[i];
                if (VB == 4) ((uint32_t*)vout)[b0 + i] = ((const uint32_t*)va)[i];
                else if (VB == 8) ((uint64_t*)vout)[b0 + i] = ((const uint64_t*)va)[i];
            }
        }
        grid.sync();
    }
}

struct HyTemp {
    uint32_t *ghist, *ticket, *status, *starts, *pile_list;  // ticket[8] = number of pile-up buckets
    size_t zero_bytes;  // ghist + ticket + status: cleared before every sort
};
static void hy_temp(int64_t n, Bump& b, HyTemp& t) {
    const size_t ntiles = (size_t)cdiv(n > 0 ? n : 1, SW_TILE);
    t.ghist = b.take<uint32_t>(HY_MAX_P * RS_RADIX);
    t.ticket = b.take<uint32_t>(64);
    t.status = b.take<uint32_t>(3 * ntiles * RS_RADIX);
    t.zero_bytes = b.off;
    size_t nb = 2;
    while (nb < ((size_t)1 << HY_MAX_T) && nb * (HY_CAP / 4) < (size_t)(n > 0 ? n : 1)) nb <<= 1;
    if (nb < ((size_t)1 << HY_POS_BITS)) nb = (size_t)1 << HY_POS_BITS;  // wide keys: the plan partitions on nbits - 52 <= 12 bits
    t.starts = b.take<uint32_t>(nb + 64);
    t.pile_list = b.take<uint32_t>(2 * (nb + 64));  // [one-block pile-ups][nb + 32 entries later: the huge ones]
}

template <int VB>
static int hy_launch(const HyPlan& pl, uint64_t* k[2], void* v[2], int64_t n, int begin_bit, int end_bit,
                     const HyTemp& t, bool iota, cudaStream_t stream, int* result_buf) {
    static bool attr_set[64] = {false};  // opt in to 48 KB of dynamic shared memory once per device
    int dev = 0;
    IV_CUDA(cudaGetDevice(&dev));
    if (dev < 0 || dev >= 64 || !attr_set[dev]) {
        IV_CUDA(cudaFuncSetAttribute(k_hy_buckets<VB>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)HY_SMEM));
        if (dev >= 0 && dev < 64) attr_set[dev] = true;
    }
    IV_CUDA(cudaMemsetAsync(t.ghist, 0, t.zero_bytes, stream));
    HyDigits D;
    D.P = pl.P;
    for (int i = 0; i < HY_MAX_P; i++) { D.shift[i] = pl.shift[i]; D.mask[i] = pl.mask[i]; }
    const int64_t ntiles = cdiv(n, SW_TILE);
    const unsigned hb = (unsigned)(ntiles < 148 * 8 ? ntiles : 148 * 8);
    k_rs_ghist<<<hb, RS_THREADS, 0, stream>>>(k[0], n, D, t.ghist, ((uintptr_t)k[0] & 15) == 0);
    IV_LAUNCH_CHECK();
    int cur = 0;
    for (int p = 0; p < pl.P; p++) {
        k_rs_sweep<VB><<<(unsigned)ntiles, RS_THREADS, 0, stream>>>(
            k[cur], k[cur ^ 1], v[cur], v[cur ^ 1], n, pl.shift[p], pl.mask[p], t.ghist + p * RS_RADIX,
            t.status + (size_t)p * ntiles * RS_RADIX, t.ticket + p, (iota && p == 0) ? 1 : 0);
        IV_LAUNCH_CHECK();
        cur ^= 1;
    }
    if (pl.low_bits > 0) {
        const uint32_t nb = 1u << pl.T;
        k_hy_starts<<<nb / 256 + 1, 256, 0, stream>>>(k[cur], n, end_bit - pl.T, nb - 1u, nb, t.starts);
        IV_LAUNCH_CHECK();
        k_hy_buckets<VB><<<nb, HY_THREADS, HY_SMEM, stream>>>(k[cur], k[cur ^ 1], v[cur], v[cur ^ 1], begin_bit,
                                                              pl.low_bits, t.starts, t.ticket + 8, t.pile_list);
        IV_LAUNCH_CHECK();
        k_hy_pileups<VB><<<148 * 2, HY_THREADS, 0, stream>>>(k[cur], k[cur ^ 1], v[cur], v[cur ^ 1], begin_bit, pl.low_bits,
                                                             t.starts, t.ticket + 8, t.pile_list);
        IV_LAUNCH_CHECK();
        {   // the huge ones: one cooperative launch (it needs every block resident: sized by the occupancy query)
            static int coop_blocks[64] = {0};
            if (dev < 0 || dev >= 64 || coop_blocks[dev] == 0) {
                int sms = 0, per_sm = 0, ok = 0;
                IV_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev < 0 ? 0 : dev));
                IV_CUDA(cudaDeviceGetAttribute(&ok, cudaDevAttrCooperativeLaunch, dev < 0 ? 0 : dev));
                IV_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, k_hy_huge<VB>, HY_THREADS, 0));
                IV_REQUIRE(ok && sms * per_sm > 0, IV_ERR_CUDA, "radix sort: the device does not support cooperative launches");
                if (dev >= 0 && dev < 64) coop_blocks[dev] = sms * per_sm;
            }
            const int blocks = (dev >= 0 && dev < 64) ? coop_blocks[dev] : 148;
            uint64_t *a0 = k[cur], *a1 = k[cur ^ 1];
            void *a2 = v[cur], *a3 = v[cur ^ 1];
            int a4 = begin_bit, a5 = pl.low_bits;
            const uint32_t *a6 = t.starts, *a7 = t.ticket + 8, *a8 = t.pile_list;
            uint32_t a9 = nb + 32;  // the second list starts gridDim.x (of k_hy_buckets) + 32 entries in
            uint32_t* a10 = t.status;
            uint32_t a11 = (uint32_t)ntiles;
            void* args[] = {&a0, &a1, &a2, &a3, &a4, &a5, &a6, &a7, &a8, &a9, &a10, &a11};
            IV_CUDA(cudaLaunchCooperativeKernel((const void*)k_hy_huge<VB>, dim3((unsigned)blocks), dim3(HY_THREADS), args, 0,
                                                stream));
            __atomic_fetch_add(&iv::g_launches, 1ull, __ATOMIC_RELAXED);
        }
        cur ^= 1;
    }
    *result_buf = cur;
    return IV_OK;
}

// ---- host driver ---------------------------------------------------------------------------------------------
static int items_for(int64_t n) { return n <= ((int64_t)1 << 20) ? RS_ITEMS_SMALL : RS_ITEMS_BIG; }
// (11-bit digits -- 3 passes instead of 4 for 32-bit keys -- were measured for small inputs: the 2048-entry
// histogram tables made each pass twice as long, 125 us instead of 86 us for the 200k-subject index.  A bitonic
// network for the buckets of the small-input path took 61 us, the shared-memory LSD passes above 31 us.)
int radix_sort_passes(int64_t n, int nbits) {
    if (n <= 1 || nbits <= 0) return 0;
    const HyPlan pl = hy_plan(n, 0, nbits);
    if (pl.use) return pl.P + (pl.low_bits > 0 ? 1 : 0);
    return bucket_path(n, nbits) ? 2 : (nbits + 7) / 8;
}

size_t radix_sort_temp_bytes(int64_t n) {
    int64_t ntiles = cdiv(n > 0 ? n : 1, (int64_t)RS_THREADS * items_for(n));
    size_t hist = (size_t)ntiles * RS_RADIX * RS_MAX_JOBS;
    size_t bytes = (hist + RS_RADIX * RS_MAX_JOBS + 128) * sizeof(uint32_t) + 512;
    if (n > BK_MAX_N && n <= HY_MAX_N) {
        Bump b(nullptr, 0);
        HyTemp t;
        hy_temp(n, b, t);
        if (b.off + 256 > bytes) bytes = b.off + 256;
    }
    return bytes;
}

template <int ITEMS>
static int radix_sort_multi_t(const SortJob* jobs, int njobs, int64_t n, int begin_bit, int end_bit, void* temp,
                              cudaStream_t stream, int* result_buf) {
    constexpr int TILE = RS_THREADS * ITEMS;
    const int orig_begin = begin_bit;
    int64_t ntiles = cdiv(n, TILE);
    uint32_t* hist = (uint32_t*)temp;
    int64_t nh = ntiles * RS_RADIX * njobs;
    uint32_t* totals = hist + ((nh + 63) & ~(int64_t)63);
    const int nrows = RS_RADIX * njobs;
    int cur = 0;
    const bool buckets = bucket_path(n, end_bit - begin_bit);
    if (buckets) begin_bit = end_bit - 8;  // the partition pass: top digit only
    for (int shift = begin_bit; shift < end_bit; shift += 8) {
        int bits = end_bit - shift < 8 ? end_bit - shift : 8;
        uint32_t mask = (1u << bits) - 1u;
        SortJobs J;
        for (int j = 0; j < njobs; j++) {
            J.kin[j] = cur ? jobs[j].k1 : jobs[j].k0;
            J.kout[j] = cur ? jobs[j].k0 : jobs[j].k1;
            J.vin[j] = cur ? jobs[j].v1 : jobs[j].v0;
            J.vout[j] = cur ? jobs[j].v0 : jobs[j].v1;
            J.vb[j] = jobs[j].vb;
        }
        dim3 grid((unsigned)ntiles, (unsigned)njobs);
        k_rs_hist<ITEMS><<<grid, RS_THREADS, 0, stream>>>(J, n, shift, mask, hist, ntiles);
        IV_LAUNCH_CHECK();
        if (ntiles >= 2048) k_rs_prefix_long<<<(unsigned)nrows, RS_THREADS, 0, stream>>>(hist, ntiles, totals);
        else k_rs_prefix<<<(unsigned)cdiv(nrows, RS_WARPS), RS_THREADS, 0, stream>>>(hist, ntiles, totals, nrows);
        IV_LAUNCH_CHECK();
        k_rs_scatter<ITEMS><<<grid, RS_THREADS, 0, stream>>>(J, n, shift, mask, hist, totals, ntiles);
        IV_LAUNCH_CHECK();
        cur ^= 1;
    }
    if (buckets) {
        static bool attr_set[64] = {false};  // opt in to 73 KB of dynamic shared memory once per device
        int dev = 0;
        IV_CUDA(cudaGetDevice(&dev));
        if (dev < 0 || dev >= 64 || !attr_set[dev]) {
            IV_CUDA(cudaFuncSetAttribute(k_rs_buckets, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)BK_SMEM));
            if (dev >= 0 && dev < 64) attr_set[dev] = true;
        }
        SortJobs J;
        for (int j = 0; j < njobs; j++) {
            J.kin[j] = jobs[j].k1; J.kout[j] = jobs[j].k0;
            J.vin[j] = jobs[j].v1; J.vout[j] = jobs[j].v0;
            J.vb[j] = jobs[j].vb;
        }
        k_rs_buckets<<<dim3(RS_RADIX, (unsigned)njobs), BK_THREADS, BK_SMEM, stream>>>(J, n, orig_begin,
                                                                                      end_bit - 8 - orig_begin, totals);
        IV_LAUNCH_CHECK();
        cur = 0;
    }
    *result_buf = cur;
    return IV_OK;
}

int radix_sort_multi(const SortJob* jobs, int njobs, int64_t n, int begin_bit, int end_bit, void* temp,
                     size_t temp_bytes, cudaStream_t stream, int* result_buf) {
    *result_buf = 0;
    if (n <= 1 || end_bit <= begin_bit || njobs <= 0) return IV_OK;
    IV_REQUIRE(njobs <= RS_MAX_JOBS, IV_ERR_INVALID, "radix sort: too many jobs");
    IV_REQUIRE((int64_t)njobs * n < ((int64_t)1 << 32) - 4096, IV_ERR_RANGE, "radix sort: n=%lld exceeds 2^32",
               (long long)n);
    for (int j = 0; j < njobs; j++)
        IV_REQUIRE(jobs[j].vb == 0 || jobs[j].vb == 4 || jobs[j].vb == 8, IV_ERR_INVALID, "radix sort: value_bytes");
    IV_REQUIRE(temp_bytes >= radix_sort_temp_bytes(n), IV_ERR_WORKSPACE, "radix sort: temp too small");
    if (njobs == 1) {
        const HyPlan pl = hy_plan(n, begin_bit, end_bit);
        if (pl.use) {
            Bump b(temp, temp_bytes);
            HyTemp t;
            hy_temp(n, b, t);
            uint64_t* k[2] = {jobs[0].k0, jobs[0].k1};
            void* v[2] = {jobs[0].v0, jobs[0].v1};
            if (jobs[0].vb == 0) return hy_launch<0>(pl, k, v, n, begin_bit, end_bit, t, false, stream, result_buf);
            if (jobs[0].vb == 4) return hy_launch<4>(pl, k, v, n, begin_bit, end_bit, t, false, stream, result_buf);
            return hy_launch<8>(pl, k, v, n, begin_bit, end_bit, t, false, stream, result_buf);
        }
    }
    if (items_for(n) == RS_ITEMS_SMALL)
        return radix_sort_multi_t<RS_ITEMS_SMALL>(jobs, njobs, n, begin_bit, end_bit, temp, stream, result_buf);
    return radix_sort_multi_t<RS_ITEMS_BIG>(jobs, njobs, n, begin_bit, end_bit, temp, stream, result_buf);
}

int radix_sort_pingpong(uint64_t* k0, uint64_t* k1, void* v0, void* v1, int value_bytes, int64_t n, int begin_bit,
                        int end_bit, void* temp, size_t temp_bytes, cudaStream_t stream, int* result_buf) {
    SortJob j{k0, k1, v0, v1, value_bytes};
    return radix_sort_multi(&j, 1, n, begin_bit, end_bit, temp, temp_bytes, stream, result_buf);
}

}  // namespace iv

// ---- exported test entry ----------------------------------------------------------------------------------------
extern "C" size_t iv_sort_workspace_bytes(int64_t n, int32_t value_bytes) {
    size_t nn = (size_t)(n > 0 ? n : 1);
    size_t b = ((nn * 8 + 255) & ~(size_t)255) + ((nn * (size_t)value_bytes + 255) & ~(size_t)255);
    return b + iv::radix_sort_temp_bytes(n) + 1024;
}

extern "C" int iv_radix_sort_u64(uint64_t* keys, void* values, int32_t value_bytes, int64_t n, int32_t begin_bit,
                                 int32_t end_bit, void* ws, size_t ws_bytes, void* stream) {
    using namespace iv;
    cudaStream_t st = (cudaStream_t)stream;
    IV_REQUIRE(n >= 0 && begin_bit >= 0 && end_bit <= 64, IV_ERR_INVALID, "iv_radix_sort_u64: bad arguments");
    IV_REQUIRE(ws_bytes >= iv_sort_workspace_bytes(n, value_bytes), IV_ERR_WORKSPACE, "iv_radix_sort_u64: workspace");
    if (n <= 1) return IV_OK;
    Bump b(ws, ws_bytes);
    uint64_t* k1 = b.take<uint64_t>((size_t)n);
    char* v1 = value_bytes ? b.take<char>((size_t)n * value_bytes) : nullptr;
    void* temp = b.take<char>(radix_sort_temp_bytes(n));
    int res = 0;
    int rc = radix_sort_pingpong(keys, k1, values, v1, value_bytes, n, begin_bit, end_bit, temp,
                                 radix_sort_temp_bytes(n), st, &res);
    if (rc != IV_OK) return rc;
    if (res == 1) {
        IV_CUDA(cudaMemcpyAsync(keys, k1, (size_t)n * 8, cudaMemcpyDeviceToDevice, st));
        if (value_bytes) IV_CUDA(cudaMemcpyAsync(values, v1, (size_t)n * value_bytes, cudaMemcpyDeviceToDevice, st));
    }
    return IV_OK;
}
