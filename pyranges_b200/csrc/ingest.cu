// ingest.cu -- BED text straight into device columns.
//
// Replaces the parse of pyranges.read_bed (pyranges/readers.py:63-153: pandas.read_csv(sep="\t", header=None|0) +
// column names) for the columns the interval path works on: the file's bytes are copied to the device once, lines
// and tab-separated fields are found there, Start / End / Score are parsed as int64, the chromosome field is hashed
// (the dictionary of the few distinct names is built by the caller from one representative span per hash), Strand is
// one byte.  Free-text fields (Name, ItemRGB, BlockSizes ...) are returned as (offset, length) spans; iv_text_gather
// packs the spans of one column into a contiguous buffer for the host.
//
//   iv_text_count_lines   newlines per 16 KB tile -> exclusive scan, number of lines (a last line without '\n' counts)
//   iv_text_line_starts   line_start[i] = offset of the first byte of line i; line_start[n] = one past the (virtual)
//                         terminator of the last line, so that line i is [line_start[i], line_start[i+1] - 1)
//   iv_bed_parse          one thread per line
//   iv_text_gather        spans -> packed bytes + offsets
#include "common.cuh"

namespace iv {

constexpr int TX_THREADS = 256;
constexpr int TX_PER_THREAD = 64;                       // bytes per thread
constexpr int TX_TILE = TX_THREADS * TX_PER_THREAD;     // 16 KB per block

__device__ __forceinline__ int nl_count16(uint4 v) {
    int c = 0;
    const uint32_t w[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
    for (int i = 0; i < 4; i++) {
        const uint32_t x = w[i] ^ 0x0a0a0a0au;  // zero bytes where the text has '\n'
        const uint32_t z = ~(((x & 0x7f7f7f7fu) + 0x7f7f7f7fu) | x | 0x7f7f7f7fu);  // 0x80 in every zero byte
        c += __popc(z);
    }
    return c;
}

// newlines in the thread's 64 bytes (the buffer is padded with zeros to a multiple of 16 bytes past nbytes)
__device__ __forceinline__ int thread_newlines(const uint8_t* __restrict__ buf, int64_t nbytes, int64_t t0) {
    int c = 0;
    if (t0 + TX_PER_THREAD <= nbytes && (((uintptr_t)(buf + t0)) & 15) == 0) {
        const uint4* p = reinterpret_cast<const uint4*>(buf + t0);
#pragma unroll
        for (int i = 0; i < TX_PER_THREAD / 16; i++) c += nl_count16(__ldg(p + i));
    } else {
        for (int64_t i = t0; i < t0 + TX_PER_THREAD && i < nbytes; i++) c += __ldg(buf + i) == '\n';
    }
    return c;
}

__global__ void __launch_bounds__(TX_THREADS)
k_tx_count(const uint8_t* __restrict__ buf, int64_t nbytes, int64_t* __restrict__ tile_cnt) {
    __shared__ int64_t sm[TX_THREADS / 32 + 1];
    const int64_t t0 = (int64_t)blockIdx.x * TX_TILE + (int64_t)threadIdx.x * TX_PER_THREAD;
    int64_t tot;
    block_excl_sum<int64_t, TX_THREADS>((int64_t)thread_newlines(buf, nbytes, t0), sm, tot);
    if (threadIdx.x == 0) tile_cnt[blockIdx.x] = tot;
}

__global__ void __launch_bounds__(TX_THREADS)
k_tx_starts(const uint8_t* __restrict__ buf, int64_t nbytes, const int64_t* __restrict__ tile_off, int64_t n_lines,
            int64_t* __restrict__ line_start) {
    __shared__ int64_t sm[TX_THREADS / 32 + 1];
    const int64_t t0 = (int64_t)blockIdx.x * TX_TILE + (int64_t)threadIdx.x * TX_PER_THREAD;
    int64_t tot;
    int64_t r = block_excl_sum<int64_t, TX_THREADS>((int64_t)thread_newlines(buf, nbytes, t0), sm, tot) +
                __ldg(tile_off + blockIdx.x);
    for (int64_t i = t0; i < t0 + TX_PER_THREAD && i < nbytes; i++)
        if (__ldg(buf + i) == '\n') line_start[++r] = i + 1;  // the line after newline number r starts here
    if (blockIdx.x == 0 && threadIdx.x == 0) {
        line_start[0] = 0;
        // a last line without terminator: its virtual newline sits at nbytes
        if (nbytes > 0 && __ldg(buf + nbytes - 1) != '\n') line_start[n_lines] = nbytes + 1;
    }
}

// FNV-1a over a field
__device__ __forceinline__ uint64_t fnv1a(const uint8_t* __restrict__ p, int n) {
    uint64_t h = 1469598103934665603ull;
    for (int i = 0; i < n; i++) h = (h ^ __ldg(p + i)) * 1099511628211ull;
    return h;
}
// decimal integer with optional sign; ok = the whole field is one
__device__ __forceinline__ int64_t parse_i64(const uint8_t* __restrict__ p, int n, bool& ok) {
    int i = 0;
    bool neg = false;
    if (n > 0 && (__ldg(p) == '-' || __ldg(p) == '+')) { neg = __ldg(p) == '-'; i = 1; }
    ok = i < n && n - i <= 18;
    int64_t v = 0;
    for (; i < n; i++) {
        const int d = (int)__ldg(p + i) - '0';
        ok = ok && d >= 0 && d <= 9;
        v = v * 10 + d;
    }
    return neg ? -v : v;
}

constexpr int BED_MAX_FIELDS = 12;

__global__ void __launch_bounds__(256)
k_bed_parse(const uint8_t* __restrict__ buf, const int64_t* __restrict__ line_start, int64_t first_line, int64_t n_rows,
            int64_t* __restrict__ start, int64_t* __restrict__ end, int64_t* __restrict__ score,
            uint64_t* __restrict__ chrom_hash, uint8_t* __restrict__ strand, uint8_t* __restrict__ n_fields,
            int64_t* __restrict__ field_off, int32_t* __restrict__ field_len, int32_t* __restrict__ bad) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n_rows) return;
    const int64_t a = __ldg(line_start + first_line + i);
    int64_t b = __ldg(line_start + first_line + i + 1) - 1;  // the newline
    if (b > a && __ldg(buf + b - 1) == '\r') b--;
    int64_t fo[BED_MAX_FIELDS];
    int fl[BED_MAX_FIELDS];
    int nf = 0;
    int64_t f0 = a;
    for (int64_t p = a; p <= b; p++) {
        if (p == b || __ldg(buf + p) == '\t') {
            if (nf < BED_MAX_FIELDS) { fo[nf] = f0; fl[nf] = (int)(p - f0); }
            nf++;
            f0 = p + 1;
        }
    }
    if (b == a) nf = 0;  // empty line
    const int kept = nf < BED_MAX_FIELDS ? nf : BED_MAX_FIELDS;
    n_fields[i] = (uint8_t)(nf < 255 ? nf : 255);
    for (int f = 0; f < BED_MAX_FIELDS; f++) {
        field_off[(int64_t)f * n_rows + i] = f < kept ? fo[f] : 0;
        field_len[(int64_t)f * n_rows + i] = f < kept ? fl[f] : 0;
    }
    bool ok_s = false, ok_e = false, ok_sc = true;
    int64_t s = 0, e = 0, sc = 0;
    if (kept >= 3) {
        s = parse_i64(buf + fo[1], fl[1], ok_s);
        e = parse_i64(buf + fo[2], fl[2], ok_e);
    }
    if (kept >= 5) sc = parse_i64(buf + fo[4], fl[4], ok_sc);
    start[i] = s;
    end[i] = e;
    score[i] = sc;
    chrom_hash[i] = kept >= 1 ? fnv1a(buf + fo[0], fl[0]) : 0ull;
    strand[i] = (kept >= 6 && fl[5] == 1) ? __ldg(buf + fo[5]) : (uint8_t)0;
    // bad[0]: rows whose Start / End are not integers (or fewer than three fields); bad[1]: non-integer Score fields;
    // bad[2]: Strand fields that are not one character
    if (!(ok_s && ok_e)) atomicAdd(bad + 0, 1);
    if (!ok_sc) atomicAdd(bad + 1, 1);
    if (kept >= 6 && fl[5] != 1) atomicAdd(bad + 2, 1);
}

// out_off[i] = exclusive sum of len (computed by the caller's scan); copies the bytes of every span
__global__ void __launch_bounds__(256)
k_tx_gather(const uint8_t* __restrict__ buf, const int64_t* __restrict__ off, const int32_t* __restrict__ len, int64_t n,
            const int64_t* __restrict__ out_off, uint8_t* __restrict__ out) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const int64_t a = __ldg(off + i), o = __ldg(out_off + i);
    const int l = __ldg(len + i);
    for (int k = 0; k < l; k++) out[o + k] = __ldg(buf + a + k);
}
__global__ void __launch_bounds__(256)
k_tx_len64(const int32_t* __restrict__ len, int64_t n, int64_t* __restrict__ out) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) out[i] = __ldg(len + i);
}

}  // namespace iv

using namespace iv;

extern "C" size_t iv_text_workspace_bytes(int64_t nbytes) {
    const size_t nt = (size_t)cdiv(nbytes > 0 ? nbytes : 1, TX_TILE) + 2;
    return (nt + scan_i64_large_temp_elems((int64_t)nt) + 64) * sizeof(int64_t) + 1024;
}

// number of lines of the text (device, *out_n_lines); leaves the per-tile line offsets in ws for iv_text_line_starts
extern "C" int iv_text_count_lines(const uint8_t* buf, int64_t nbytes, void* ws, size_t ws_bytes, int64_t* out_n_lines,
                                   void* stream) {
    cudaStream_t st = (cudaStream_t)stream;
    IV_REQUIRE(buf && nbytes >= 0 && ws && out_n_lines, IV_ERR_INVALID, "iv_text_count_lines: null argument");
    IV_REQUIRE(ws_bytes >= iv_text_workspace_bytes(nbytes), IV_ERR_WORKSPACE, "iv_text_count_lines: workspace");
    const int64_t nt = cdiv(nbytes > 0 ? nbytes : 1, TX_TILE);
    Bump b(ws, ws_bytes);
    int64_t* tile = b.take<int64_t>((size_t)nt + 2);
    int64_t* tmp = b.take<int64_t>(scan_i64_large_temp_elems(nt));
    k_tx_count<<<(unsigned)nt, TX_THREADS, 0, st>>>(buf, nbytes, tile);
    IV_LAUNCH_CHECK();
    int rc = scan_exclusive_i64_large(tile, nt, tmp, out_n_lines, st);  // *out_n_lines = newlines
    if (rc != IV_OK) return rc;
    return IV_OK;
}

// line_start[n_lines + 1]; n_lines = newlines + (the text does not end in a newline ? 1 : 0), computed by the caller from
// iv_text_count_lines' result and the last byte
extern "C" int iv_text_line_starts(const uint8_t* buf, int64_t nbytes, const void* ws, int64_t n_lines,
                                   int64_t* out_line_start, void* stream) {
    cudaStream_t st = (cudaStream_t)stream;
    IV_REQUIRE(buf && ws && out_line_start && n_lines >= 0, IV_ERR_INVALID, "iv_text_line_starts: null argument");
    const int64_t nt = cdiv(nbytes > 0 ? nbytes : 1, TX_TILE);
    k_tx_starts<<<(unsigned)nt, TX_THREADS, 0, st>>>(buf, nbytes, (const int64_t*)ws, n_lines, out_line_start);
    IV_LAUNCH_CHECK();
    return IV_OK;
}

// rows = lines [first_line, first_line + n_rows).  field_off / field_len: [12][n_rows] spans of the tab-separated fields
// (0 / 0 beyond the line's last field); out_bad[3] (device, zeroed by the call): rows with a non-integer Start / End,
// non-integer Score fields, Strand fields that are not one character.
extern "C" int iv_bed_parse(const uint8_t* buf, const int64_t* line_start, int64_t first_line, int64_t n_rows,
                            int64_t* out_start, int64_t* out_end, int64_t* out_score, uint64_t* out_chrom_hash,
                            uint8_t* out_strand, uint8_t* out_n_fields, int64_t* out_field_off, int32_t* out_field_len,
                            int32_t* out_bad, void* stream) {
    cudaStream_t st = (cudaStream_t)stream;
    IV_REQUIRE(buf && line_start && out_bad, IV_ERR_INVALID, "iv_bed_parse: null argument");
    IV_CUDA(cudaMemsetAsync(out_bad, 0, 3 * sizeof(int32_t), st));
    if (n_rows <= 0) return IV_OK;
    IV_REQUIRE(out_start && out_end && out_score && out_chrom_hash && out_strand && out_n_fields && out_field_off &&
                   out_field_len, IV_ERR_INVALID, "iv_bed_parse: null output");
    k_bed_parse<<<(unsigned)cdiv(n_rows, 256), 256, 0, st>>>(buf, line_start, first_line, n_rows, out_start, out_end,
                                                            out_score, out_chrom_hash, out_strand, out_n_fields,
                                                            out_field_off, out_field_len, out_bad);
    IV_LAUNCH_CHECK();
    return IV_OK;
}

// packs n spans: out_off[n + 1] (exclusive sums of the lengths, total in out_off[n]) and, when out_bytes is given, the
// bytes.  Two calls: out_bytes = NULL to learn the total, then with a buffer of out_off[n] bytes.
extern "C" int iv_text_gather(const uint8_t* buf, const int64_t* off, const int32_t* len, int64_t n, int64_t* out_off,
                              uint8_t* out_bytes, void* ws, size_t ws_bytes, void* stream) {
    cudaStream_t st = (cudaStream_t)stream;
    IV_REQUIRE(buf && out_off && n >= 0, IV_ERR_INVALID, "iv_text_gather: null argument");
    if (n == 0) {
        IV_CUDA(cudaMemsetAsync(out_off, 0, 8, st));
        return IV_OK;
    }
    IV_REQUIRE(off && len, IV_ERR_INVALID, "iv_text_gather: null argument");
    if (!out_bytes) {
        IV_REQUIRE(ws && ws_bytes >= (scan_i64_large_temp_elems(n) + 8) * sizeof(int64_t), IV_ERR_WORKSPACE,
                   "iv_text_gather: workspace");
        k_tx_len64<<<(unsigned)cdiv(n, 256), 256, 0, st>>>(len, n, out_off);
        IV_LAUNCH_CHECK();
        return scan_exclusive_i64_large(out_off, n, (int64_t*)ws, out_off + n, st);
    }
    k_tx_gather<<<(unsigned)cdiv(n, 256), 256, 0, st>>>(buf, off, len, n, out_off, out_bytes);
    IV_LAUNCH_CHECK();
    return IV_OK;
}
