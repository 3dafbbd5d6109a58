// materialize.cu -- device-side materialisation of the joined frame.
//
// Replaces, for the columns that already live in HBM, the two `reindex` gathers of the reference's join tail
//   scdf.reindex(_self_indexes) / ocdf.reindex(_other_indexes)          pyranges/methods/join.py:95-123
//   df["Overlap"] = min(End, End_b) - max(Start, Start_b)               pyranges/methods/join.py:147-151
// and the clip of intersect                                              pyranges/methods/intersection.py:39-48
//
// k_materialize   one pair per thread: (query row, subject row) -> Start, End, Start_b, End_b, Overlap.  The pair
//                 lists are in query-row order, so the query gathers walk the query columns forward (sector
//                 reuse in L1/L2) and the subject gathers hit the small subject table in L2.  88 algorithmic bytes
//                 per pair with all five outputs: HBM-bound.
// k_gather<T>     out[i] = column[rows[i]] for a payload column of 1/2/4/8-byte elements; rows < 0 (the "-1" filler of
//                 outer joins) produce `fill`.
#include "common.cuh"

namespace iv {

constexpr int MT_THREADS = 256;

__global__ void __launch_bounds__(MT_THREADS)
k_materialize(const int64_t* __restrict__ q_start, const int64_t* __restrict__ q_end,
              const int64_t* __restrict__ s_start, const int64_t* __restrict__ s_end,
              const int64_t* __restrict__ pair_q, const int64_t* __restrict__ pair_s, int64_t n, int64_t q_off,
              int64_t s_off, int64_t* __restrict__ o_qs, int64_t* __restrict__ o_qe, int64_t* __restrict__ o_ss,
              int64_t* __restrict__ o_se, int64_t* __restrict__ o_ov, int32_t clip) {
    const int64_t i = (int64_t)blockIdx.x * MT_THREADS + threadIdx.x;
    if (i >= n) return;
    const int64_t qr = __ldg(pair_q + i) - q_off, sr = __ldg(pair_s + i) - s_off;
    int64_t qs = __ldg(q_start + qr), qe = __ldg(q_end + qr);
    const int64_t ss = __ldg(s_start + sr), se = __ldg(s_end + sr);
    if (o_ov) o_ov[i] = min(qe, se) - max(qs, ss);
    if (clip) {  // intersect: the query row clipped to the overlap
        qs = max(qs, ss);
        qe = min(qe, se);
    }
    if (o_qs) o_qs[i] = qs;
    if (o_qe) o_qe[i] = qe;
    if (o_ss) o_ss[i] = ss;
    if (o_se) o_se[i] = se;
}

template <typename T>
__global__ void __launch_bounds__(MT_THREADS)
k_gather(const T* __restrict__ col, const int64_t* __restrict__ rows, int64_t n, T* __restrict__ out, T fill) {
    const int64_t i = (int64_t)blockIdx.x * MT_THREADS + threadIdx.x;
    if (i >= n) return;
    const int64_t r = __ldg(rows + i);
    out[i] = r >= 0 ? __ldg(col + r) : fill;
}

}  // namespace iv

using namespace iv;

extern "C" int iv_materialize_pairs(const int64_t* q_start, const int64_t* q_end, const int64_t* s_start,
                                    const int64_t* s_end, const int64_t* pair_q, const int64_t* pair_s,
                                    int64_t n_pairs, int64_t q_row_offset, int64_t s_row_offset, int32_t clip,
                                    int64_t* out_q_start, int64_t* out_q_end, int64_t* out_s_start,
                                    int64_t* out_s_end, int64_t* out_overlap, void* stream) {
    cudaStream_t st = (cudaStream_t)stream;
    IV_REQUIRE(n_pairs >= 0, IV_ERR_INVALID, "iv_materialize_pairs: negative pair count");
    if (n_pairs == 0) return IV_OK;
    IV_REQUIRE(q_start && q_end && s_start && s_end && pair_q && pair_s, IV_ERR_INVALID,
               "iv_materialize_pairs: null input column");
    k_materialize<<<(unsigned)cdiv(n_pairs, MT_THREADS), MT_THREADS, 0, st>>>(
        q_start, q_end, s_start, s_end, pair_q, pair_s, n_pairs, q_row_offset, s_row_offset, out_q_start, out_q_end,
        out_s_start, out_s_end, out_overlap, clip);
    IV_LAUNCH_CHECK();
    return IV_OK;
}

extern "C" int iv_gather(const void* column, int32_t elem_bytes, const int64_t* rows, int64_t n, void* out,
                         const void* fill, void* stream) {
    cudaStream_t st = (cudaStream_t)stream;
    IV_REQUIRE(n >= 0, IV_ERR_INVALID, "iv_gather: negative row count");
    if (n == 0) return IV_OK;
    IV_REQUIRE(column && rows && out, IV_ERR_INVALID, "iv_gather: null argument");
    IV_REQUIRE(elem_bytes == 1 || elem_bytes == 2 || elem_bytes == 4 || elem_bytes == 8, IV_ERR_INVALID,
               "iv_gather: element size %d is not 1, 2, 4 or 8", (int)elem_bytes);
    const unsigned nb = (unsigned)cdiv(n, MT_THREADS);
    uint64_t f = 0;
    if (fill) memcpy(&f, fill, (size_t)elem_bytes);  // little-endian: the low bytes are the element
    switch (elem_bytes) {
        case 1: k_gather<uint8_t><<<nb, MT_THREADS, 0, st>>>((const uint8_t*)column, rows, n, (uint8_t*)out, (uint8_t)f); break;
        case 2: k_gather<uint16_t><<<nb, MT_THREADS, 0, st>>>((const uint16_t*)column, rows, n, (uint16_t*)out, (uint16_t)f); break;
        case 4: k_gather<uint32_t><<<nb, MT_THREADS, 0, st>>>((const uint32_t*)column, rows, n, (uint32_t*)out, (uint32_t)f); break;
        default: k_gather<uint64_t><<<nb, MT_THREADS, 0, st>>>((const uint64_t*)column, rows, n, (uint64_t*)out, f); break;
    }
    IV_LAUNCH_CHECK();
    return IV_OK;
}
