/*
 * pyranges_b200.h -- C ABI of the B200-native interval engine (libpyranges_b200.so).
 *
 * Drop-in boundary B2 of SURVEY.md section 8(b): these entry points replace the numpy-in/numpy-out
 * calls pyranges v0 makes into its three native dependencies (ncls, sorted_nearest, pyrle; none of
 * them is in the reference tree, the cited file:line are the reference's CALL SITES):
 *
 *   iv_index_build            <- NCLS(starts, ends, ids)                 pyranges/methods/join.py:12,
 *                                                                        intersection.py:19,71, coverage.py:20,58
 *   iv_count_overlaps         <- NCLS.all_overlaps_both + value_counts   pyranges/methods/coverage.py:26-40
 *                                NCLS.has_overlaps (mode ANY)            pyranges/methods/intersection.py:78
 *   iv_emit_pairs             <- NCLS.all_overlaps_both                  pyranges/methods/join.py:15,23
 *                                NCLS.all_overlaps_self                  pyranges/methods/intersection.py:74
 *                                NCLS.all_containments_both              pyranges/methods/join.py:17, intersection.py:76
 *                                NCLS.first_overlap_both                 pyranges/methods/join.py:19 (via nearest.py:32)
 *                                NCLS.last_overlap_both                  pyranges/methods/join.py:21, intersection.py:28
 *   iv_lists_build            <- NCLS(starts, ends, ids)                 pyranges/methods/join.py:12, coverage.py:20
 *   iv_join_pairs             <- NCLS.all_overlaps_both                  pyranges/methods/join.py:15,23 (one kernel)
 *   iv_join_count             <- NCLS.all_overlaps_both + value_counts   pyranges/methods/coverage.py:26-40
 *                                NCLS.has_overlaps                       pyranges/methods/intersection.py:78
 *   iv_coverage_bp            <- NCLS.coverage                           pyranges/methods/coverage.py:64-68
 *   iv_subtract_count/_emit   <- NCLS.set_difference_helper              pyranges/methods/subtraction.py:63-65
 *   iv_nearest                <- first_overlap_both + nearest_previous_nonoverlapping +
 *                                nearest_next_nonoverlapping + nearest_nonoverlapping
 *                                                                        pyranges/methods/nearest.py:29-53,59,69,113
 *   iv_k_nearest_count/_emit  <- k_nearest_previous_nonoverlapping + k_nearest_next_nonoverlapping
 *                                                                        pyranges/methods/k_nearest.py:6-48
 *   iv_cluster                <- sort_values("Start") + annotate_clusters pyranges/methods/cluster.py:11-15,
 *                                + global id offsets                      pyranges/pyranges_main.py:1210-1223
 *   iv_merge_count/_emit      <- sort_values("Start") + find_clusters     pyranges/methods/merge.py:12-14
 *   iv_rle_count/_emit        <- pyrle.methods.coverage                   pyranges/methods/to_rle.py:18
 *   iv_split_count/_emit      <- concat(Start, End).sort_values().drop_duplicates() + consecutive pairs
 *                                                                        pyranges/methods/split.py:4-24
 *   iv_text_* / iv_bed_parse  <- pandas.read_csv(f, sep="\t") of read_bed      pyranges/readers.py:63-153
 *   iv_materialize_pairs      <- scdf.reindex / ocdf.reindex + "Overlap"  pyranges/methods/join.py:95-123,147-151
 *                                and the clip of intersect                pyranges/methods/intersection.py:39-48
 *   iv_gather                 <- the same reindex for a payload column    pyranges/methods/join.py:95-123
 *
 * Where the reference loops over (chromosome, strand) keys in Python and calls the native library
 * once per key (pyranges/multithreaded.py:182-303, :306-372), every entry point here is BATCHED over
 * all keys: rows of all keys are concatenated (key-segmented, natsorted key order) and one launch
 * sequence handles every key.  Groups ("partner frames") are contiguous row ranges of the subject
 * columns; each query segment names its partner group (or -1 = no partner).
 *
 * Conventions
 *   - every pointer is a DEVICE pointer unless the name ends in _host or the type is a struct below
 *     (structs live in host memory, their members are device pointers / scalars);
 *   - all coordinates, labels, counts are int64 (pyranges/methods/init.py:13-16);
 *   - the caller owns every buffer including the workspace (size from *_workspace_bytes); the library
 *     never allocates or frees device memory and does not synchronise the stream (one exception:
 *     iv_cluster / iv_merge_count read back one flag to skip the sort of already sorted input);
 *   - `stream` is a cudaStream_t passed as void*;
 *   - return value: IV_OK or a negative error; iv_last_error_string() describes the last failure of
 *     the calling thread.
 *   - intervals are half open [start, end); overlap iff S < e && s < E.  start <= end is required.
 */
#ifndef PYRANGES_B200_H
#define PYRANGES_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define IV_ABI_VERSION 8

#define IV_OK 0
#define IV_ERR_INVALID (-1)   /* bad argument */
#define IV_ERR_CUDA (-2)      /* CUDA runtime error (see iv_last_error_string) */
#define IV_ERR_WORKSPACE (-3) /* workspace too small */
#define IV_ERR_RANGE (-4)     /* coordinate span / row count exceeds what the flattened keys can hold */

/* emit / count modes */
#define IV_MODE_ALL 0     /* every overlapping (query, subject) pair            all_overlaps_both      */
#define IV_MODE_ANY 1     /* one entry per query with >= 1 hit                   has_overlaps / first   */
#define IV_MODE_CONTAIN 2 /* pairs with subject.start <= q.start && q.end <= subject.end             */
#define IV_MODE_LAST 3    /* like ANY, the subject is the hit that ENDS LAST (ties: smallest start, row)  last_overlap_both */

/* nearest direction per query segment (pyranges/methods/nearest.py:85-90) */
#define IV_NEAREST_BOTH 0
#define IV_NEAREST_PREVIOUS 1
#define IV_NEAREST_NEXT 2

/* index flags */
#define IV_INDEX_ENDS_PERM 1 /* also keep the row permutation of the end-sorted order (nearest) */
#define IV_INDEX_GRAPH 2     /* opt in: the second build with identical arguments (pointers, sizes, stream) captures its
                                launch sequence into a CUDA graph, later ones replay it (kernel arguments are by value,
                                data is read at replay time: contents may change, addresses may not).  At most 8
                                sequences are kept per process; never used on the legacy default stream; the environment
                                variable IV_NO_GRAPH=1 disables it globally.  Without the flag every build is plain
                                launches. */

/* Row groups of a set of intervals.  Group g owns rows [row_off[g], row_off[g+1]).
 * lo[g] <= every start, hi[g] >= every end of the group.  base[] is the flattened-coordinate origin:
 * base[0] = 0, base[g+1] >= base[g] + (hi[g]-lo[g]) + gap, base[n_groups] = total_span.
 * sum_len >= sum over all rows of (end - start): sizes the stabbing lists of iv_index_build;
 * max_len >= the largest (end - start), <= 0 = unknown (both ignored by the unary operations). */
typedef struct iv_groups {
    int32_t n_groups;
    int32_t _pad;
    const int64_t* row_off; /* [n_groups+1] */
    const int64_t* lo;      /* [n_groups]   */
    const int64_t* hi;      /* [n_groups]   */
    const int64_t* base;    /* [n_groups+1] */
    int64_t total_span;     /* host copy of base[n_groups] */
    int64_t sum_len;
    int64_t max_len;
} iv_groups_t;

/* One 32-byte record (= one L2 sector, fetched with a single 256-bit load) per coordinate bin
 * b = flattened coordinate >> shift.  It answers "how many starts / ends lie before x" for any x of the
 * bin without touching another array as long as the bin holds at most four starts and four ends. */
typedef struct iv_binrec {
    uint32_t bs, be;  /* #starts < b<<shift, #ends < b<<shift                                   */
    uint32_t co;      /* stabbing list of the bin: cross[co, co + nc)                            */
    uint8_t ns, ne, nc; /* starts / ends / stabbing entries of the bin; 255 = "255 or more": the
                           exact count is the difference to the next record                      */
    uint8_t flags;    /* IV_BIN_INLINE: so[] / eo[] are valid (shift <= 15: offsets are 15-bit); IV_BIN_FAST */
    uint16_t so[4];   /* offsets (x - (b<<shift)) of the bin's first four starts, 0x7FFF = none   */
    uint16_t eo[4];   /* ... and first four ends                                                  */
} iv_binrec_t;
#define IV_BIN_INLINE 1
#define IV_BIN_FAST 2 /* inline and ns <= 4 and ne <= 4: both ranks come out of the record alone */

/* Stabbing-list entry: a subject that starts before the bin and ends after the bin's first coordinate. */
typedef struct iv_cross {
    uint32_t row;  /* subject row */
    uint32_t eoff; /* min(flattened end - (b<<shift), 2^32-1) */
} iv_cross_t;

/* Subject index descriptor, filled by iv_index_build; all arrays live in the caller's workspace. */
typedef struct iv_index {
    int64_t n;             /* subjects */
    int64_t total_span;
    int32_t shift;         /* bin = flattened coordinate >> shift (shift <= 31) */
    int32_t flags;
    int64_t n_bins;        /* bins[] holds n_bins + 1 records */
    int64_t cross_cap;     /* capacity of cross[] */
    const uint64_t* key_s; /* [n] flattened starts, ascending */
    const uint32_t* row_s; /* [n] subject row of each entry of key_s */
    const uint64_t* end_s; /* [n] flattened end of the same entries (start order) */
    const uint64_t* key_e; /* [n] flattened ends, ascending */
    const uint32_t* row_e; /* [n] subject row of key_e entries (IV_INDEX_ENDS_PERM) or NULL */
    const uint32_t* soff;  /* [n] key_s - (bin << shift)                      (start order) */
    const uint32_t* seoff; /* [n] min(end_s - (bin of key_s << shift), 2^32-1) (start order) */
    const uint32_t* eoff;  /* [n] key_e - (bin << shift)                      (end order)   */
    const iv_binrec_t* bins;
    const iv_cross_t* cross;
    const uint32_t* cross_p; /* start-order position of every cross[] entry */
    iv_groups_t groups;
} iv_index_t;

/* Query rows: key segments in the same concatenated layout. */
typedef struct iv_queries {
    int64_t n;
    const int64_t* start;
    const int64_t* end;
    const int64_t* label;     /* NULL -> row number */
    int32_t n_seg;
    int32_t _pad;
    const int64_t* seg_off;   /* [n_seg+1] */
    const int32_t* seg_group; /* [n_seg] partner group of the subject index, -1 = none */
    const int32_t* seg_mode;  /* [n_seg] IV_NEAREST_* (iv_nearest only; NULL -> BOTH) */
    int64_t slack;            /* join(slack=k): start = max(start-k, 0), end += k (multithreaded.py:420-423) */
} iv_queries_t;

int iv_abi_version(void);
const char* iv_last_error_string(void);
int iv_set_device(int device);
/* number of CUDA kernels this library has launched in this process (bench.py: gpu_launches) */
uint64_t iv_launch_count(void);

/* ---- building blocks (exported for tests and the bench) -------------------------------------- */

size_t iv_sort_workspace_bytes(int64_t n, int32_t value_bytes);
/* LSD radix sort of 64-bit keys on bits [begin_bit, end_bit), optional 4/8-byte payload, stable.
 * Sorted data ends up in keys / values. */
int iv_radix_sort_u64(uint64_t* keys, void* values, int32_t value_bytes, int64_t n, int32_t begin_bit,
                      int32_t end_bit, void* ws, size_t ws_bytes, void* stream);

/* out_lo[k] = min start, out_hi[k] = max end, out_maxlen[k] = max(end-start), out_sumlen[k] = sum(end-start),
 * out_minlen[k] = min(end-start) of segment k (INT64_MAX / INT64_MIN / 0 / 0 / INT64_MAX for an empty segment).
 * out_minlen may be NULL. */
int iv_segment_bounds(const int64_t* starts, const int64_t* ends, const int64_t* seg_off, int32_t n_seg,
                      int64_t* out_lo, int64_t* out_hi, int64_t* out_maxlen, int64_t* out_sumlen, int64_t* out_minlen,
                      void* stream);

/* ---- binary operations ------------------------------------------------------------------------ */

size_t iv_index_workspace_bytes(int64_t n, int64_t total_span, int64_t sum_len, int32_t flags);
int iv_index_build(const int64_t* starts, const int64_t* ends, int64_t n, const iv_groups_t* groups,
                   int32_t flags, void* ws, size_t ws_bytes, iv_index_t* out, void* stream);

size_t iv_overlap_workspace_bytes(int64_t n_queries);
/* out_count[i] = number of entries query i emits under `mode`.  With a workspace (ws != NULL, 256-byte aligned,
 * iv_overlap_workspace_bytes(n) bytes) the call also leaves what iv_emit_pairs works from -- per 1024-row tile the
 * output offset, per 64-row granule the list of its queries WITH hits (row, count, flattened interval) -- and the
 * grand total in *out_total (device); out_count may then be NULL (join does not need the per-row counts).
 * Without a workspace only out_count is written. */
int iv_count_overlaps(const iv_index_t* index, const iv_queries_t* q, int32_t mode, int64_t* out_count,
                      void* ws, size_t ws_bytes, int64_t* out_total, void* stream);
/* Writes the pairs of every query whose entries all lie below out_capacity (all of them when total <= out_capacity)
 * from the hit lists iv_count_overlaps left in ws (`counts` is not read
 * and may be NULL; the parameter is kept for callers of ABI v3).  out_q gets the query label, out_s the subject row (or
 * s_label[row]).  Either may be NULL.  out_capacity = entries the output buffers hold: the exact total read back
 * from iv_count_overlaps, or a guess when the emit is launched speculatively before the total is known (entries
 * beyond the capacity are dropped, never written; the caller compares the total and repeats the call if needed).
 * In mode ANY out_s receives the FIRST hit: smallest start, then largest end, then smallest row (the head of the
 * nested-containment-list traversal); in mode LAST the last one: largest start, then smallest end, then largest
 * row. */
int iv_emit_pairs(const iv_index_t* index, const iv_queries_t* q, int32_t mode, const int64_t* counts,
                  const void* ws, int64_t* out_q, int64_t* out_s, const int64_t* s_label, int64_t out_capacity,
                  void* stream);
/* out_seg_pairs[k] = number of entries the query rows of segment k emit, read off the workspace and the total that
 * iv_count_overlaps (same index, queries, mode) left: the per-key result counts of pyrange_apply
 * (pyranges/multithreaded.py:77-103) are known BEFORE the pairs exist, so a multi-GPU caller can exchange them
 * (global offsets of the concatenated result) while iv_emit_pairs runs. */
int iv_segment_pair_counts(const iv_index_t* index, const iv_queries_t* q, const void* ws, const int64_t* total,
                           int64_t* out_seg_pairs, void* stream);
/* out_bp[i] = sum over hits of min(e,E)-max(s,S); out_fraction[i] = out_bp / (end-start), NaN -> 0 */
int iv_coverage_bp(const iv_index_t* index, const iv_queries_t* q, int64_t* out_bp, double* out_fraction,
                   void* stream);
/* subtract: NCLS.set_difference_helper (pyranges/methods/subtraction.py:63-65).  The index must hold DISJOINT
 * subjects per group (the caller merges `other` first, pyranges/pyranges_main.py:4868).  Phase 1 leaves the
 * exclusive piece offsets in out_offsets[nq] and the number of pieces in *out_total (device); phase 2 writes one
 * (query label, start, end) per surviving piece, query order.  A query without overlap survives as one piece, a
 * completely covered query yields none. */
size_t iv_subtract_workspace_bytes(int64_t n_queries);
int iv_subtract_count(const iv_index_t* index, const iv_queries_t* q, int64_t* out_offsets, void* ws, size_t ws_bytes,
                      int64_t* out_total, void* stream);
int iv_subtract_emit(const iv_index_t* index, const iv_queries_t* q, const int64_t* offsets, int64_t* out_q,
                     int64_t* out_start, int64_t* out_end, void* stream);
/* out_row[i] = subject row of the nearest interval or -1, out_dist[i] = 0 (overlap), gap+1, or -1 */
int iv_nearest(const iv_index_t* index, const iv_queries_t* q, int32_t overlap, int64_t* out_row,
               int64_t* out_dist, void* stream);

/* k-nearest candidates (pyranges/methods/k_nearest.py:6-48): for every query the non-overlapping subjects before it
 * (End <= query Start), nearest first, then those behind it (Start >= query End), nearest first; the direction(s) follow
 * q->seg_mode like iv_nearest.  k[i] = how many to keep for query row i, per direction; ties: */
#define IV_TIES_NONE 0      /* the k nearest subjects                                                       */
#define IV_TIES_FIRST 1     /* one subject per distinct distance (the first met), k distinct distances      */
#define IV_TIES_LAST 2      /* ... the last met                                                             */
#define IV_TIES_DIFFERENT 3 /* every subject of the k nearest distinct distances                            */
/* Subjects at equal distance are met in descending row order before the query, ascending row order behind it.  Phase 1
 * leaves the exclusive entry offsets in out_offsets[nq] and the number of entries in *out_total (device); phase 2
 * writes (query label, subject row, distance = gap + 1 >= 1) per entry, query order.  The index needs
 * IV_INDEX_ENDS_PERM.  Overlapping subjects are not reported (the caller joins them in, pyranges_main.py:2737-2746). */
size_t iv_k_nearest_workspace_bytes(int64_t n_queries);
int iv_k_nearest_count(const iv_index_t* index, const iv_queries_t* q, const int64_t* k, int32_t ties,
                       int64_t* out_offsets, void* ws, size_t ws_bytes, int64_t* out_total, void* stream);
int iv_k_nearest_emit(const iv_index_t* index, const iv_queries_t* q, const int64_t* k, int32_t ties,
                      const int64_t* offsets, int64_t* out_q, int64_t* out_s, int64_t* out_dist, void* stream);

/* ---- list index + fused join: the join / count_overlaps / overlap fast path ------------------------------------ */
/* Every coordinate bin (bin = flattened coordinate >> shift, shift <= 14) owns its TOUCH LIST: the subjects that
 * intersect the bin, i.e. those that start before it and are still open at its origin ("crossing") followed by those
 * that start inside it.  A query whose start falls into bin b overlaps exactly
 *      the entries of list b with  (crossing or S' < e')  and  E' > s'
 *    + for every later bin the query reaches: its NON-crossing entries with  S' < e',
 * every subject once.  No sort is needed to build the lists (difference array -> offsets -> fill -> per-bin ordering),
 * and one 32-byte record (= one L2 sector, one 256-bit load) holds a bin's list offset, length and first three entries:
 * most queries are answered -- counted AND emitted -- from that single gather.
 *
 * Entry test word (se): with sp = 0 for a crossing subject, else S' - origin + 1 (< 2^14), and eo = min(E' - origin,
 * 0x7fff):      se = 0x80008000 | (0x4000 - sp) | eo << 16.
 * Query word:   qw = (0x4000 - (e' - origin)) | (s' - origin + 1) << 16   for a query that ends within bin width + reach
 * of the origin, else (0x4000 - bin width) | ...: only the subjects of this bin, the later bins are walked.
 * Then   hit  <=>  ((se - qw) & 0x80008000) == 0x80008000:   the low half keeps its guard bit iff
 * sp <= e' - origin (the subject starts before the query ends), the high half iff eo >= s' - origin + 1 (it ends after
 * the query starts); no borrow crosses the halves.
 * Zero-length subjects (start == end) are handled exactly (hit iff s < start < e) as long as the group window is one
 * coordinate wider than its rows on both sides (lo <= min start - 1, hi >= max end + 1). */
#define IV_L_MAX_SHIFT 14
#define IV_L_GUARD 0x80008000u
#define IV_L_SMAX 0x4000u /* low half = IV_L_SMAX - sp */
#define IV_L_EMAX 0x7fffu /* saturated end offset */
typedef struct iv_lentry {
    uint32_t row; /* subject row */
    uint32_t se;  /* test word */
} iv_lentry_t;
typedef struct iv_lrec {
    uint32_t co;     /* the bin's list is cand_se / cand_row [co, co + n) */
    uint32_t n8;     /* min(n, 255) in the low byte; n = co of the next record - co when it says 255 */
    iv_lentry_t e[3]; /* the first three entries of the list (crossing entries by row, then starts by (start, row));
                         unused ones carry se = IV_L_GUARD (end offset 0: never a hit) */
} iv_lrec_t;
typedef struct iv_lists {
    int64_t n;          /* subjects */
    int64_t total_span;
    int32_t shift;
    int32_t reach;      /* queries up to this long are answered from the list of their first bin alone (see below) */
    int64_t n_bins;     /* rec[] holds n_bins + 2 records (the last two are sentinels) */
    int64_t cand_cap;   /* capacity of cand_se[] / cand_row[] */
    const iv_lrec_t* rec;
    const uint32_t* cand_se;  /* test word of every list entry (all lists back to back, bin order) */
    const uint32_t* cand_row; /* subject row of the same entries */
    iv_groups_t groups;
} iv_lists_t;

/* 0 = these subjects cannot be indexed this way (flattened span beyond 2^41, or so many long subjects that the lists
 * would exceed 2^31 entries): use iv_index_build */
size_t iv_lists_workspace_bytes(int64_t n, int64_t total_span, int64_t sum_len);
/* bin width (log2) iv_lists_build will use for these subjects; < 0 = unsupported.  Callers route queries that are
 * much longer than a bin to the rank index (a query costs one gather per bin it reaches). */
int32_t iv_lists_shift(int64_t n, int64_t total_span, int64_t sum_len);
/* reach (>= 0, a hint: e.g. the longest query that will be asked; clipped to one bin width): subjects that start within
 * `reach` of a bin's origin are ALSO listed in the bin before (sp > bin width), so that a query of at most that length
 * never has to look at a second bin.  Longer queries stay exact: they ignore those entries and walk the later bins. */
int iv_lists_build(const int64_t* starts, const int64_t* ends, int64_t n, const iv_groups_t* groups, int64_t reach,
                   void* ws, size_t ws_bytes, iv_lists_t* out, void* stream);

size_t iv_join_workspace_bytes(int64_t n_queries);
/* All (query label | row, subject row | s_label[row]) pairs in query order, ONE kernel: every 256-row tile (a warp) counts its
 * hits, stages its pairs in shared memory, obtains its output offset by a decoupled look-back over the tiles before
 * it and writes its pairs with coalesced stores.  out_capacity = entries the output buffers hold; pairs beyond it are
 * dropped (never written) and *out_total (device) receives the true total, so the caller can repeat the call with
 * larger buffers.  ws: iv_join_workspace_bytes(q->n) bytes, 256-byte aligned.
 * mode IV_MODE_ANY: one entry per query WITH hits -- out_q only (has_overlaps / overlap(how="first"),
 * pyranges/methods/intersection.py:78), out_s is not written. */
/* mode flag of iv_join_pairs (IV_MODE_ALL | IV_JOIN_HINT_ORDERED): the query rows come in position order (a BAM-sorted
 * read set): rows with many hits then sit next to each other, whole tiles outgrow the staging areas, and the kernel
 * with one larger staging area per warp (k_join) serves them better than the pipelined one (k_join_pipe).  A hint:
 * results do not depend on it. */
#define IV_JOIN_HINT_ORDERED 0x100
int iv_join_pairs(const iv_lists_t* lists, const iv_queries_t* q, int32_t mode, int64_t* out_q, int64_t* out_s,
                  const int64_t* s_label, int64_t out_capacity, void* ws, size_t ws_bytes, int64_t* out_total,
                  void* stream);
/* out_count[i] = number of overlapping subjects (IV_MODE_ALL) or 0/1 (IV_MODE_ANY) of query i */
int iv_join_count(const iv_lists_t* lists, const iv_queries_t* q, int32_t mode, int64_t* out_count, void* stream);
/* out_seg_pairs[k] = number of pairs of query segment k in a pair list written in query order WITHOUT labels
 * (pair_q ascending rows): binary searches of the segment boundaries over the first min(*total, capacity) pairs
 * (total: device). */
int iv_sorted_pair_segments(const int64_t* pair_q, const int64_t* total, int64_t capacity, const int64_t* seg_off,
                            int32_t n_seg, int64_t* out_seg_pairs, void* stream);

/* ---- unary operations ------------------------------------------------------------------------- */
/* groups->base must leave a gap of max(slack,0)+1 between consecutive groups. `len_bits` = number of
 * bits of the largest (end-start). */

size_t iv_cluster_workspace_bytes(int64_t n);
/* out_row[p] = input row at sorted position p (sorted by group, start, end, row);
 * out_id[p] = 1-based cluster id, global over groups in order;  *out_n_clusters (device). */
int iv_cluster(const int64_t* starts, const int64_t* ends, int64_t n, const iv_groups_t* groups, int64_t slack,
               int32_t len_bits, int64_t* out_row, int64_t* out_id, int64_t* out_n_clusters, void* ws,
               size_t ws_bytes, void* stream);
/* phase 1: sort + scans, *out_n_clusters (device).  phase 2: one row per cluster. */
int iv_merge_count(const int64_t* starts, const int64_t* ends, int64_t n, const iv_groups_t* groups,
                   int64_t slack, int32_t len_bits, int64_t* out_n_clusters, void* ws, size_t ws_bytes,
                   void* stream);
int iv_merge_emit(int64_t n, int64_t n_clusters, const iv_groups_t* groups, int32_t len_bits, const void* ws,
                  int64_t* out_start, int64_t* out_end, int64_t* out_count, int32_t* out_group, void* stream);

size_t iv_rle_workspace_bytes(int64_t n, int32_t n_groups);
/* groups->lo must be 0 (coverage is counted from coordinate 0).  out_run_off [n_groups+1] (device):
 * runs of group g are [out_run_off[g], out_run_off[g+1]); out_run_off[n_groups] = total runs. */
int iv_rle_count(const int64_t* starts, const int64_t* ends, int64_t n, const iv_groups_t* groups, void* ws,
                 size_t ws_bytes, int64_t* out_run_off, void* stream);
/* n_runs = out_run_off[n_groups] read back by the caller; run_off = the array iv_rle_count filled. */
int iv_rle_emit(int64_t n, const iv_groups_t* groups, const void* ws, int64_t n_runs, const int64_t* run_off,
                int64_t* out_runs, double* out_values, void* stream);

/* to_rle(value_col=...): the same with float64 weights per row (pyranges/pyranges_main.py:5838-5856).  The level
 * is a float64 sum restarted at every group; a parallel sum can differ from pyrle's sequential one in the last
 * bits (exact for integer-valued weights). */
size_t iv_rle_weighted_workspace_bytes(int64_t n, int32_t n_groups);
int iv_rle_weighted_count(const int64_t* starts, const int64_t* ends, const double* weights, int64_t n,
                          const iv_groups_t* groups, void* ws, size_t ws_bytes, int64_t* out_run_off, void* stream);
int iv_rle_weighted_emit(int64_t n, const iv_groups_t* groups, const void* ws, int64_t n_runs, const int64_t* run_off,
                         int64_t* out_runs, double* out_values, void* stream);

/* ---- split (feeds the multi-sample count matrix, pyranges/multioverlap.py:146) ---- */
/* Every pair of consecutive distinct Start/End points of a group is one feature [point, next point) -- with the gaps
 * between intervals (split(between=True)); the caller drops those with an overlap pass.  groups->lo / hi = min Start /
 * max End of the group.  Phase 1 leaves, in out_group_off [n_groups+1] (device), the features before each group and
 * the total; phase 2 writes them group-major in coordinate order. */
size_t iv_split_workspace_bytes(int64_t n);
int iv_split_count(const int64_t* starts, const int64_t* ends, int64_t n, const iv_groups_t* groups, void* ws,
                   size_t ws_bytes, int64_t* out_group_off, void* stream);
int iv_split_emit(int64_t n, const iv_groups_t* groups, const void* ws, int64_t n_features, int64_t* out_start,
                  int64_t* out_end, int32_t* out_group, void* stream);

/* ---- device-side materialisation of the joined frame --------------------- */
/* For every pair i: qr = pair_q[i] - q_row_offset, sr = pair_s[i] - s_row_offset (rows of the query / subject
 * columns the pair lists refer to; iv_emit_pairs without labels writes global rows, offsets 0).
 *   out_q_start/out_q_end = the query row's Start / End   (clip != 0: clipped to the overlap, i.e. intersect)
 *   out_s_start/out_s_end = the subject row's Start / End (the "_b" columns of join)
 *   out_overlap           = min(End, End_b) - max(Start, Start_b)           (join(report_overlap=True))
 * Any output may be NULL. */
int iv_materialize_pairs(const int64_t* q_start, const int64_t* q_end, const int64_t* s_start,
                         const int64_t* s_end, const int64_t* pair_q, const int64_t* pair_s, int64_t n_pairs,
                         int64_t q_row_offset, int64_t s_row_offset, int32_t clip, int64_t* out_q_start,
                         int64_t* out_q_end, int64_t* out_s_start, int64_t* out_s_end, int64_t* out_overlap,
                         void* stream);
/* out[i] = column[rows[i]] for elements of 1, 2, 4 or 8 bytes; rows[i] < 0 (the filler row of outer joins) gives
 * *fill (host pointer to one element, NULL = zero bytes). */
int iv_gather(const void* column, int32_t elem_bytes, const int64_t* rows, int64_t n, void* out, const void* fill,
              void* stream);

/* ---- ingest: BED text -> device columns (pyranges/readers.py:63-153) ----------------------------------------------
 * The file's bytes live on the device (buf, nbytes; readable up to the next multiple of 16).
 *   iv_text_count_lines: *out_newlines (device) = number of '\n'; per-tile offsets stay in ws for iv_text_line_starts.
 *     n_lines = newlines + 1 if the text does not end in '\n' (the caller looks at the last byte).
 *   iv_text_line_starts: out_line_start[n_lines + 1]; line i = bytes [line_start[i], line_start[i + 1] - 1).
 *   iv_bed_parse: rows = lines [first_line, first_line + n_rows) (first_line = 1 skips a header).  Start / End / Score
 *     (fields 1, 2, 4) as int64, FNV-1a hash of field 0, field 5 as one byte (0: absent), number of fields, and the
 *     spans of the first 12 fields as [12][n_rows] arrays (a trailing '\r' is not part of the last field).
 *     out_bad[3] (device): rows whose Start / End are not integers, Score fields that are not, Strand fields longer than
 *     one character -- the caller decides what to do with such a file.
 *   iv_text_gather: the bytes of n spans packed back to back + their offsets (call with out_bytes = NULL first: the
 *     total lands in out_off[n]). */
size_t iv_text_workspace_bytes(int64_t nbytes);
int iv_text_count_lines(const uint8_t* buf, int64_t nbytes, void* ws, size_t ws_bytes, int64_t* out_newlines, void* stream);
int iv_text_line_starts(const uint8_t* buf, int64_t nbytes, const void* ws, int64_t n_lines, int64_t* out_line_start,
                        void* stream);
int iv_bed_parse(const uint8_t* buf, const int64_t* line_start, int64_t first_line, int64_t n_rows, int64_t* out_start,
                 int64_t* out_end, int64_t* out_score, uint64_t* out_chrom_hash, uint8_t* out_strand,
                 uint8_t* out_n_fields, int64_t* out_field_off, int32_t* out_field_len, int32_t* out_bad, void* stream);
int iv_text_gather(const uint8_t* buf, const int64_t* off, const int32_t* len, int64_t n, int64_t* out_off,
                   uint8_t* out_bytes, void* ws, size_t ws_bytes, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* PYRANGES_B200_H */
