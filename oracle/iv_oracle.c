/*
 * iv_oracle.c -- CPU ORACLE for the pyranges interval hot path.
 *
 * THIS IS TEST INFRASTRUCTURE, NOT PRODUCT CODE.  Only tests/, bench.py's
 * cpu_baseline / --impl reference legs and __graft_entry__.smoke() may load it.
 * The product path (pyranges_b200/) never imports, links or calls anything here.
 *
 * What it restates.  pyranges v0 (/root/reference, pure Python) delegates all
 * interval arithmetic of the hot path to three un-vendored PyPI dependencies
 * that are absent from /root/reference and from this machine:
 *     ncls >= 0.0.63, sorted_nearest >= 0.0.33 (pyproject.toml:24), pyrle (methods/to_rle.py:6-9).
 * Their *published* algorithms are restated here in plain C, anchored on the
 * reference's own call sites (cited per function) and pinned by replaying the
 * reference's doctest / unit-test known answers through the unmodified
 * reference Python layer (oracle/make_golden.py, tests/test_oracle_known_answers.py).
 *
 *   ncls            nested containment list (Alekseyenko & Lee, Bioinformatics 2007):
 *                   sort by (start asc, end desc), peel contained intervals into
 *                   sublists so every list is monotone in start AND end; query =
 *                   binary search for first end > q_start, walk while start < q_end,
 *                   recurse into sublists.           call sites: methods/join.py:12-23,
 *                   methods/intersection.py:19-28,71-78, methods/coverage.py:20-26,58-64
 *   sorted_nearest  linear scans over sorted arrays. call sites: methods/cluster.py:13,
 *                   methods/merge.py:14, methods/nearest.py:59,69,113
 *   pyrle           event sort + cumulative sum.     call site: methods/to_rle.py:18
 *
 * Interval convention: half-open [start, end), int64, overlap iff S < e && s < E
 * (doctests pyranges_main.py:3563-3572, :2355-2362).
 */
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

typedef int64_t i64;

/* ------------------------------------------------------------------ */
/* growable int64 buffer (mirrors ncls' growing output arrays)          */
/* ------------------------------------------------------------------ */
typedef struct { i64 *p; i64 n, cap; } vec64;
static void vpush(vec64 *v, i64 x) {
    if (v->n == v->cap) {
        v->cap = v->cap ? v->cap * 2 : 1024;
        v->p = (i64 *)realloc(v->p, (size_t)v->cap * sizeof(i64));
    }
    v->p[v->n++] = x;
}

void ivo_free(void *p) { free(p); }

/* ------------------------------------------------------------------ */
/* NCLS restatement                                                      */
/* ------------------------------------------------------------------ */
typedef struct {
    i64 n;
    i64 n_top;          /* length of the top-level list (positions [0, n_top)) */
    i64 *start, *end, *id;
    i64 *sub_off;       /* first position of this interval's sublist, -1 if none */
    i64 *sub_len;
} ncls_t;

typedef struct { i64 s, e, id, pos; } item_t;

static int cmp_item(const void *a, const void *b) {
    const item_t *x = (const item_t *)a, *y = (const item_t *)b;
    if (x->s != y->s) return x->s < y->s ? -1 : 1;
    if (x->e != y->e) return x->e > y->e ? -1 : 1;      /* end DESC: container first */
    if (x->pos != y->pos) return x->pos < y->pos ? -1 : 1; /* deterministic (upstream: qsort, unstable) */
    return 0;
}

/* NCLS(starts, ends, ids)   methods/join.py:12 */
ncls_t *ivo_ncls_build(const i64 *starts, const i64 *ends, const i64 *ids, i64 n) {
    ncls_t *t = (ncls_t *)calloc(1, sizeof(ncls_t));
    t->n = n;
    if (n == 0) return t;
    item_t *it = (item_t *)malloc((size_t)n * sizeof(item_t));
    for (i64 i = 0; i < n; i++) { it[i].s = starts[i]; it[i].e = ends[i]; it[i].id = ids[i]; it[i].pos = i; }
    qsort(it, (size_t)n, sizeof(item_t), cmp_item);

    /* parent[i] = innermost earlier interval that contains i, or -1 */
    i64 *parent = (i64 *)malloc((size_t)n * sizeof(i64));
    i64 *stack = (i64 *)malloc((size_t)n * sizeof(i64));
    i64 sp = 0;
    for (i64 i = 0; i < n; i++) {
        while (sp > 0 && it[stack[sp - 1]].e < it[i].e) sp--;   /* top does not contain i */
        parent[i] = sp > 0 ? stack[sp - 1] : -1;
        stack[sp++] = i;
    }
    /* counting sort by parent (+1), stable => every list stays in start order */
    i64 *cnt = (i64 *)calloc((size_t)n + 2, sizeof(i64));
    for (i64 i = 0; i < n; i++) cnt[parent[i] + 1 + 1]++;
    for (i64 k = 1; k <= n + 1; k++) cnt[k] += cnt[k - 1];      /* cnt[p+1] = offset of list p */
    i64 *off = (i64 *)malloc(((size_t)n + 1) * sizeof(i64));
    memcpy(off, cnt, ((size_t)n + 1) * sizeof(i64));
    i64 *newpos = (i64 *)malloc((size_t)n * sizeof(i64));
    for (i64 i = 0; i < n; i++) newpos[i] = cnt[parent[i] + 1]++;

    t->start = (i64 *)malloc((size_t)n * sizeof(i64));
    t->end = (i64 *)malloc((size_t)n * sizeof(i64));
    t->id = (i64 *)malloc((size_t)n * sizeof(i64));
    t->sub_off = (i64 *)malloc((size_t)n * sizeof(i64));
    t->sub_len = (i64 *)malloc((size_t)n * sizeof(i64));
    t->n_top = (n >= 1) ? off[1] - off[0] : 0;
    for (i64 i = 0; i < n; i++) {
        i64 p = newpos[i];
        t->start[p] = it[i].s; t->end[p] = it[i].e; t->id[p] = it[i].id;
        i64 len = (i + 2 <= n) ? off[i + 2] - off[i + 1] : n - off[i + 1];
        t->sub_len[p] = len;
        t->sub_off[p] = len > 0 ? off[i + 1] : -1;
    }
    free(it); free(parent); free(stack); free(cnt); free(off); free(newpos);
    return t;
}

void ivo_ncls_free(ncls_t *t) {
    if (!t) return;
    free(t->start); free(t->end); free(t->id); free(t->sub_off); free(t->sub_len); free(t);
}

/* first position in list [a, b) whose end > s (ends are non-decreasing inside a list) */
static i64 list_first(const ncls_t *t, i64 a, i64 b, i64 s) {
    while (a < b) { i64 m = a + ((b - a) >> 1); if (t->end[m] > s) b = m; else a = m + 1; }
    return a;
}

typedef void (*hit_fn)(void *ctx, i64 pos);

/* pre-order walk of all hits = the "nested-list traversal order" */
static void ncls_walk(const ncls_t *t, i64 a, i64 b, i64 s, i64 e, hit_fn f, void *ctx) {
    for (i64 i = list_first(t, a, b, s); i < b && t->start[i] < e; i++) {
        f(ctx, i);
        if (t->sub_off[i] >= 0) ncls_walk(t, t->sub_off[i], t->sub_off[i] + t->sub_len[i], s, e, f, ctx);
    }
}

typedef struct { const ncls_t *t; vec64 *q, *s; i64 qid, qs, qe; i64 cnt; i64 first, last; i64 cov; } qctx;

static void hit_all(void *c, i64 pos) { qctx *x = (qctx *)c; vpush(x->q, x->qid); vpush(x->s, x->t->id[pos]); }
static void hit_contain(void *c, i64 pos) {
    qctx *x = (qctx *)c;
    if (x->t->start[pos] <= x->qs && x->qe <= x->t->end[pos]) { vpush(x->q, x->qid); vpush(x->s, x->t->id[pos]); }
}
static void hit_count(void *c, i64 pos) { qctx *x = (qctx *)c; (void)pos; x->cnt++; }
/* first: the first hit of the traversal.  last: the hit that ends last (the first such hit met) -- ncls keeps the hit with
 * the largest end; pinned by the reference's tests/unit/k_nearest.py:163-211 (last overlap of [15,20) among [11,16) and
 * [11,20) is [11,20)); "the last hit of the traversal" would give [11,16) there. */
static void hit_firstlast(void *c, i64 pos) {
    qctx *x = (qctx *)c;
    if (x->cnt++ == 0) { x->first = pos; x->last = pos; }
    else if (x->t->end[pos] > x->t->end[x->last]) x->last = pos;
}
static void hit_cov(void *c, i64 pos) {
    qctx *x = (qctx *)c;
    i64 lo = x->t->start[pos] > x->qs ? x->t->start[pos] : x->qs;
    i64 hi = x->t->end[pos] < x->qe ? x->t->end[pos] : x->qe;
    x->cov += hi - lo;
}

/* NCLS.all_overlaps_both  methods/join.py:15,23  coverage.py:26  intersection.py:22
 * mode 0: all overlaps, mode 1: all_containments_both (join.py:17, intersection.py:24,76) */
i64 ivo_ncls_all_overlaps_both(const ncls_t *t, const i64 *qs, const i64 *qe, const i64 *qid, i64 nq,
                               int mode, i64 **out_q, i64 **out_s) {
    vec64 vq = {0, 0, 0}, vs = {0, 0, 0};
    qctx c; memset(&c, 0, sizeof c); c.t = t; c.q = &vq; c.s = &vs;
    if (t->n > 0)
        for (i64 i = 0; i < nq; i++) {
            c.qid = qid[i]; c.qs = qs[i]; c.qe = qe[i];
            ncls_walk(t, 0, t->n_top, qs[i], qe[i], mode ? hit_contain : hit_all, &c);
        }
    *out_q = vq.p; *out_s = vs.p;
    return vq.n;
}

/* NCLS.first_overlap_both / last_overlap_both  methods/join.py:19,21 ; which: 0 first, 1 last */
i64 ivo_ncls_first_overlap_both(const ncls_t *t, const i64 *qs, const i64 *qe, const i64 *qid, i64 nq,
                                int which, i64 **out_q, i64 **out_s) {
    vec64 vq = {0, 0, 0}, vs = {0, 0, 0};
    qctx c; memset(&c, 0, sizeof c); c.t = t;
    if (t->n > 0)
        for (i64 i = 0; i < nq; i++) {
            c.cnt = 0;
            ncls_walk(t, 0, t->n_top, qs[i], qe[i], hit_firstlast, &c);
            if (c.cnt) { vpush(&vq, qid[i]); vpush(&vs, t->id[which ? c.last : c.first]); }
        }
    *out_q = vq.p; *out_s = vs.p;
    return vq.n;
}

/* per-query hit count (has_overlaps intersection.py:78 = count > 0; derived count of SURVEY App. A) */
void ivo_ncls_count(const ncls_t *t, const i64 *qs, const i64 *qe, i64 nq, i64 *out_count) {
    qctx c; memset(&c, 0, sizeof c); c.t = t;
    for (i64 i = 0; i < nq; i++) {
        c.cnt = 0;
        if (t->n > 0) ncls_walk(t, 0, t->n_top, qs[i], qe[i], hit_count, &c);
        out_count[i] = c.cnt;
    }
}

/* NCLS.coverage  methods/coverage.py:64 : sum over hits of (min(e,E) - max(s,S)) */
void ivo_ncls_coverage(const ncls_t *t, const i64 *qs, const i64 *qe, i64 nq, i64 *out_bp) {
    qctx c; memset(&c, 0, sizeof c); c.t = t;
    for (i64 i = 0; i < nq; i++) {
        c.cov = 0; c.qs = qs[i]; c.qe = qe[i];
        if (t->n > 0) ncls_walk(t, 0, t->n_top, qs[i], qe[i], hit_cov, &c);
        out_bp[i] = c.cov;
    }
}

/* NCLS.set_difference_helper(starts, ends, indexes, num)  methods/subtraction.py:63-65
 * The subjects are the MERGED other intervals (pyranges_main.py:4868: disjoint).  For every query with at
 * least one hit: the parts of [s, e) not covered by its hits, in coordinate order, as (id, start, end);
 * a query that is covered completely yields one (id, -1, -1) marker (dropped by the caller,
 * subtraction.py:69-73).  Queries without hit are not reported (they survive unchanged, :67). */
static int cmp_i64_pair(const void *a, const void *b) {
    const i64 *x = (const i64 *)a, *y = (const i64 *)b;
    return x[0] < y[0] ? -1 : (x[0] > y[0] ? 1 : 0);
}
typedef struct { const ncls_t *t; vec64 *h; } dctx;
static void hit_collect(void *c, i64 pos) { dctx *x = (dctx *)c; vpush(x->h, x->t->start[pos]); vpush(x->h, x->t->end[pos]); }
i64 ivo_ncls_set_difference(const ncls_t *t, const i64 *qs, const i64 *qe, const i64 *qid, i64 nq,
                            i64 **out_id, i64 **out_s, i64 **out_e) {
    vec64 vi = {0, 0, 0}, vs = {0, 0, 0}, ve = {0, 0, 0}, h = {0, 0, 0};
    dctx c; c.t = t; c.h = &h;
    for (i64 i = 0; i < nq && t->n > 0; i++) {
        h.n = 0;
        ncls_walk(t, 0, t->n_top, qs[i], qe[i], hit_collect, &c);
        if (h.n == 0) continue;
        qsort(h.p, (size_t)(h.n / 2), 2 * sizeof(i64), cmp_i64_pair);
        i64 cur = qs[i], pieces = 0;
        for (i64 k = 0; k < h.n; k += 2) {
            if (h.p[k] > cur) { vpush(&vi, qid[i]); vpush(&vs, cur); vpush(&ve, h.p[k]); pieces++; }
            if (h.p[k + 1] > cur) cur = h.p[k + 1];
        }
        if (cur < qe[i]) { vpush(&vi, qid[i]); vpush(&vs, cur); vpush(&ve, qe[i]); pieces++; }
        if (!pieces) { vpush(&vi, qid[i]); vpush(&vs, -1); vpush(&ve, -1); }
    }
    free(h.p);
    *out_id = vi.p; *out_s = vs.p; *out_e = ve.p;
    return vi.n;
}

/* ------------------------------------------------------------------ */
/* sorted_nearest restatement                                            */
/* ------------------------------------------------------------------ */

/* annotate_clusters(starts, ends, slack)  methods/cluster.py:13
 * input sorted by start; row joins the running cluster iff start - slack <= max_end */
void ivo_annotate_clusters(const i64 *S, const i64 *E, i64 n, i64 slack, i64 *out_id) {
    if (n == 0) return;
    i64 id = 1, m = E[0];
    out_id[0] = 1;
    for (i64 i = 1; i < n; i++) {
        if (S[i] - slack <= m) { if (E[i] > m) m = E[i]; }
        else { id++; m = E[i]; }
        out_id[i] = id;
    }
}

/* find_clusters(starts, ends, slack)  methods/merge.py:14 -> (cluster_start, cluster_max_end, n_members) */
i64 ivo_find_clusters(const i64 *S, const i64 *E, i64 n, i64 slack, i64 *out_s, i64 *out_e, i64 *out_n) {
    if (n == 0) return 0;
    i64 k = 0, m = E[0], cs = S[0], cnt = 1;
    for (i64 i = 1; i < n; i++) {
        if (S[i] - slack <= m) { if (E[i] > m) m = E[i]; cnt++; }
        else { out_s[k] = cs; out_e[k] = m; out_n[k] = cnt; k++; cs = S[i]; m = E[i]; cnt = 1; }
    }
    out_s[k] = cs; out_e[k] = m; out_n[k] = cnt; k++;
    return k;
}

/* cluster_by(starts, ends, by_ids, slack)  methods/cluster.py:48 (sorted_nearest, un-vendored): rows sorted by
 * (by_id, Start); the annotate_clusters scan, except that a change of by_id always opens a new cluster. */
void ivo_cluster_by(const i64 *S, const i64 *E, const i64 *B, i64 n, i64 slack, i64 *out_id) {
    if (n == 0) return;
    i64 id = 1, m = E[0];
    out_id[0] = 1;
    for (i64 i = 1; i < n; i++) {
        if (B[i] == B[i - 1] && S[i] - slack <= m) { if (E[i] > m) m = E[i]; }
        else { id++; m = E[i]; }
        out_id[i] = id;
    }
}

/* merge_by(starts, ends, by_ids, slack)  methods/merge.py:65 -> per cluster (by_id, start, max end, n_members) */
i64 ivo_merge_by(const i64 *S, const i64 *E, const i64 *B, i64 n, i64 slack, i64 *out_b, i64 *out_s, i64 *out_e,
                 i64 *out_n) {
    if (n == 0) return 0;
    i64 k = 0, m = E[0], cs = S[0], cnt = 1, cb = B[0];
    for (i64 i = 1; i < n; i++) {
        if (B[i] == cb && S[i] - slack <= m) { if (E[i] > m) m = E[i]; cnt++; }
        else { out_b[k] = cb; out_s[k] = cs; out_e[k] = m; out_n[k] = cnt; k++; cb = B[i]; cs = S[i]; m = E[i]; cnt = 1; }
    }
    out_b[k] = cb; out_s[k] = cs; out_e[k] = m; out_n[k] = cnt; k++;
    return k;
}

/* nearest_previous_nonoverlapping(l_starts, r_ends - 1, r_idx)  methods/nearest.py:69
 * Ls sorted ascending, Re1 sorted ascending (= End - 1), Ridx aligned with Re1.
 * per left: last j with Re1[j] < Ls[i]; ties among equal Re1 -> FIRST of the equal run
 * (convention: earliest row of the partner frame under a stable sort). distance = Ls - Re1. */
void ivo_nearest_previous(const i64 *Ls, i64 nl, const i64 *Re1, const i64 *Ridx, i64 nr,
                          i64 *out_idx, i64 *out_dist) {
    i64 j = 0;   /* number of Re1 < Ls[i] */
    for (i64 i = 0; i < nl; i++) {
        while (j < nr && Re1[j] < Ls[i]) j++;
        if (j == 0) { out_idx[i] = -1; out_dist[i] = -1; continue; }
        i64 k = j - 1;
        while (k > 0 && Re1[k - 1] == Re1[k]) k--;      /* first of equal run */
        out_idx[i] = Ridx[k]; out_dist[i] = Ls[i] - Re1[k];
    }
}

/* nearest_next_nonoverlapping(l_ends - 1, r_starts, r_idx)  methods/nearest.py:59
 * first j with Rs[j] > Le1[i]; distance = Rs - Le1 */
void ivo_nearest_next(const i64 *Le1, i64 nl, const i64 *Rs, const i64 *Ridx, i64 nr,
                      i64 *out_idx, i64 *out_dist) {
    i64 j = 0;
    for (i64 i = 0; i < nl; i++) {
        while (j < nr && Rs[j] <= Le1[i]) j++;
        if (j == nr) { out_idx[i] = -1; out_dist[i] = -1; continue; }
        out_idx[i] = Ridx[j]; out_dist[i] = Rs[j] - Le1[i];
    }
}

/* nearest_nonoverlapping(prev_idx, prev_dist, next_idx, next_dist)  methods/nearest.py:113
 * smaller distance wins; -1 = missing; tie -> previous (convention) */
void ivo_nearest_nonoverlapping(const i64 *pi, const i64 *pd, const i64 *ni, const i64 *nd, i64 n,
                                i64 *out_idx, i64 *out_dist) {
    for (i64 i = 0; i < n; i++) {
        int hp = pi[i] >= 0, hn = ni[i] >= 0;
        if (hp && (!hn || pd[i] <= nd[i])) { out_idx[i] = pi[i]; out_dist[i] = pd[i]; }
        else if (hn) { out_idx[i] = ni[i]; out_dist[i] = nd[i]; }
        else { out_idx[i] = -1; out_dist[i] = -1; }
    }
}

/* ------------------------------------------------------------------ */
/* pyrle.methods.coverage restatement   methods/to_rle.py:18             */
/* ------------------------------------------------------------------ */
typedef struct { i64 pos; double v; i64 ord; } ev_t;
static int cmp_ev(const void *a, const void *b) {
    const ev_t *x = (const ev_t *)a, *y = (const ev_t *)b;
    if (x->pos != y->pos) return x->pos < y->pos ? -1 : 1;
    if (x->ord != y->ord) return x->ord < y->ord ? -1 : 1;   /* stable (mergesort upstream) */
    return 0;
}

/* events (Start,+v),(End,-v); stable sort by position; sum per distinct position; cumsum;
 * runs = diff(positions); leading zero run if first position > 0; zero-length runs dropped;
 * adjacent equal values coalesced; no trailing zero run (doctests pyranges_main.py:5804-5872).
 * out_runs / out_vals must hold 2n+1 entries.  weights may be NULL (unit weights). */
i64 ivo_coverage_rle(const i64 *S, const i64 *E, const double *w, i64 n, i64 *out_runs, double *out_vals) {
    if (n == 0) return 0;
    ev_t *ev = (ev_t *)malloc((size_t)(2 * n) * sizeof(ev_t));
    for (i64 i = 0; i < n; i++) {
        double v = w ? w[i] : 1.0;
        ev[i].pos = S[i]; ev[i].v = v; ev[i].ord = i;
        ev[n + i].pos = E[i]; ev[n + i].v = -v; ev[n + i].ord = n + i;
    }
    qsort(ev, (size_t)(2 * n), sizeof(ev_t), cmp_ev);
    i64 k = 0;
    double level = 0.0;
    i64 prev_pos = 0;          /* coverage starts at coordinate 0 with value 0 */
    double prev_val = 0.0;
    i64 i = 0;
    while (i < 2 * n) {
        i64 p = ev[i].pos;
        double d = 0.0;
        while (i < 2 * n && ev[i].pos == p) { d += ev[i].v; i++; }
        /* segment [prev_pos, p) has value prev_val */
        i64 len = p - prev_pos;
        if (len > 0) {
            if (k > 0 && out_vals[k - 1] == prev_val) out_runs[k - 1] += len;
            else { out_runs[k] = len; out_vals[k] = prev_val; k++; }
        }
        level += d;
        prev_pos = p; prev_val = level;
    }
    free(ev);
    return k;
}

/* ------------------------------------------------------------------ */
/* brute-force O(nq * ns) reference of the reference (tiny cases only)   */
/* ------------------------------------------------------------------ */
void ivo_brute_count(const i64 *S, const i64 *E, i64 ns, const i64 *qs, const i64 *qe, i64 nq, i64 *out) {
    for (i64 i = 0; i < nq; i++) {
        i64 c = 0;
        for (i64 j = 0; j < ns; j++) c += (S[j] < qe[i] && qs[i] < E[j]);
        out[i] = c;
    }
}
