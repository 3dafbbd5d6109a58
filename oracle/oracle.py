"""ctypes front-end of the CPU oracle (oracle/iv_oracle.c).

TEST INFRASTRUCTURE ONLY: importable from tests/, bench.py (cpu_baseline / --impl reference)
and __graft_entry__.smoke().  Nothing under pyranges_b200/ may import this module.

The classes/functions mirror the numpy-in / numpy-out API of the three third-party
libraries the reference calls: ``ncls.NCLS``, ``sorted_nearest.*``
and ``pyrle.methods.coverage`` so that oracle/shims/* can put them under the *unmodified*
reference Python layer (only in the authoring container; /root/reference does not travel).
"""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SRC = os.path.join(_HERE, "iv_oracle.c")
_LIB = os.path.join(_HERE, "libiv_oracle.so")
_lib = None

_P64 = C.POINTER(C.c_int64)
_PD = C.POINTER(C.c_double)


def build(force=False):
    """gcc -O3 the C restatement into oracle/libiv_oracle.so (git-ignored, travels with gpurun)."""
    if force or not os.path.exists(_LIB) or os.path.getmtime(_LIB) < os.path.getmtime(_SRC):
        subprocess.check_call(
            ["gcc", "-O3", "-march=x86-64-v2", "-fPIC", "-shared", "-o", _LIB, _SRC]
        )
    return _LIB


def lib():
    global _lib
    if _lib is None:
        build()
        L = C.CDLL(_LIB)
        L.ivo_ncls_build.restype = C.c_void_p
        L.ivo_ncls_build.argtypes = [_P64, _P64, _P64, C.c_int64]
        L.ivo_ncls_free.argtypes = [C.c_void_p]
        L.ivo_free.argtypes = [C.c_void_p]
        for f in (L.ivo_ncls_all_overlaps_both, L.ivo_ncls_first_overlap_both):
            f.restype = C.c_int64
            f.argtypes = [C.c_void_p, _P64, _P64, _P64, C.c_int64, C.c_int, C.POINTER(_P64), C.POINTER(_P64)]
        L.ivo_ncls_set_difference.restype = C.c_int64
        L.ivo_ncls_set_difference.argtypes = [C.c_void_p, _P64, _P64, _P64, C.c_int64, C.POINTER(_P64), C.POINTER(_P64),
                                              C.POINTER(_P64)]
        L.ivo_ncls_count.argtypes = [C.c_void_p, _P64, _P64, C.c_int64, _P64]
        L.ivo_ncls_coverage.argtypes = [C.c_void_p, _P64, _P64, C.c_int64, _P64]
        L.ivo_annotate_clusters.argtypes = [_P64, _P64, C.c_int64, C.c_int64, _P64]
        L.ivo_find_clusters.restype = C.c_int64
        L.ivo_find_clusters.argtypes = [_P64, _P64, C.c_int64, C.c_int64, _P64, _P64, _P64]
        L.ivo_cluster_by.argtypes = [_P64, _P64, _P64, C.c_int64, C.c_int64, _P64]
        L.ivo_merge_by.restype = C.c_int64
        L.ivo_merge_by.argtypes = [_P64, _P64, _P64, C.c_int64, C.c_int64, _P64, _P64, _P64, _P64]
        L.ivo_nearest_previous.argtypes = [_P64, C.c_int64, _P64, _P64, C.c_int64, _P64, _P64]
        L.ivo_nearest_next.argtypes = [_P64, C.c_int64, _P64, _P64, C.c_int64, _P64, _P64]
        L.ivo_nearest_nonoverlapping.argtypes = [_P64, _P64, _P64, _P64, C.c_int64, _P64, _P64]
        L.ivo_coverage_rle.restype = C.c_int64
        L.ivo_coverage_rle.argtypes = [_P64, _P64, _PD, C.c_int64, _P64, _PD]
        L.ivo_brute_count.argtypes = [_P64, _P64, C.c_int64, _P64, _P64, C.c_int64, _P64]
        _lib = L
    return _lib


def _a64(x):
    return np.ascontiguousarray(np.asarray(x), dtype=np.int64)


def _p(a):
    return a.ctypes.data_as(_P64)


def _take(ptr, n):
    """copy a malloc'ed int64 buffer of length n into numpy and free it"""
    if n == 0 or not ptr:
        out = np.empty(0, dtype=np.int64)
    else:
        out = np.ctypeslib.as_array(ptr, shape=(n,)).copy()
    if ptr:
        lib().ivo_free(C.cast(ptr, C.c_void_p))
    return out


class NCLS:
    """API twin of ``ncls.NCLS`` (reference call sites: methods/join.py:12, intersection.py:19,71,
    coverage.py:20,58)."""

    def __init__(self, starts, ends, ids):
        s, e, i = _a64(starts), _a64(ends), _a64(ids)
        self._n = len(s)
        self._h = lib().ivo_ncls_build(_p(s), _p(e), _p(i), len(s))

    def __del__(self):
        h, self._h = getattr(self, "_h", None), None
        if h:
            lib().ivo_ncls_free(h)

    def _both(self, fn, starts, ends, indexes, mode):
        s, e, i = _a64(starts), _a64(ends), _a64(indexes)
        oq, os_ = _P64(), _P64()
        n = fn(self._h, _p(s), _p(e), _p(i), len(s), mode, C.byref(oq), C.byref(os_))
        return _take(oq, n), _take(os_, n)

    def all_overlaps_both(self, starts, ends, indexes):
        return self._both(lib().ivo_ncls_all_overlaps_both, starts, ends, indexes, 0)

    def all_containments_both(self, starts, ends, indexes):
        return self._both(lib().ivo_ncls_all_overlaps_both, starts, ends, indexes, 1)

    def first_overlap_both(self, starts, ends, indexes):
        return self._both(lib().ivo_ncls_first_overlap_both, starts, ends, indexes, 0)

    def last_overlap_both(self, starts, ends, indexes):
        return self._both(lib().ivo_ncls_first_overlap_both, starts, ends, indexes, 1)

    def all_overlaps_self(self, starts, ends, indexes):
        return self.all_overlaps_both(starts, ends, indexes)[0]

    def count(self, starts, ends):
        s, e = _a64(starts), _a64(ends)
        out = np.zeros(len(s), dtype=np.int64)
        lib().ivo_ncls_count(self._h, _p(s), _p(e), len(s), _p(out))
        return out

    def has_overlaps(self, starts, ends, indexes):
        return _a64(indexes)[self.count(starts, ends) > 0]

    def set_difference_helper(self, starts, ends, indexes, num=None):
        """-> (idx, new_starts, new_ends); see iv_oracle.c (reference call site methods/subtraction.py:63-65)"""
        s, e, i = _a64(starts), _a64(ends), _a64(indexes)
        oi, os_, oe = _P64(), _P64(), _P64()
        n = lib().ivo_ncls_set_difference(self._h, _p(s), _p(e), _p(i), len(s), C.byref(oi), C.byref(os_), C.byref(oe))
        return _take(oi, n), _take(os_, n), _take(oe, n)

    def coverage(self, starts, ends, indexes):
        s, e = _a64(starts), _a64(ends)
        out = np.zeros(len(s), dtype=np.int64)
        lib().ivo_ncls_coverage(self._h, _p(s), _p(e), len(s), _p(out))
        return out


def annotate_clusters(starts, ends, slack):
    s, e = _a64(starts), _a64(ends)
    out = np.empty(len(s), dtype=np.int64)
    lib().ivo_annotate_clusters(_p(s), _p(e), len(s), int(slack), _p(out))
    return out


def find_clusters(starts, ends, slack):
    s, e = _a64(starts), _a64(ends)
    n = len(s)
    os_, oe, on = (np.empty(n, dtype=np.int64) for _ in range(3))
    k = lib().ivo_find_clusters(_p(s), _p(e), n, int(slack), _p(os_), _p(oe), _p(on))
    return os_[:k].copy(), oe[:k].copy(), on[:k].copy()


def split_points(starts, ends):
    """pyranges/methods/split.py:11-20: concat(Start, End).sort_values().drop_duplicates(); consecutive points are the
    features (points, points.shift(-1))[:-1].  -> (feature starts, feature ends)"""
    pts = np.unique(np.concatenate([_a64(starts), _a64(ends)]))
    return pts[:-1].copy(), pts[1:].copy()


def cluster_by(starts, ends, ids, slack):
    s, e, b = _a64(starts), _a64(ends), _a64(ids)
    out = np.empty(len(s), dtype=np.int64)
    lib().ivo_cluster_by(_p(s), _p(e), _p(b), len(s), int(slack), _p(out))
    return out


def merge_by(starts, ends, ids, slack):
    s, e, b = _a64(starts), _a64(ends), _a64(ids)
    n = len(s)
    ob, os_, oe, on = (np.empty(n, dtype=np.int64) for _ in range(4))
    k = lib().ivo_merge_by(_p(s), _p(e), _p(b), n, int(slack), _p(ob), _p(os_), _p(oe), _p(on))
    return ob[:k].copy(), os_[:k].copy(), oe[:k].copy(), on[:k].copy()


def nearest_previous_nonoverlapping(l_starts, r_ends, r_indexes):
    ls, re_, ri = _a64(l_starts), _a64(r_ends), _a64(r_indexes)
    oi, od = np.empty(len(ls), dtype=np.int64), np.empty(len(ls), dtype=np.int64)
    lib().ivo_nearest_previous(_p(ls), len(ls), _p(re_), _p(ri), len(re_), _p(oi), _p(od))
    return oi, od


def nearest_next_nonoverlapping(l_ends, r_starts, r_indexes):
    le, rs, ri = _a64(l_ends), _a64(r_starts), _a64(r_indexes)
    oi, od = np.empty(len(le), dtype=np.int64), np.empty(len(le), dtype=np.int64)
    lib().ivo_nearest_next(_p(le), len(le), _p(rs), _p(ri), len(rs), _p(oi), _p(od))
    return oi, od


def nearest_nonoverlapping(prev_idx, prev_dist, next_idx, next_dist):
    pi, pd_, ni, nd = _a64(prev_idx), _a64(prev_dist), _a64(next_idx), _a64(next_dist)
    oi, od = np.empty(len(pi), dtype=np.int64), np.empty(len(pi), dtype=np.int64)
    lib().ivo_nearest_nonoverlapping(_p(pi), _p(pd_), _p(ni), _p(nd), len(pi), _p(oi), _p(od))
    return oi, od


def coverage_rle(starts, ends, weights=None):
    """-> (runs int64, values float64) of one key; restates pyrle.methods.coverage (to_rle.py:18)."""
    s, e = _a64(starts), _a64(ends)
    n = len(s)
    runs = np.empty(2 * n + 1, dtype=np.int64)
    vals = np.empty(2 * n + 1, dtype=np.float64)
    w = None
    wp = _PD()
    if weights is not None:
        w = np.ascontiguousarray(np.asarray(weights), dtype=np.float64)
        wp = w.ctypes.data_as(_PD)
    k = lib().ivo_coverage_rle(_p(s), _p(e), wp, n, _p(runs), vals.ctypes.data_as(_PD))
    return runs[:k].copy(), vals[:k].copy()


def brute_count(S, E, qs, qe):
    S, E, qs, qe = _a64(S), _a64(E), _a64(qs), _a64(qe)
    out = np.empty(len(qs), dtype=np.int64)
    lib().ivo_brute_count(_p(S), _p(E), len(S), _p(qs), _p(qe), len(qs), _p(out))
    return out


def brute_pairs(S, E, sid, qs, qe, qid):
    """numpy brute force of all_overlaps_both -- tiny cases only; canonical (q, s) sorted."""
    S, E, qs, qe = _a64(S), _a64(E), _a64(qs), _a64(qe)
    m = (S[None, :] < qe[:, None]) & (qs[:, None] < E[None, :])
    qi, sj = np.nonzero(m)
    a, b = _a64(qid)[qi], _a64(sid)[sj]
    o = np.lexsort([b, a])
    return a[o], b[o]


# ---- k-nearest (sorted_nearest 0.0.39: k_nearest_*_nonoverlapping, get_all_ties, get_different_ties) ----------------
# The package is not vendored in the reference and not installable here; these are restatements of its published
# behaviour as the reference uses it (pyranges/methods/k_nearest.py:6-48, pyranges/pyranges_main.py:2726-2800), pinned by
# the reference's doctests (pyranges_main.py:2560-2700) and unit tests (tests/unit/k_nearest.py) through
# oracle/make_golden.py.  What those answers do not pin is said where it applies.  Plain Python loops: small cases only.
def _k_walk(values, positions, k, ties):
    """positions in walking order (nearest first), values = distance key at each; -> kept positions.
    ties None: the k nearest; "first" / "last": one entry (the first / last met) per distinct value, k distinct values;
    "different": every entry of the k nearest distinct values."""
    out = []
    if k <= 0:
        return out
    if ties is None or ties is False:
        return list(positions[:k])
    distinct, last_v, pend = 0, None, None
    for p, v in zip(positions, values):
        if last_v is None or v != last_v:
            if ties == "last" and pend is not None:
                out.append(pend)
            if distinct == k:
                pend = None
                break
            distinct += 1
            last_v = v
            if ties == "first":
                out.append(p)
        if ties == "different":
            out.append(p)
        pend = p
    if ties == "last" and pend is not None:
        out.append(pend)
    return out


def _k_of(k, labels, i):
    """k of the i-th query of the sorted / filtered arrays: the reference passes d1.__k__.values (frame order) next to
    arrays in sorted order (k_nearest.py:23,46); its per-key frames carry a range index, so the row label finds the
    row's own k.  (Positional vs. label lookup is not pinned by the reference's tests -- both agree there; a too large k
    here is harmless because pyranges_main.py:2745-2790 trims with the row's own k again.)"""
    k = np.asarray(k)
    if k.ndim == 0:
        return int(k)
    lab = int(labels[i])
    return int(k[lab]) if 0 <= lab < len(k) else int(k[min(i, len(k) - 1)])


def k_nearest_previous_nonoverlapping(l_s, r_e, l_idx, ix, k, ties=None):
    """l_s: sorted query starts, r_e: sorted subject ends, ix[i]: position of the last end <= l_s[i].
    -> (query labels, positions into r_e, l_s - r_e) walking leftwards from ix[i]."""
    l_s, r_e, l_idx, ix = _a64(l_s), _a64(r_e), np.asarray(l_idx), _a64(ix)
    lo, ro, do = [], [], []
    for i in range(len(l_s)):
        pos = list(range(int(ix[i]), -1, -1))
        keep = _k_walk([int(r_e[p]) for p in pos], pos, _k_of(k, l_idx, i), ties)
        for p in keep:
            lo.append(l_idx[i])
            ro.append(p)
            do.append(int(l_s[i] - r_e[p]))
    return np.asarray(lo, dtype=np.int64), np.asarray(ro, dtype=np.int64), np.asarray(do, dtype=np.int64)


def k_nearest_next_nonoverlapping(l_e, r_s, l_idx, ix, k, ties=None):
    """l_e: sorted query ends, r_s: sorted subject starts, ix[i]: position of the first start >= l_e[i].
    -> (query labels, positions into r_s, r_s - l_e) walking rightwards from ix[i]."""
    l_e, r_s, l_idx, ix = _a64(l_e), _a64(r_s), np.asarray(l_idx), _a64(ix)
    lo, ro, do = [], [], []
    for i in range(len(l_e)):
        pos = list(range(int(ix[i]), len(r_s)))
        keep = _k_walk([int(r_s[p]) for p in pos], pos, _k_of(k, l_idx, i), ties)
        for p in keep:
            lo.append(l_idx[i])
            ro.append(p)
            do.append(int(r_s[p] - l_e[i]))
    return np.asarray(lo, dtype=np.int64), np.asarray(ro, dtype=np.int64), np.asarray(do, dtype=np.int64)


def get_all_ties(index, ix, dist, k):
    """rows (labels `index`, sorted by (ix, dist)) of every query ix up to its k-th row plus the rows tied with it."""
    index, ix, dist = np.asarray(index), _a64(ix), _a64(dist)
    out, i, n = [], 0, len(ix)
    while i < n:
        j = i
        while j < n and ix[j] == ix[i]:
            j += 1
        m = j - i
        if k > 0:
            cut = dist[i + min(k, m) - 1]
            out.extend(index[t] for t in range(i, j) if dist[t] <= cut)
        i = j
    return np.asarray(out, dtype=index.dtype if len(out) else np.int64)


def get_different_ties(index, ix, dist, k):
    """rows of every query ix whose distance is among its k smallest distinct distances."""
    index, ix, dist = np.asarray(index), _a64(ix), _a64(dist)
    out, i, n = [], 0, len(ix)
    while i < n:
        j = i
        while j < n and ix[j] == ix[i]:
            j += 1
        distinct, last = 0, None
        for t in range(i, j):
            if last is None or dist[t] != last:
                distinct += 1
                last = dist[t]
            if distinct > k:
                break
            out.append(index[t])
        i = j
    return np.asarray(out, dtype=index.dtype if len(out) else np.int64)
