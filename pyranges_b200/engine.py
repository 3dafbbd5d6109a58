"""Batched interval engine: device-resident, key-segmented int64 columns + the C-ABI kernels.

This is the layer that replaces the reference's per-key calls into ncls / sorted_nearest / pyrle
(pyranges/methods/join.py:12-23, intersection.py:71-78, coverage.py:20-26,58-64, nearest.py:59-113,
cluster.py:13, merge.py:14, to_rle.py:18).  Where the reference loops over (chromosome, strand) keys
and builds one NCLS per key (pyranges/multithreaded.py:221-294), here the rows of all keys sit in one
concatenated int64 column pair in HBM (natsorted key order) and ONE launch sequence serves every key.

torch is used for device memory and streams only; every computation is a kernel of
libpyranges_b200.so reached through ctypes (pyranges_b200/_cabi.py).  No CPU fallback exists.
"""
import ctypes as C

import numpy as np
import torch

from . import _cabi as cabi
from ._cabi import (IV_INDEX_ENDS_PERM, IV_MODE_ALL, IV_MODE_ANY, IV_MODE_CONTAIN, IV_MODE_LAST, IV_NEAREST_BOTH,  # noqa: F401
                    IV_NEAREST_NEXT, IV_NEAREST_PREVIOUS, IvGroups, IvIndex, IvQueries)

I64_MAX = np.iinfo(np.int64).max
I64_MIN = np.iinfo(np.int64).min


def _ptr(t):
    return C.c_void_p(t.data_ptr()) if t is not None else None


def _stream():
    return C.c_void_p(torch.cuda.current_stream().cuda_stream)


def _use(device):
    """Make `device` current for this thread in the library's CUDA runtime as well as torch's."""
    device = torch.device(device)
    if device.type != "cuda":
        raise cabi.CabiError("pyranges_b200 runs on CUDA devices only (got %r); there is no CPU path" % (device,))
    idx = device.index if device.index is not None else torch.cuda.current_device()
    cabi.check(cabi.lib().iv_set_device(idx), "iv_set_device")
    return torch.device("cuda", idx)


def bits_for(v):
    v = int(v)
    b = 1
    while b < 64 and (v >> b):
        b += 1
    return b


def _bytes(nbytes, device):
    return torch.empty(max(int(nbytes), 256), dtype=torch.uint8, device=device)


def _to_dev(a, device):
    return torch.from_numpy(np.ascontiguousarray(a)).to(device, non_blocking=False)


_desc_cache = {}
_groups_cache = {}


def _desc_to_dev(a, device):
    """Device copy of a SMALL read-only descriptor array (segment offsets, group table, partner map), memoised by
    content: the same chromosome table comes back call after call, and a pageable host->device copy costs a stream
    synchronisation each time (and queues behind bulk transfers on the copy engine)."""
    a = np.ascontiguousarray(a)
    if a.nbytes > (1 << 16):
        return _to_dev(a, device)
    key = (str(device), a.dtype.str, a.tobytes())
    t = _desc_cache.get(key)
    if t is None:
        if len(_desc_cache) >= 512:
            _desc_cache.clear()
        t = _desc_cache[key] = _to_dev(a, device)
    return t


class DeviceIntervals:
    """Key-segmented interval columns on one GPU.

    start, end : int64 device tensors, rows of all keys concatenated (segment k = rows
                 [seg_off[k], seg_off[k+1])); mirrors df.Start.values / df.End.values of every
                 DataFrame of PyRanges.dfs (pyranges/methods/init.py:35-45) laid end to end.
    lo, hi, maxlen : per-segment min Start / max End / max length, computed once at ingest.
    """

    def __init__(self, start, end, seg_off, bounds=None, seg_off_dev=None):
        assert start.dtype == torch.int64 and end.dtype == torch.int64 and start.is_cuda and end.is_cuda
        self.device = _use(start.device)
        self.start, self.end = start.contiguous(), end.contiguous()
        self.n = int(start.numel())
        self.seg_off = np.ascontiguousarray(seg_off, dtype=np.int64)
        assert self.seg_off[0] == 0 and self.seg_off[-1] == self.n
        self.n_seg = len(self.seg_off) - 1
        # (a caller that pipelines uploads passes the device copy: a small pageable copy issued now would queue
        # behind its bulk transfers on the copy engine)
        self.seg_off_dev = seg_off_dev if seg_off_dev is not None else _desc_to_dev(self.seg_off, self.device)
        # (lo, hi, maxlen, sumlen[, minlen]) per segment; computed on first use (subjects / unary ops)
        self._b = bounds if bounds is None or len(bounds) == 5 else tuple(bounds) + (np.ones(self.n_seg, np.int64),)
        self._groups = {}

    def groups(self, seg_ranges=None, gap=1, zero_lo=False, widen=False):
        """iv_groups_t descriptor of these rows (ingest-time metadata: memoised per grouping).  widen: windows reach
        one coordinate beyond the rows on both sides (list index over zero-length subjects)."""
        key = (None if seg_ranges is None else tuple(map(tuple, seg_ranges)), int(gap), bool(zero_lo), bool(widen))
        g = self._groups.get(key)
        if g is None:
            # the descriptor only depends on the segment table and its bounds, not on the rows: the same chromosome
            # table comes back with every new upload of the same frame (content-keyed, small)
            ckey = None
            if self.n_seg <= 4096:
                ckey = (key, str(self.device), self.seg_off.tobytes()) + tuple(np.ascontiguousarray(b).tobytes() for b in self.bounds)
                g = _groups_cache.get(ckey)
            if g is None:
                g = Groups(self, seg_ranges, gap=gap, zero_lo=zero_lo, widen=widen)
                if ckey is not None:
                    if len(_groups_cache) >= 128:
                        _groups_cache.clear()
                    _groups_cache[ckey] = g
            self._groups[key] = g
        return g

    @property
    def bounds(self):
        if self._b is None:
            self._b = self._bounds()
        return self._b

    @property
    def lo(self):
        return self.bounds[0]

    @property
    def hi(self):
        return self.bounds[1]

    @property
    def maxlen(self):
        return self.bounds[2]

    @property
    def position_ordered(self):
        """True when the rows come (mostly) in Start order inside their segments, e.g. a BAM-sorted read set: judged
        once per upload from a strided sample of adjacent row pairs (ingest-time metadata, like the bounds)."""
        v = self.__dict__.get("_ordered")
        if v is None:
            v = False
            if self.n >= 4096:
                step = max(1, (self.n - 1) // (1 << 20))
                a = self.start[0:self.n - 1:step]
                b = self.start[1:self.n:step]
                v = bool(float((b >= a).float().mean()) >= 0.9)
            self.__dict__["_ordered"] = v
        return v

    @property
    def sumlen(self):
        return self.bounds[3]

    @property
    def minlen(self):
        return self.bounds[4]

    @property
    def has_zero_length(self):
        """any row with Start == End"""
        m = self.minlen
        return bool(len(m)) and bool((m[np.diff(self.seg_off) > 0] <= 0).any())

    @classmethod
    def from_numpy(cls, starts, ends, seg_off, device="cuda", pinned=False):
        dev = _use(device)
        s = torch.from_numpy(np.ascontiguousarray(starts, dtype=np.int64))
        e = torch.from_numpy(np.ascontiguousarray(ends, dtype=np.int64))
        return cls(s.to(dev), e.to(dev), seg_off)

    def _bounds(self):
        k = self.n_seg
        if k == 0:
            z = np.zeros(0, np.int64)
            return z, z.copy(), z.copy(), z.copy(), z.copy()
        out = torch.empty(5 * k, dtype=torch.int64, device=self.device)
        p = out.data_ptr()
        cabi.check(cabi.lib().iv_segment_bounds(_ptr(self.start), _ptr(self.end), _ptr(self.seg_off_dev), k,
                                                C.c_void_p(p), C.c_void_p(p + 8 * k), C.c_void_p(p + 16 * k),
                                                C.c_void_p(p + 24 * k), C.c_void_p(p + 32 * k), _stream()),
                   "iv_segment_bounds")
        h = out.cpu().numpy()
        return tuple(h[i * k:(i + 1) * k].copy() for i in range(5))


class Groups:
    """iv_groups_t: contiguous row groups (partner frames) with their flattened-coordinate origins."""

    def __init__(self, iv, seg_ranges=None, gap=1, zero_lo=False, widen=False):
        k = iv.n_seg
        if seg_ranges is None:
            seg_ranges = [(i, i + 1) for i in range(k)]
        a = np.asarray([r[0] for r in seg_ranges], dtype=np.int64)
        b = np.asarray([r[1] for r in seg_ranges], dtype=np.int64)
        g = len(seg_ranges)
        if g:
            assert a[0] == 0 and b[-1] == k and np.all(a[1:] == b[:-1]), "groups must tile the segments in order"
        self.n_groups = g
        self.row_off = np.concatenate([iv.seg_off[a], iv.seg_off[b[-1:]]]) if g else np.zeros(1, np.int64)
        lo = np.full(g, I64_MAX, np.int64)
        hi = np.full(g, I64_MIN, np.int64)
        ml = np.zeros(g, np.int64)
        if g and k:
            # per-group reduce of the per-segment bounds (segments of a group are consecutive)
            lo = np.minimum.reduceat(iv.lo, a)
            hi = np.maximum.reduceat(iv.hi, a)
            ml = np.maximum.reduceat(iv.maxlen, a)
        empty = lo > hi
        lo = np.where(empty, 0, lo)
        hi = np.where(empty, 0, hi)
        if widen:
            lo, hi = lo - 1, hi + 1
        if zero_lo:
            if g and lo.min() < 0:
                raise ValueError("coverage needs non-negative coordinates")
            lo = np.zeros(g, np.int64)
        self.lo, self.hi, self.maxlen = lo, hi, int(ml.max()) if g else 0
        self.sumlen = int(iv.sumlen.sum()) if k else 0
        base = np.zeros(g + 1, np.int64)
        np.cumsum((hi - lo) + int(gap), out=base[1:])
        self.base = base
        self.total_span = int(base[-1])
        packed = np.concatenate([self.row_off, lo, hi, base]).astype(np.int64)
        self.dev = _desc_to_dev(packed, iv.device)
        p = self.dev.data_ptr()
        self.c = IvGroups(g, 0, p, p + 8 * (g + 1), p + 8 * (2 * g + 1), p + 8 * (3 * g + 1), self.total_span,
                          self.sumlen, self.maxlen)


class Queries:
    """iv_queries_t over a DeviceIntervals: every segment names its partner group (-1 = none)."""

    def __init__(self, iv, seg_group, seg_mode=None, slack=0, label=None, desc_dev=None):
        self.iv = iv
        sg = np.ascontiguousarray(seg_group, dtype=np.int32)
        assert len(sg) == iv.n_seg
        arr = sg if seg_mode is None else np.concatenate([sg, np.ascontiguousarray(seg_mode, dtype=np.int32)])
        # the partner map is a few hundred bytes; its device copy is memoised on the columns (a pageable
        # host->device copy would otherwise synchronise the stream on every call) or handed in (desc_dev)
        key = arr.tobytes()
        cache = iv.__dict__.setdefault("_qdesc", {})
        if desc_dev is not None:
            cache[key] = desc_dev
        if key not in cache:
            cache[key] = _desc_to_dev(arr, iv.device) if len(arr) else torch.zeros(1, dtype=torch.int32, device=iv.device)
        self.dev = cache[key]
        p = self.dev.data_ptr()
        self.label = label
        self.c = IvQueries(iv.n, iv.start.data_ptr(), iv.end.data_ptr(), label.data_ptr() if label is not None else None,
                           iv.n_seg, 0, iv.seg_off_dev.data_ptr(), p,
                           (p + 4 * iv.n_seg) if seg_mode is not None else None, int(slack))


def _max_len(q):
    """longest query of a Queries object, slack included (ingest-time metadata of its columns)"""
    ml = q.iv.maxlen
    return (int(ml.max()) if len(ml) else 0) + 2 * int(q.c.slack)


class SubjectIndex:
    """Device RANK index over the subject ("other") intervals of ALL groups -- the NCLS replacement
    (pyranges/methods/join.py:12) behind nearest, coverage, subtract, containment and first / last overlap.  Built by
    iv_index_build: radix sorts of the flattened starts and ends, 32-byte rank records per coordinate bin, per-bin
    stabbing lists.  (join / count_overlaps / overlap use ListIndex.)"""

    def __init__(self, iv, seg_ranges=None, ends_perm=False, graph=True):
        """graph: let repeated builds over the same buffers replay a captured launch sequence (IV_INDEX_GRAPH)"""
        self.iv = iv
        if iv.has_zero_length:
            # the rank index counts with #{S < e} - #{E <= s}, which needs S < E; join / count_overlaps / overlap go
            # through ListIndex, which handles such rows exactly
            raise ValueError("zero-length intervals (Start == End) in `other` are not supported by this operation")
        self.groups = iv.groups(seg_ranges, gap=1)
        self.flags = (IV_INDEX_ENDS_PERM if ends_perm else 0) | (cabi.IV_INDEX_GRAPH if graph else 0)
        L = cabi.lib()
        nbytes = L.iv_index_workspace_bytes(iv.n, self.groups.total_span, self.groups.sumlen, self.flags)
        self.ws = _bytes(nbytes, iv.device)
        self.c = IvIndex()
        self.build()

    def build(self):
        """(re)run the device build into this index' workspace: sorts, bin records, stabbing lists"""
        iv = self.iv
        cabi.check(cabi.lib().iv_index_build(_ptr(iv.start), _ptr(iv.end), iv.n, C.byref(self.groups.c), self.flags,
                                             _ptr(self.ws), self.ws.numel(), C.byref(self.c), _stream()),
                   "iv_index_build")
        return self

    # -- counts ------------------------------------------------------------------------------------
    def count(self, q, mode=IV_MODE_ALL, for_emit=True, want_counts=True):
        """-> (counts int64[nq] | None, ws, total int64[1] device).  NCLS count / has_overlaps /
        containment count per query (pyranges/methods/coverage.py:26-40, intersection.py:73-78).  for_emit: also
        leave the hit lists and tile offsets iv_emit_pairs works from in `ws`; the per-row counts themselves are
        then optional (want_counts=False saves their 8 bytes per row)."""
        L = cabi.lib()
        dev = self.iv.device
        assert for_emit or want_counts
        counts = torch.empty(q.iv.n, dtype=torch.int64, device=dev) if want_counts else None
        ws = _bytes(L.iv_overlap_workspace_bytes(q.iv.n), dev) if for_emit else None
        total = torch.zeros(1, dtype=torch.int64, device=dev) if for_emit else None
        cabi.check(L.iv_count_overlaps(C.byref(self.c), C.byref(q.c), mode, _ptr(counts), _ptr(ws),
                                       ws.numel() if ws is not None else 0, _ptr(total), _stream()),
                   "iv_count_overlaps")
        return counts, ws, total

    def segment_pair_counts(self, q, ws, total):
        """Pairs per query segment (key), from what count(for_emit=True) left in `ws`: available before the emit, so
        that the per-key result counts can be exchanged between GPUs while the pairs are written."""
        out = torch.empty(max(q.iv.n_seg, 1), dtype=torch.int64, device=self.iv.device)
        cabi.check(cabi.lib().iv_segment_pair_counts(C.byref(self.c), C.byref(q.c), _ptr(ws), _ptr(total), _ptr(out),
                                                     _stream()), "iv_segment_pair_counts")
        return out[:q.iv.n_seg]

    def pairs(self, q, mode=IV_MODE_ALL, want_q=True, want_s=True, s_label=None, capacity=None, want_counts=True):
        """All (query row | label, subject row | label) pairs in query order: count -> scan -> emit
        (NCLS.all_overlaps_both / all_containments_both / first_overlap_both, join.py:15-23).

        The output size is only known after the count, which costs one host sync before the emit can be
        launched.  `capacity` (e.g. the pair count of a previous, similar join) lets the emit be launched
        speculatively into buffers of that size right behind the count; the total is read afterwards and the
        emit is simply repeated into exact buffers if the guess was too small."""
        counts, ws, total = self.count(q, mode, for_emit=True, want_counts=want_counts)
        dev = self.iv.device

        def emit(n):
            oq = torch.empty(n, dtype=torch.int64, device=dev) if want_q else None
            os_ = torch.empty(n, dtype=torch.int64, device=dev) if want_s else None
            cabi.check(cabi.lib().iv_emit_pairs(C.byref(self.c), C.byref(q.c), mode, None, _ptr(ws),
                                                _ptr(oq), _ptr(os_), _ptr(s_label), n, _stream()), "iv_emit_pairs")
            return oq, os_

        if capacity is not None and capacity > 0:
            out_q, out_s = emit(int(capacity))
            p = int(total.item())
            if p <= capacity:
                return counts, (out_q[:p] if want_q else None), (out_s[:p] if want_s else None)
            # guess too small: the kernel dropped the entries beyond `capacity`; emit again into exact buffers
        p = int(total.item())  # the one host sync: output size
        if p == 0:
            z = torch.empty(0, dtype=torch.int64, device=dev)
            return counts, (z if want_q else None), (z.clone() if want_s else None)
        out_q, out_s = emit(p)
        return counts, out_q, out_s

    def coverage_bp(self, q, want_fraction=True):
        """NCLS.coverage (pyranges/methods/coverage.py:64-68): covered bp and bp / (End - Start)."""
        dev = self.iv.device
        bp = torch.empty(q.iv.n, dtype=torch.int64, device=dev)
        fr = torch.empty(q.iv.n, dtype=torch.float64, device=dev) if want_fraction else None
        cabi.check(cabi.lib().iv_coverage_bp(C.byref(self.c), C.byref(q.c), _ptr(bp), _ptr(fr), _stream()),
                   "iv_coverage_bp")
        return bp, fr

    def subtract(self, q):
        """NCLS.set_difference_helper (pyranges/methods/subtraction.py:63-65); the index must hold disjoint
        (merged) subjects.  -> (query row, start, end) of every surviving piece, query order."""
        L = cabi.lib()
        dev = self.iv.device
        off = torch.empty(max(q.iv.n, 1), dtype=torch.int64, device=dev)
        ws = _bytes(L.iv_subtract_workspace_bytes(q.iv.n), dev)
        total = torch.zeros(1, dtype=torch.int64, device=dev)
        cabi.check(L.iv_subtract_count(C.byref(self.c), C.byref(q.c), _ptr(off), _ptr(ws), ws.numel(), _ptr(total),
                                       _stream()), "iv_subtract_count")
        m = int(total.item())
        oq = torch.empty(m, dtype=torch.int64, device=dev)
        os_ = torch.empty(m, dtype=torch.int64, device=dev)
        oe = torch.empty(m, dtype=torch.int64, device=dev)
        if m:
            cabi.check(L.iv_subtract_emit(C.byref(self.c), C.byref(q.c), _ptr(off), _ptr(oq), _ptr(os_), _ptr(oe),
                                          _stream()), "iv_subtract_emit")
        return oq, os_, oe

    def nearest(self, q, overlap=True):
        """-> (subject row or -1, distance) per query (pyranges/methods/nearest.py:29-53,59,69,113)."""
        assert self.flags & IV_INDEX_ENDS_PERM, "build the index with ends_perm=True for nearest"
        dev = self.iv.device
        row = torch.empty(q.iv.n, dtype=torch.int64, device=dev)
        dist = torch.empty(q.iv.n, dtype=torch.int64, device=dev)
        cabi.check(cabi.lib().iv_nearest(C.byref(self.c), C.byref(q.c), 1 if overlap else 0, _ptr(row), _ptr(dist),
                                         _stream()), "iv_nearest")
        return row, dist

    def k_nearest(self, q, k, ties=None):
        """k_nearest_previous_nonoverlapping + k_nearest_next_nonoverlapping (pyranges/methods/k_nearest.py:6-48) for
        all keys: k = int64 device tensor, one entry per query row; ties in (None, "first", "last", "different").
        -> (query row, subject row, distance >= 1) per candidate, query order, previous candidates before next ones."""
        assert self.flags & IV_INDEX_ENDS_PERM, "build the index with ends_perm=True for k_nearest"
        L = cabi.lib()
        dev = self.iv.device
        code = {None: 0, False: 0, "first": 1, "last": 2, "different": 3}[ties]
        assert k.dtype == torch.int64 and k.numel() == q.iv.n and k.is_cuda
        k = k.contiguous()
        off = torch.empty(max(q.iv.n, 1), dtype=torch.int64, device=dev)
        ws = _bytes(L.iv_k_nearest_workspace_bytes(q.iv.n), dev)
        total = torch.zeros(1, dtype=torch.int64, device=dev)
        cabi.check(L.iv_k_nearest_count(C.byref(self.c), C.byref(q.c), _ptr(k), code, _ptr(off), _ptr(ws), ws.numel(),
                                        _ptr(total), _stream()), "iv_k_nearest_count")
        m = int(total.item())
        oq = torch.empty(m, dtype=torch.int64, device=dev)
        os_ = torch.empty(m, dtype=torch.int64, device=dev)
        od = torch.empty(m, dtype=torch.int64, device=dev)
        if m:
            cabi.check(L.iv_k_nearest_emit(C.byref(self.c), C.byref(q.c), _ptr(k), code, _ptr(off), _ptr(oq), _ptr(os_),
                                           _ptr(od), _stream()), "iv_k_nearest_emit")
        return oq, os_, od


def to_device_i64(a, device):
    """int64 host array -> device tensor (per-row parameters of an operation, e.g. the k of k_nearest)."""
    return torch.from_numpy(np.ascontiguousarray(a, dtype=np.int64)).to(_use(device))


# ---- unary operations -------------------------------------------------------------------------------
def _cluster_groups(iv, seg_ranges, slack):
    g = iv.groups(seg_ranges, gap=max(int(slack), 0) + 1)
    return g, bits_for(g.maxlen)


def cluster(iv, seg_ranges=None, slack=0):
    """sort_values("Start") + annotate_clusters + global id offsets (pyranges/methods/cluster.py:11-15,
    pyranges/pyranges_main.py:1210-1223) for all groups at once.
    -> (row int64[n]: input row at each sorted position, id int64[n]: 1-based global cluster id, n_clusters)."""
    L = cabi.lib()
    g, lbits = _cluster_groups(iv, seg_ranges, slack)
    dev = iv.device
    ws = _bytes(L.iv_cluster_workspace_bytes(iv.n), dev)
    row = torch.empty(iv.n, dtype=torch.int64, device=dev)
    cid = torch.empty(iv.n, dtype=torch.int64, device=dev)
    nc = torch.zeros(1, dtype=torch.int64, device=dev)
    cabi.check(L.iv_cluster(_ptr(iv.start), _ptr(iv.end), iv.n, C.byref(g.c), int(slack), lbits, _ptr(row), _ptr(cid),
                            _ptr(nc), _ptr(ws), ws.numel(), _stream()), "iv_cluster")
    return row, cid, nc


def merge(iv, seg_ranges=None, slack=0, want_count=True):
    """sort_values("Start") + find_clusters (pyranges/methods/merge.py:12-14) for all groups at once.
    -> (start, end, count | None, group int32) one row per cluster, group-major, Start order."""
    L = cabi.lib()
    g, lbits = _cluster_groups(iv, seg_ranges, slack)
    dev = iv.device
    ws = _bytes(L.iv_cluster_workspace_bytes(iv.n), dev)
    nc = torch.zeros(1, dtype=torch.int64, device=dev)
    cabi.check(L.iv_merge_count(_ptr(iv.start), _ptr(iv.end), iv.n, C.byref(g.c), int(slack), lbits, _ptr(nc),
                                _ptr(ws), ws.numel(), _stream()), "iv_merge_count")
    m = int(nc.item())
    st = torch.empty(m, dtype=torch.int64, device=dev)
    en = torch.empty(m, dtype=torch.int64, device=dev)
    cnt = torch.empty(m, dtype=torch.int64, device=dev) if want_count else None
    grp = torch.empty(m, dtype=torch.int32, device=dev)
    cabi.check(L.iv_merge_emit(iv.n, m, C.byref(g.c), lbits, _ptr(ws), _ptr(st), _ptr(en), _ptr(cnt), _ptr(grp),
                               _stream()), "iv_merge_emit")
    return st, en, cnt, grp


def split(iv, seg_ranges=None):
    """PyRanges.split(between=True) (pyranges/methods/split.py:4-24) for all groups at once: every pair of consecutive
    distinct Start/End points of a group.  -> (group_off int64[n_groups+1] host, start, end) group-major."""
    L = cabi.lib()
    g = iv.groups(seg_ranges, gap=1)
    dev = iv.device
    ws = _bytes(L.iv_split_workspace_bytes(iv.n), dev)
    off = torch.zeros(g.n_groups + 1, dtype=torch.int64, device=dev)
    cabi.check(L.iv_split_count(_ptr(iv.start), _ptr(iv.end), iv.n, C.byref(g.c), _ptr(ws), ws.numel(), _ptr(off),
                                _stream()), "iv_split_count")
    off = off.cpu().numpy()
    m = int(off[-1])
    st = torch.empty(m, dtype=torch.int64, device=dev)
    en = torch.empty(m, dtype=torch.int64, device=dev)
    cabi.check(L.iv_split_emit(iv.n, C.byref(g.c), _ptr(ws), m, _ptr(st), _ptr(en), None, _stream()), "iv_split_emit")
    return off, st, en


def coverage_rle(iv, seg_ranges=None, weights=None):
    """pyrle.methods.coverage (pyranges/methods/to_rle.py:18) for all groups at once; weights = float64 device
    tensor aligned with the rows (value_col), None = unit weights (bit-exact integer levels).
    -> (run_off int64[n_groups+1] host, runs int64[R], values float64[R])."""
    L = cabi.lib()
    g = iv.groups(seg_ranges, gap=1, zero_lo=True)
    dev = iv.device
    run_off = torch.zeros(g.n_groups + 1, dtype=torch.int64, device=dev)
    if weights is None:
        ws = _bytes(L.iv_rle_workspace_bytes(iv.n, g.n_groups), dev)
        cabi.check(L.iv_rle_count(_ptr(iv.start), _ptr(iv.end), iv.n, C.byref(g.c), _ptr(ws), ws.numel(), _ptr(run_off),
                                  _stream()), "iv_rle_count")
    else:
        assert weights.dtype == torch.float64 and weights.numel() == iv.n and weights.is_cuda
        weights = weights.contiguous()
        ws = _bytes(L.iv_rle_weighted_workspace_bytes(iv.n, g.n_groups), dev)
        cabi.check(L.iv_rle_weighted_count(_ptr(iv.start), _ptr(iv.end), _ptr(weights), iv.n, C.byref(g.c), _ptr(ws),
                                           ws.numel(), _ptr(run_off), _stream()), "iv_rle_weighted_count")
    off_h = run_off.cpu().numpy()
    r = int(off_h[-1])
    runs = torch.empty(r, dtype=torch.int64, device=dev)
    vals = torch.empty(r, dtype=torch.float64, device=dev)
    emit = L.iv_rle_emit if weights is None else L.iv_rle_weighted_emit
    cabi.check(emit(iv.n, C.byref(g.c), _ptr(ws), r, _ptr(run_off), _ptr(runs), _ptr(vals), _stream()), "iv_rle_emit")
    return off_h, runs, vals


def radix_sort_u64(keys, values=None, begin_bit=0, end_bit=64):
    """In-place stable LSD radix sort of a uint64 (viewed as int64 tensor) key column (+ optional payload)."""
    L = cabi.lib()
    dev = _use(keys.device)
    vb = 0 if values is None else values.element_size()
    ws = _bytes(L.iv_sort_workspace_bytes(keys.numel(), vb), dev)
    cabi.check(L.iv_radix_sort_u64(_ptr(keys), _ptr(values), vb, keys.numel(), begin_bit, end_bit, _ptr(ws), ws.numel(),
                                   _stream()), "iv_radix_sort_u64")
    return keys, values


def materialize_pairs(q_iv, s_iv, pair_q, pair_s, clip=False, want=("q_start", "q_end", "s_start", "s_end"),
                      q_row_offset=0, s_row_offset=0):
    """Coordinate columns of the joined frame for the pair lists of SubjectIndex.pairs, gathered on the device
    (pyranges/methods/join.py:95-123,147-151; clip=True: intersect, intersection.py:39-48).  `want` names the
    outputs: q_start, q_end, s_start, s_end, overlap.  -> dict name -> int64 tensor [n_pairs]."""
    L = cabi.lib()
    dev = _use(q_iv.device)
    n = int(pair_q.numel())
    out = {k: torch.empty(n, dtype=torch.int64, device=dev) for k in want}
    g = lambda k: _ptr(out[k]) if k in out else None  # noqa: E731
    cabi.check(L.iv_materialize_pairs(_ptr(q_iv.start), _ptr(q_iv.end), _ptr(s_iv.start), _ptr(s_iv.end),
                                      _ptr(pair_q), _ptr(pair_s), n, q_row_offset, s_row_offset, 1 if clip else 0,
                                      g("q_start"), g("q_end"), g("s_start"), g("s_end"), g("overlap"), _stream()),
               "iv_materialize_pairs")
    return out


def gather(column, rows, fill=None):
    """out[i] = column[rows[i]] on the device for a 1/2/4/8-byte column (payload columns of the joined frame);
    rows < 0 give `fill` (a Python scalar of the column's dtype, default 0)."""
    L = cabi.lib()
    dev = _use(column.device)
    out = torch.empty(rows.numel(), dtype=column.dtype, device=dev)
    fb = None
    if fill is not None:
        fb = torch.tensor([fill], dtype=column.dtype).numpy().tobytes()
    cabi.check(L.iv_gather(_ptr(column), column.element_size(), _ptr(rows), rows.numel(), _ptr(out), fb, _stream()),
               "iv_gather")
    return out


# ---- path selection: list index (fused join) where it applies, rank index otherwise --------------------------------------
def _mean_len(q_iv):
    return float(q_iv.sumlen.sum()) / q_iv.n if q_iv.n else 0.0


def overlap_pairs(s_iv, group_ranges, q, want_counts=False, capacity=None):
    """NCLS.all_overlaps_both for all keys (pyranges/methods/join.py:15,23): -> (counts | None, query rows, subject
    rows), pairs in query order."""
    if ListIndex.supported(s_iv, group_ranges, _mean_len(q.iv)):
        ix = ListIndex(s_iv, group_ranges, reach=_max_len(q))
        a, b = ix.pairs(q, capacity=capacity)
        return (ix.count(q, IV_MODE_ALL) if want_counts else None), a, b
    return SubjectIndex(s_iv, group_ranges).pairs(q, IV_MODE_ALL, want_counts=want_counts, capacity=capacity)


def overlap_counts(s_iv, group_ranges, q, mode=IV_MODE_ALL):
    """overlaps per query row: all (IV_MODE_ALL), 0/1 (IV_MODE_ANY) or containments (IV_MODE_CONTAIN)"""
    if mode in (IV_MODE_ALL, IV_MODE_ANY) and ListIndex.supported(s_iv, group_ranges, _mean_len(q.iv)):
        return ListIndex(s_iv, group_ranges, reach=_max_len(q)).count(q, mode)
    return SubjectIndex(s_iv, group_ranges).count(q, mode, for_emit=False)[0]


# ---- repeated operations on resident columns ------------------------------------------------------------------------
class _Totals:
    """Device-side result sizes of asynchronous steps: every step copies its total into a pinned host slot without
    synchronising; check() waits once and validates all of them against the capacity the step emitted into."""

    def __init__(self):
        self.slots = []

    def push(self, total_dev, cap):
        h = torch.empty(1, dtype=torch.int64).pin_memory()
        h.copy_(total_dev, non_blocking=True)
        self.slots.append((h, cap))

    def check(self, what):
        torch.cuda.current_stream().synchronize()
        worst = 0
        for h, cap in self.slots:
            p = int(h.item())
            if p > cap:
                raise cabi.CabiError("%s: a step produced %d entries for buffers of %d" % (what, p, cap))
            worst = max(worst, p)
        self.slots = []
        return worst


class JoinStep:
    """gr.join(other) over and over on the same resident columns, the way a pipeline (or a benchmark) calls it: every
    step REBUILDS the subject index (the reference builds a new NCLS per call, pyranges/methods/join.py:12) and writes
    all (query row, subject row) pairs in query order.  Output buffers are sized by the first, synchronous pass
    (run_sync); later steps launch everything without a host synchronisation and leave their pair total in a pinned
    slot that check() validates afterwards (a step whose pairs did not fit raises there).

    path = "twopass": iv_index_build -> iv_count_overlaps (hit lists) -> iv_emit_pairs
    path = "fused"  : iv_lists_build -> iv_join_pairs (one kernel: count, look-back offsets, staged emission)"""

    def __init__(self, s_iv, group_ranges, q_iv, seg_group, path="auto", slack=0, seg_counts=False):
        self.s_iv, self.q_iv, self.group_ranges = s_iv, q_iv, group_ranges
        self.q = Queries(q_iv, seg_group, slack=slack)
        self.reach = _max_len(self.q)  # ingest-time metadata of the query columns
        self.want_seg_counts = seg_counts  # also leave the pairs per query key (device) in self.seg_counts
        self.seg_counts = None
        if path == "auto":
            path = "fused" if ListIndex.supported(s_iv, group_ranges, _mean_len(q_iv)) else "twopass"
        self.path = path
        self.cap = 0
        self.totals = _Totals()
        self.phase_events = None
        self.out = None
        self._graph = None
        self.graph_replays = 0
        self.graph_launches = 0

    def _ev(self):
        e = torch.cuda.Event(enable_timing=True)
        e.record()
        return e

    def _run(self, cap):
        dev = self.q_iv.device
        ev = [self._ev()] if self.phase_events is not None else None
        out_q = torch.empty(cap, dtype=torch.int64, device=dev)
        out_s = torch.empty(cap, dtype=torch.int64, device=dev)
        if self.path == "fused":
            ix = ListIndex(self.s_iv, self.group_ranges, reach=self.reach)
            if ev:
                ev.append(self._ev())
            total = ix.join_pairs(self.q, out_q, out_s, cap)
            if self.want_seg_counts:
                self.seg_counts = ix.segment_pair_counts(self.q, out_q, total, cap)
            if ev:
                ev.append(self._ev())
        else:
            ix = SubjectIndex(self.s_iv, self.group_ranges)
            if ev:
                ev.append(self._ev())
            _, ws, total = ix.count(self.q, IV_MODE_ALL, want_counts=False)
            if self.want_seg_counts:
                self.seg_counts = ix.segment_pair_counts(self.q, ws, total)
            if ev:
                ev.append(self._ev())
            cabi.check(cabi.lib().iv_emit_pairs(C.byref(ix.c), C.byref(self.q.c), IV_MODE_ALL, None, _ptr(ws), _ptr(out_q),
                                                _ptr(out_s), None, cap, _stream()), "iv_emit_pairs")
            if ev:
                ev.append(self._ev())
        if ev:
            self.phase_events.append(ev)
        return out_q, out_s, total

    def run_sync(self):
        """-> (query rows, subject rows, number of pairs); sizes the buffers of the following step() calls"""
        cap = self.cap or max(2 * self.q_iv.n, 1 << 20)
        while True:
            out_q, out_s, total = self._run(cap)
            p = int(total.item())
            if p <= cap:
                break
            cap = p + 4096
        self.cap = int(p * 1.02) + 4096
        return out_q[:p], out_s[:p], p

    def step(self):
        if not self.cap:
            return self.run_sync()
        if self._graph is not None:
            self._graph.replay()
            self.graph_replays += 1
            out_q, out_s, total = self.out
        else:
            out_q, out_s, total = self._run(self.cap)
        self.totals.push(total, self.cap)
        self.out = (out_q, out_s, total)
        return self.out

    def capture(self):
        """Capture the whole step (index build, fused join, per-key counts) into ONE CUDA graph; later step() calls
        replay it: one launch per step instead of a dozen, no Python between the kernels.  The outputs then live in
        fixed buffers that every replay overwrites.  Explicit opt-in (bench.py uses it for the short per-rank steps
        of a multi-GPU run); needs run_sync() first (buffer sizes) and cannot record per-phase events."""
        assert self.cap, "run_sync() first"
        self.phase_events = None
        L = cabi.lib()
        torch.cuda.synchronize()
        side = torch.cuda.Stream()
        side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(side):  # lazy one-time work (function attributes, descriptor uploads) before the capture
            self._run(self.cap)
        torch.cuda.current_stream().wait_stream(side)
        torch.cuda.synchronize()
        g = torch.cuda.CUDAGraph()
        n0 = L.iv_launch_count()
        with torch.cuda.graph(g):
            self.out = self._run(self.cap)
        self.graph_launches = int(L.iv_launch_count() - n0)  # kernels per replay (the library counts at capture time)
        self._graph = g
        return self

    def check(self):
        return self.totals.check("JoinStep")

    def kernel_of(self, phase):
        names = {"index_build": "iv_lists_build" if self.path == "fused" else "iv_index_build",
                 "join": "k_join" if self.q.iv.position_ordered else "k_join_pipe", "count": "k_count_lean",
                 "emit": "k_emit"}
        return names.get(phase, phase)

    def phase_summary(self, last_n):
        """mean device milliseconds per phase over the last `last_n` recorded steps"""
        names = ["index_build", "join"] if self.path == "fused" else ["index_build", "count", "emit"]
        evs = self.phase_events[-last_n:]
        return {n: float(np.mean([e[i].elapsed_time(e[i + 1]) for e in evs])) for i, n in enumerate(names)}


class CountStep:
    """gr.count_overlaps(other) on resident columns: index rebuild + one int64 count per query row, asynchronous."""

    def __init__(self, s_iv, group_ranges, q_iv, seg_group, mode=IV_MODE_ALL, path="auto"):
        self.s_iv, self.q_iv, self.group_ranges, self.mode = s_iv, q_iv, group_ranges, mode
        self.q = Queries(q_iv, seg_group)
        if path == "auto":
            path = "lists" if ListIndex.supported(s_iv, group_ranges, _mean_len(q_iv)) else "ranks"
        self.path = path
        self.reach = _max_len(self.q)

    def step(self):
        if self.path == "lists":
            return ListIndex(self.s_iv, self.group_ranges, reach=self.reach).count(self.q, self.mode)
        c, _, _ = SubjectIndex(self.s_iv, self.group_ranges).count(self.q, self.mode, for_emit=False)
        return c


class OverlapFirstStep:
    """gr.overlap(other, how="first") on resident columns: the self rows with at least one hit, in row order
    (NCLS.has_overlaps, pyranges/methods/intersection.py:78).  Same asynchronous protocol as JoinStep."""

    def __init__(self, s_iv, group_ranges, q_iv, seg_group, path="auto"):
        self.s_iv, self.q_iv, self.group_ranges = s_iv, q_iv, group_ranges
        self.q = Queries(q_iv, seg_group)
        self.cap = 0
        self.totals = _Totals()
        if path == "auto":
            path = "fused" if ListIndex.supported(s_iv, group_ranges, _mean_len(q_iv)) else "twopass"
        self.path = path
        self.reach = _max_len(self.q)

    def _run(self, cap):
        if self.path == "fused":
            ix = ListIndex(self.s_iv, self.group_ranges, reach=self.reach)
            rows = torch.empty(cap, dtype=torch.int64, device=self.q_iv.device)
            return rows, ix.join_pairs(self.q, rows, None, cap, mode=IV_MODE_ANY)
        ix = SubjectIndex(self.s_iv, self.group_ranges)
        _, ws, total = ix.count(self.q, IV_MODE_ANY, want_counts=False)
        rows = torch.empty(cap, dtype=torch.int64, device=self.q_iv.device)
        cabi.check(cabi.lib().iv_emit_pairs(C.byref(ix.c), C.byref(self.q.c), IV_MODE_ANY, None, _ptr(ws), _ptr(rows), None,
                                            None, cap, _stream()), "iv_emit_pairs")
        return rows, total

    def run_sync(self):
        rows, total = self._run(self.q_iv.n)
        p = int(total.item())
        self.cap = min(self.q_iv.n, int(p * 1.02) + 4096)
        return rows[:p]

    def step(self):
        rows, total = self._run(self.cap)
        self.totals.push(total, self.cap)
        return rows, total

    def check(self):
        return self.totals.check("OverlapFirstStep")


class ListIndex:
    """Device index of the join / count_overlaps / overlap fast path -- the NCLS replacement of
    pyranges/methods/join.py:12 for all keys at once, built without a sort (csrc/lists.cu): per coordinate bin the list
    of subjects touching it, 32-byte bin records with the first three entries inline.  iv_join_pairs / iv_join_count
    (csrc/join.cu) answer a query from one record gather in the common case."""

    LONG_QUERY_BINS = 4  # queries longer than this many bins on average go to the rank index (one gather per bin)

    @staticmethod
    def supported(s_iv, group_ranges, q_mean_len=None):
        """can (and should) these subjects / queries of this mean length use the list index?"""
        g = s_iv.groups(group_ranges, gap=1, widen=s_iv.has_zero_length)
        shift = cabi.lib().iv_lists_shift(s_iv.n, g.total_span, g.sumlen)
        if shift < 0:
            return False
        if s_iv.has_zero_length:
            return True  # the only index that handles zero-length subjects
        return q_mean_len is None or q_mean_len <= ListIndex.LONG_QUERY_BINS * (1 << shift)

    def __init__(self, iv, seg_ranges=None, reach=0):
        """reach: the longest query that will be asked (a hint, see iv_lists_build): such queries never look at a second
        bin.  Queries.max_len() gives it for resident query columns."""
        self.iv = iv
        self.reach = int(reach)
        self.groups = iv.groups(seg_ranges, gap=1, widen=iv.has_zero_length)
        L = cabi.lib()
        nbytes = L.iv_lists_workspace_bytes(iv.n, self.groups.total_span, self.groups.sumlen)
        if nbytes == 0:
            raise cabi.CabiError("ListIndex: span / subject lengths beyond the list index (use SubjectIndex)")
        self.ws = _bytes(nbytes, iv.device)
        self.c = cabi.IvLists()
        self.build()

    def build(self):
        iv = self.iv
        cabi.check(cabi.lib().iv_lists_build(_ptr(iv.start), _ptr(iv.end), iv.n, C.byref(self.groups.c), self.reach,
                                             _ptr(self.ws), self.ws.numel(), C.byref(self.c), _stream()), "iv_lists_build")
        return self

    def join_pairs(self, q, out_q, out_s, capacity, s_label=None, mode=IV_MODE_ALL):
        """launch the fused join into caller buffers of `capacity` entries; -> total (device int64[1]); pairs beyond
        the capacity are dropped: compare the total and repeat with larger buffers.  mode IV_MODE_ANY: one entry per
        query with hits (out_q only)."""
        L = cabi.lib()
        dev = self.iv.device
        ws = _bytes(L.iv_join_workspace_bytes(q.iv.n), dev)
        total = torch.empty(1, dtype=torch.int64, device=dev)
        if mode == IV_MODE_ALL and q.iv.position_ordered:
            mode |= cabi.IV_JOIN_HINT_ORDERED
        cabi.check(L.iv_join_pairs(C.byref(self.c), C.byref(q.c), mode, _ptr(out_q), _ptr(out_s), _ptr(s_label),
                                   int(capacity), _ptr(ws), ws.numel(), _ptr(total), _stream()), "iv_join_pairs")
        return total

    def pairs(self, q, want_q=True, want_s=True, s_label=None, capacity=None):
        """all (query row | label, subject row | label) pairs in query order (NCLS.all_overlaps_both) -> (out_q, out_s)"""
        dev = self.iv.device
        cap = int(capacity) if capacity else max(2 * q.iv.n, 1 << 16)
        while True:
            oq = torch.empty(cap, dtype=torch.int64, device=dev) if want_q else None
            os_ = torch.empty(cap, dtype=torch.int64, device=dev) if want_s else None
            p = int(self.join_pairs(q, oq, os_, cap, s_label).item())
            if p <= cap:
                return (oq[:p] if want_q else None), (os_[:p] if want_s else None)
            cap = p

    def count(self, q, mode=IV_MODE_ALL):
        """overlaps per query (IV_MODE_ALL) or 0/1 (IV_MODE_ANY), int64[nq]"""
        counts = torch.empty(q.iv.n, dtype=torch.int64, device=self.iv.device)
        cabi.check(cabi.lib().iv_join_count(C.byref(self.c), C.byref(q.c), mode, _ptr(counts), _stream()), "iv_join_count")
        return counts

    def segment_pair_counts(self, q, out_q, total, capacity):
        """pairs per query segment (key) of a pair list written WITHOUT query labels (query rows ascending)"""
        out = torch.empty(max(q.iv.n_seg, 1), dtype=torch.int64, device=self.iv.device)
        cabi.check(cabi.lib().iv_sorted_pair_segments(_ptr(out_q), _ptr(total), int(capacity), _ptr(q.iv.seg_off_dev),
                                                      q.iv.n_seg, _ptr(out), _stream()), "iv_sorted_pair_segments")
        return out[:q.iv.n_seg]
