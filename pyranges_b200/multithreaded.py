"""Batched GPU replacement of the reference's per-key dispatcher.

Same two entry points, same signatures and the same result contract as
pyranges/multithreaded.py:182-303 (``pyrange_apply``) and :306-372 (``pyrange_apply_single``):
a dict {key -> DataFrame} holding only the non-empty results, each with a RangeIndex.  Where the
reference loops over the (chromosome[, strand]) keys and calls ``function(df, odf, **kwargs)`` once per
key (building an NCLS per key), this module recognises the hot-path callables by name

    _write_both, _number_overlapping, _overlap, _intersection, _subtraction, _count_overlaps, _coverage,
    _nearest                                                                (binary)
    _cluster, _cluster_by, _merge, _merge_by, _split, coverage (pyrle.methods.coverage)   (unary)

concatenates the Start/End columns of all keys into device-resident int64 columns and runs ONE batched
kernel sequence (pyranges_b200.engine) for every key; the pandas glue around the index arrays (row
gathers, suffix joins, -1 fillers) follows pyranges/methods/*.py.  Any other callable raises
NotImplementedError -- there is no CPU fallback.
"""
import numpy as np
import pandas as pd

from . import engine as E
from ._natsort import natsorted

_OTHER_STRAND = {"+": "-", "-": "+"}
_SAME_STRAND = {"+": "+", "-": "-"}


def _name(function):
    return function if isinstance(function, str) else getattr(function, "__name__", repr(function))


# ---- plumbing shared with the reference --------------------------------------------------------------
def merge_dfs(df1, df2):
    """pyranges/multithreaded.py:64-74"""
    if not df1.empty and not df2.empty:
        return pd.concat([df1, df2], sort=False).reset_index(drop=True)
    elif df1.empty and df2.empty:
        return None
    elif df1.empty:
        return df2
    return df1


def process_results(results, keys):
    """pyranges/multithreaded.py:77-103: drop None/empty, RangeIndex.  (The reference also deep-copies each frame
    because its per-key callables may return views of the inputs; every frame built in this module is new.)"""
    results_dict = {k: r for k, r in zip(keys, results) if r is not None}
    try:
        first_item = next(iter(results_dict.values()))
    except StopIteration:
        return results_dict
    if not isinstance(first_item, pd.DataFrame):
        return results_dict
    out = {}
    for k, v in results_dict.items():
        if v is None or v.empty:
            continue
        if not (isinstance(v.index, pd.RangeIndex) and v.index.start == 0 and v.index.step == 1):
            v.index = pd.RangeIndex(len(v))
        out[k] = v
    return out


def make_sparse(df):
    cols = "Chromosome Start End Strand".split() if "Strand" in df else "Chromosome Start End".split()
    return df[cols]


def _dummy():
    return pd.DataFrame(columns="Chromosome Start End".split())


def _is_stranded(gr):
    keys = list(gr.dfs.keys())
    return True if not keys else isinstance(keys[0], tuple)


# ---- device ingest --------------------------------------------------------------------------------------
class _Frames:
    """A list of (key, DataFrame) in natsorted key order plus their concatenated device columns."""

    def __init__(self, items, device, upload=True):
        self.keys = [k for k, _ in items]
        self.dfs = [d for _, d in items]
        lens = np.asarray([len(d) for d in self.dfs], dtype=np.int64)
        self.seg_off = np.zeros(len(lens) + 1, np.int64)
        np.cumsum(lens, out=self.seg_off[1:])
        if not upload:  # the caller lays the rows out differently (cluster / merge by=)
            self.starts = self.ends = self.iv = None
            return
        if len(self.dfs):
            s = np.concatenate([np.asarray(d["Start"].values, dtype=np.int64) for d in self.dfs])
            e = np.concatenate([np.asarray(d["End"].values, dtype=np.int64) for d in self.dfs])
        else:
            s = e = np.zeros(0, np.int64)
        self.starts, self.ends = s, e
        self.iv = E.DeviceIntervals.from_numpy(s, e, self.seg_off, device)

    def with_frames(self, items):
        """the same device columns under other (key, DataFrame) items holding the same coordinates"""
        f = _Frames.__new__(_Frames)
        f.__dict__.update(self.__dict__)
        f.keys = [k for k, _ in items]
        f.dfs = [d for _, d in items]
        f.__dict__.pop("_merged", None)  # concatenated partner frames belong to the frame objects
        return f


# Device copies of the coordinate columns persist across calls (north-star subsystem 1: "device-resident int64
# Start/End/index columns grouped by (Chromosome, Strand)"): a PyRanges keeps its _Frames on the object, and a small
# table keyed by the identity of the Start / End buffers serves the derived objects the API creates per call (sparse
# copies, unstranded views) -- pandas >= 3 column subsets share their buffers.  An entry holds references to those
# buffers, so an address can never come back with other contents while it is cached; replacing a column (PyRanges
# attribute assignment, any pandas operation: copy-on-write) changes the buffer and misses.  Writing INTO a buffer
# behind pandas' back (ndarray views) is not detected.
_FRAME_CACHE = {}
_FRAME_CACHE_MAX = 8


def _col_array(df, col):
    """the array behind a column without boxing a Series (25x cheaper than df[col].values; pandas-internal accessor
    with the public one as fallback)"""
    try:
        return df._get_column_array(df.columns.get_loc(col))
    except Exception:
        return df[col].values


_COLUMN_INDEX = {}  # column names -> pd.Index (the same few tuples come back for every key of every call)


def _new_frame(data, n):
    """DataFrame over ready column arrays with a RangeIndex: no re-validation of the columns, no consolidation of
    equal-typed columns into one block (a copy of every column), the Index of the column names built once per set of
    names -- pandas-internal constructors with the public one as fallback."""
    names = tuple(data.keys())
    cols = _COLUMN_INDEX.get(names)
    if cols is None:
        if len(_COLUMN_INDEX) > 256:
            _COLUMN_INDEX.clear()
        cols = _COLUMN_INDEX[names] = pd.Index(list(names))
    arrays = list(data.values())
    try:
        from pandas.core.internals.construction import arrays_to_mgr

        mgr = arrays_to_mgr(arrays, cols, pd.RangeIndex(n), verify_integrity=False, consolidate=False)
        return pd.DataFrame._from_mgr(mgr, axes=mgr.axes)
    except Exception:
        pass
    try:
        return pd.DataFrame._from_arrays(arrays, cols, pd.RangeIndex(n), verify_integrity=False)
    except Exception:
        return pd.DataFrame(data, copy=False)


def _col_ptr(df, col):
    v = _col_array(df, col)
    if not isinstance(v, np.ndarray) or v.dtype != np.int64:
        return None, None
    return v.__array_interface__["data"][0], v


def _frames_of(gr, device):
    """the (cached) device ingest of a PyRanges-like object"""
    items = natsorted(gr.dfs.items())
    sig, held = [str(device), id(E)], []
    for k, d in items:
        ps, vs = _col_ptr(d, "Start")
        pe, ve = _col_ptr(d, "End")
        if ps is None or pe is None:
            return _Frames(items, device)
        sig.append((k, len(d), ps, pe, vs.strides[0], ve.strides[0]))
        held.append((vs, ve))
    sig = tuple(sig)
    hit = _FRAME_CACHE.get(sig)
    if hit is not None:
        f = hit[0]
        if [id(d) for d in f.dfs] != [id(d) for _, d in items]:
            # same coordinates, other frame objects (payload columns may differ): share the device columns only
            f = f.with_frames(items)
        return f
    f = _Frames(items, device)
    if len(_FRAME_CACHE) >= _FRAME_CACHE_MAX:
        _FRAME_CACHE.pop(next(iter(_FRAME_CACHE)))
    _FRAME_CACHE[sig] = (f, held)
    return f


def seed_frame_cache(gr, device, frames):
    """A device ingest made elsewhere (pyranges_b200.readers: columns parsed on the device) becomes the cached one of
    `gr`; False when the frames do not qualify (non-int64 coordinate columns)."""
    items = natsorted(gr.dfs.items())
    sig, held = [str(device), id(E)], []
    for k, d in items:
        ps, vs = _col_ptr(d, "Start")
        pe, ve = _col_ptr(d, "End")
        if ps is None or pe is None:
            return False
        sig.append((k, len(d), ps, pe, vs.strides[0], ve.strides[0]))
        held.append((vs, ve))
    if len(_FRAME_CACHE) >= _FRAME_CACHE_MAX:
        _FRAME_CACHE.pop(next(iter(_FRAME_CACHE)))
    _FRAME_CACHE[tuple(sig)] = (frames, held)
    return True


def clear_device_cache():
    """drop every cached device copy of coordinate columns"""
    _FRAME_CACHE.clear()


def _device_of(kwargs):
    import torch

    dev = kwargs.get("device")
    if dev is None:
        if not torch.cuda.is_available():
            raise E.cabi.CabiError("pyranges_b200 needs a CUDA device (B200); there is no CPU fallback")
        dev = torch.device("cuda", torch.cuda.current_device())
    return dev


class _Plan:
    """Partner selection of pyrange_apply (pyranges/multithreaded.py:221-294) for all keys at once."""

    def __init__(self, self_gr, other_gr, strandedness, device, q_frames=None):
        assert strandedness in ["same", "opposite", False, None]
        self_stranded, other_stranded = _is_stranded(self_gr), _is_stranded(other_gr)
        if strandedness:
            assert self_stranded and other_stranded, \
                "Can only do stranded operations when both PyRanges contain strand info"
        self.q = q_frames if q_frames is not None else _frames_of(self_gr, device)
        self.s = _frames_of(other_gr, device)
        okeys = self.s.keys
        if strandedness or not other_stranded:
            # every key of other is one partner frame
            self.group_ranges = [(i, i + 1) for i in range(len(okeys))]
            gkey = {k: i for i, k in enumerate(okeys)}
            self._frames = list(self.s.dfs)
        else:
            # strand ignored, other stranded: partner = '+' and '-' frames of the chromosome, concatenated
            self.group_ranges, gkey, self._frames = [], {}, []
            i = 0
            while i < len(okeys):
                j = i
                while j < len(okeys) and okeys[j][0] == okeys[i][0]:
                    j += 1
                gkey[okeys[i][0]] = len(self.group_ranges)
                self.group_ranges.append((i, j))
                self._frames.append(None)  # merged lazily
                i = j
        strand_dict = _OTHER_STRAND if strandedness == "opposite" else _SAME_STRAND
        sg = []
        for k in self.q.keys:
            if strandedness:
                pk = (k[0], strand_dict[k[1]])
            elif other_stranded:
                pk = k[0] if self_stranded else k
            else:
                pk = k[0] if self_stranded else k
            sg.append(gkey.get(pk, -1))
        self.seg_group = np.asarray(sg, dtype=np.int32)
        self.group_row_off = np.asarray([self.s.seg_off[r[0]] for r in self.group_ranges] + [self.s.iv.n], dtype=np.int64)

    def partner(self, k):
        """partner DataFrame of self key number k (0-row dummy when other has none)."""
        g = self.seg_group[k]
        if g < 0:
            return _dummy()
        f = self._frames[g]
        if f is None:
            a, b = self.group_ranges[g]
            # the concatenated partner frame of a chromosome lives with the (cached) device ingest of `other`: building
            # it costs a pandas concat per chromosome and call otherwise.  Every key of the group is concatenated, in
            # key order -- the row layout of the device index.
            memo = self.s.__dict__.setdefault("_merged", {})
            f = memo.get((a, b))
            if f is None:
                dfs = self.s.dfs[a:b]
                f = dfs[0] if len(dfs) == 1 else pd.concat(dfs, sort=False).reset_index(drop=True)
                memo[(a, b)] = f
            self._frames[g] = f
        return f

    def index(self, ends_perm=False):
        return E.SubjectIndex(self.s.iv, self.group_ranges if self.group_ranges else None, ends_perm=ends_perm)

    def queries(self, seg_mode=None, slack=0):
        return E.Queries(self.q.iv, self.seg_group, seg_mode=seg_mode, slack=slack)


def _sparse(kwargs, df, odf):
    sparse = kwargs.get("sparse")
    if not sparse:
        return df, odf
    if sparse.get("self"):
        df = make_sparse(df)
    if sparse.get("other") and odf is not None and len(odf.columns):
        odf = make_sparse(odf)
    return df, odf


# ---- binary operations --------------------------------------------------------------------------------------
def null_types(h):
    """pyranges/methods/join.py:62-92: one-row frame of "missing" values (-1 / "-1")."""
    h2 = h.copy()
    for n, d in zip(h, h.dtypes):
        if n in ["Chromosome", "Strand"]:
            continue
        d = str(d)
        if "bool" in d:
            h2[n] = h2[n].astype("category")
            d = "category"
        if "int" in d or "float" in d:
            null = -1
        elif "string" in d or d in ("object", "str"):
            null = "-1"
        elif d == "category":
            h2[n] = h2[n].cat.add_categories("-1")
            null = "-1"
        else:
            raise Exception("Unknown dtype {} in a column {}".format(d, n))
        h2.loc[:, n] = null
    return h2


def _take_with_null(df, pos):
    """df.reindex(labels) of the reference with positions: -1 selects the null row."""
    pos = np.asarray(pos, dtype=np.int64)
    if (pos < 0).any():
        nh = null_types(df.head(1))
        df = pd.concat([df, nh])
        pos = np.where(pos < 0, len(df) - 1, pos)
    return df.take(pos)


def _assemble(scdf, ocdf, self_pos, other_pos, suffix, core_self=None, core_other=None, drop_other=("Chromosome",),
              extra=None, extra_is_other=False):
    """The inner-join tail of _write_both (pyranges/methods/join.py:95-154) column by column: rows `self_pos` of scdf
    next to rows `other_pos` of ocdf, names of ocdf that collide get `suffix`.  `core_self` / `core_other` hold the
    columns of scdf / ocdf that were already gathered on the device (engine.materialize_pairs), keyed by the SOURCE
    column name; `extra` is appended last (as columns of ocdf, i.e. suffixed on collision, if `extra_is_other`).  One take per remaining column instead of two DataFrame.take + drop +
    DataFrame.join."""
    core_self, core_other = core_self or {}, core_other or {}
    data = {}
    for c in scdf.columns:
        col = _col_array(scdf, c)
        data[c] = core_self[c].astype(col.dtype, copy=False) if c in core_self else col.take(self_pos)
    for c in ocdf.columns:
        if c in drop_other:
            continue
        name = c + suffix if c in data else c
        col = _col_array(ocdf, c)
        data[name] = core_other[c].astype(col.dtype, copy=False) if c in core_other else col.take(other_pos)
    for n, v in (extra or {}).items():
        data[n + suffix if (extra_is_other and n in data) else n] = v
    return _new_frame(data, len(self_pos))


def _join_frames(scdf, ocdf, self_pos, other_pos, counts_self, kwargs):
    """_both_indexes tail + _both_dfs + _write_both (pyranges/methods/join.py:22-154) given the inner pairs
    as row positions into scdf / ocdf."""
    how = kwargs.get("how")
    suffix = kwargs.get("suffix", "_b") if not kwargs.get("new_pos") else kwargs.get("suffixes", "_a _b".split())[1]
    if how in ["outer", "left", "right"]:
        slab, olab = scdf.index.values, ocdf.index.values
        ms = np.flatnonzero(counts_self == 0)
        ms = ms[np.argsort(slab[ms], kind="stable")]  # Index.difference returns sorted labels
        hit_o = np.zeros(len(ocdf), dtype=bool)
        hit_o[other_pos] = True
        mo = np.flatnonzero(~hit_o)
        mo = mo[np.argsort(olab[mo], kind="stable")]
        fs, fo = np.full(len(mo), -1, np.int64), np.full(len(ms), -1, np.int64)
        if how == "outer":
            self_pos, other_pos = np.concatenate([self_pos, ms, fs]), np.concatenate([other_pos, fo, mo])
        elif how == "left":
            self_pos, other_pos = np.concatenate([self_pos, ms]), np.concatenate([other_pos, fo])
        else:
            self_pos, other_pos = np.concatenate([self_pos, fs]), np.concatenate([other_pos, mo])
        if kwargs.get("preserve_order"):
            sl = np.where(self_pos >= 0, slab[np.maximum(self_pos, 0)], -1) if len(slab) else self_pos
            ol = np.where(other_pos >= 0, olab[np.maximum(other_pos, 0)], -1) if len(olab) else other_pos
            if how == "outer":
                o = np.lexsort([ol, ol < 0, sl, sl < 0])
            elif how == "left":
                o = np.argsort(sl, kind="stable")
            else:
                o = np.argsort(ol, kind="stable")
            self_pos, other_pos = self_pos[o], other_pos[o]
        s2, o2 = _take_with_null(scdf, self_pos), _take_with_null(ocdf, other_pos)
        if "Strand" in s2 and "Strand" in o2:
            if how == "left":
                x = other_pos < 0
                if x.any():
                    o2 = o2.copy()
                    o2["Strand"] = np.where(x, s2["Strand"].astype(object).values, o2["Strand"].astype(object).values)
            elif how == "right":
                x = self_pos < 0
                if x.any():
                    s2 = s2.copy()
                    s2["Strand"] = np.where(x, o2["Strand"].astype(object).values, s2["Strand"].astype(object).values)
    else:
        s2, o2 = scdf.take(self_pos), ocdf.take(other_pos)
    nix = pd.RangeIndex(len(s2))
    s2.index = nix
    o2.index = nix
    o2 = o2.drop("Chromosome", axis=1)
    df = s2.join(o2, rsuffix=suffix)
    if kwargs.get("report_overlap"):
        df["Overlap"] = df[["End", "End" + suffix]].min(axis=1) - df[["Start", "Start" + suffix]].max(axis=1)
    return df


def _b_write_both(plan, kwargs):
    """join: pyranges/methods/join.py:125-154 for all keys."""
    how = kwargs.get("how")
    assert (how in "containment first last outer right left".split() + [False, None]) or isinstance(how, int)
    mode = {"containment": E.IV_MODE_CONTAIN, "first": E.IV_MODE_ANY, "last": E.IV_MODE_LAST}.get(how, E.IV_MODE_ALL)
    inner = how not in ["outer", "left", "right"]
    if mode == E.IV_MODE_ALL:  # counts: rows without partner (outer joins)
        counts, out_q, out_s = E.overlap_pairs(plan.s.iv, plan.group_ranges or None, plan.queries(), want_counts=not inner)
    else:
        counts, out_q, out_s = plan.index().pairs(plan.queries(), mode, want_counts=not inner)
    core = None
    if inner:
        # the coordinate columns of the joined frame are gathered on the device, where both
        # tables already live; the host only gathers the payload columns (by position, one take per column)
        want = ("q_start", "q_end", "s_start", "s_end") + (("overlap",) if kwargs.get("report_overlap") else ())
        core = {n: t.cpu().numpy() for n, t in E.materialize_pairs(plan.q.iv, plan.s.iv, out_q, out_s, want=want).items()}
    counts = counts.cpu().numpy() if counts is not None else None
    out_q, out_s = out_q.cpu().numpy(), out_s.cpu().numpy()
    first = np.searchsorted(out_q, plan.q.seg_off)  # pairs are emitted in query-row order
    suffix = kwargs.get("suffix", "_b") if not kwargs.get("new_pos") else kwargs.get("suffixes", "_a _b".split())[1]
    results = []
    for k, scdf in enumerate(plan.q.dfs):
        ocdf = plan.partner(k)
        scdf, ocdf = _sparse(kwargs, scdf, ocdf)
        a, b = first[k], first[k + 1]
        if inner:
            if scdf.empty or ocdf.empty or a == b:
                results.append(None)
                continue
            cq = {"Start": core["q_start"][a:b], "End": core["q_end"][a:b]}
            co = {"Start": core["s_start"][a:b], "End": core["s_end"][a:b]}
            extra = {"Overlap": core["overlap"][a:b]} if "overlap" in core else None
            results.append(_assemble(scdf, ocdf, out_q[a:b] - plan.q.seg_off[k],
                                     out_s[a:b] - plan.group_row_off[plan.seg_group[k]], suffix, cq, co, extra=extra))
            continue
        if scdf.empty or ocdf.empty:
            if how in ["left", "outer"] and ocdf.empty:
                ocdf = null_types(kwargs["example_header_other"])
                cs = np.zeros(len(scdf), np.int64)
                results.append(_join_frames(scdf, ocdf, np.zeros(0, np.int64), np.zeros(0, np.int64), cs, kwargs))
            else:
                results.append(None)
            continue
        g = plan.seg_group[k]
        self_pos = out_q[a:b] - plan.q.seg_off[k]
        other_pos = out_s[a:b] - plan.group_row_off[g]
        cs = counts[plan.q.seg_off[k]:plan.q.seg_off[k + 1]]
        results.append(_join_frames(scdf, ocdf, self_pos, other_pos, cs, kwargs))
    return results


def _b_number_overlapping(plan, kwargs):
    """count_overlaps: pyranges/methods/coverage.py:6-45 for all keys (no pair materialisation)."""
    keep = kwargs.get("keep_nonoverlapping", True)
    col = kwargs.get("overlap_col", True)
    counts = E.overlap_counts(plan.s.iv, plan.group_ranges or None, plan.queries(), E.IV_MODE_ALL).cpu().numpy()
    results = []
    for k, scdf in enumerate(plan.q.dfs):
        ocdf = plan.partner(k)
        if scdf.empty or (ocdf.empty and not keep):
            results.append(None)
            continue
        df = scdf.copy(deep=False)  # new frame object, columns shared: only a column is added
        df.insert(df.shape[1], col, counts[plan.q.seg_off[k]:plan.q.seg_off[k + 1]])
        results.append(df if keep else df[df[col] != 0])
    return results


def _b_overlap(plan, kwargs):
    """overlap: pyranges/methods/intersection.py:58-83 for all keys; only counts are needed."""
    how = kwargs["how"]
    assert how in "containment first".split() + [False, None]
    mode = E.IV_MODE_ALL if not how else (E.IV_MODE_CONTAIN if how == "containment" else E.IV_MODE_ANY)
    counts = E.overlap_counts(plan.s.iv, plan.group_ranges or None, plan.queries(), mode).cpu().numpy()
    results = []
    for k, scdf in enumerate(plan.q.dfs):
        ocdf = plan.partner(k)
        # (only the self frame is used below: the sparse copy of the partner frame, intersection.py:63, is skipped)
        scdf, _ = _sparse(kwargs, scdf, None)
        if scdf.empty or ocdf.empty:
            results.append(None)
            continue
        c = counts[plan.q.seg_off[k]:plan.q.seg_off[k + 1]]
        pos = np.repeat(np.arange(len(c)), c)
        if kwargs.get("return_indexes", False):
            results.append(scdf.index.values[pos])
        else:
            results.append(scdf.take(pos))
    return results


def _b_intersection(plan, kwargs):
    """intersect / set_intersect: pyranges/methods/intersection.py:6-55 for all keys -- the self rows of every
    overlapping pair, clipped to the overlap."""
    how = kwargs["how"]
    assert how in "containment first last".split() + [False, None]
    mode = {"containment": E.IV_MODE_CONTAIN, "first": E.IV_MODE_ANY, "last": E.IV_MODE_LAST}.get(how, E.IV_MODE_ALL)
    if mode == E.IV_MODE_ALL:
        _, out_q, out_s = E.overlap_pairs(plan.s.iv, plan.group_ranges or None, plan.queries())
    else:
        _, out_q, out_s = plan.index().pairs(plan.queries(), mode, want_counts=False)
    # the clip max(Start, Start_b) / min(End, End_b) happens where the columns live (engine.materialize_pairs)
    core = E.materialize_pairs(plan.q.iv, plan.s.iv, out_q, out_s, clip=True, want=("q_start", "q_end"))
    cs, ce = core["q_start"].cpu().numpy(), core["q_end"].cpu().numpy()
    out_q = out_q.cpu().numpy()
    first = np.searchsorted(out_q, plan.q.seg_off)
    results = []
    for k, scdf in enumerate(plan.q.dfs):
        ocdf = plan.partner(k)
        scdf, _ = _sparse(kwargs, scdf, None)  # (of the partner frame only emptiness and the Start dtype are used)
        a, b = first[k], first[k + 1]
        if scdf.empty or ocdf.empty or a == b:
            results.append(None)
            continue
        in_dtype = ocdf.Start.dtype
        sp = out_q[a:b] - plan.q.seg_off[k]
        data = {}
        for c in scdf.columns:
            if c == "Start":
                data[c] = cs[a:b].astype(in_dtype, copy=False)
            elif c == "End":
                data[c] = ce[a:b].astype(in_dtype, copy=False)
            else:
                data[c] = scdf[c].values.take(sp)
        results.append(pd.DataFrame(data, copy=False))
    return results


def _b_subtraction(plan, kwargs):
    """subtract: pyranges/methods/subtraction.py:47-89 for all keys.  `other` must already be merged (disjoint), as
    PyRanges.subtract does (pyranges/pyranges_main.py:4868).  Rows without overlap come back unchanged, overlapped
    rows once per surviving piece with the piece's Start / End, completely covered rows not at all."""
    oq, os_, oe = plan.index().subtract(plan.queries())
    oq, os_, oe = oq.cpu().numpy(), os_.cpu().numpy(), oe.cpu().numpy()
    first = np.searchsorted(oq, plan.q.seg_off)
    results = []
    for k, scdf in enumerate(plan.q.dfs):
        a, b = first[k], first[k + 1]
        if scdf.empty or a == b:
            results.append(None)
            continue
        df = scdf.take(oq[a:b] - plan.q.seg_off[k])
        df["Start"] = os_[a:b].astype(scdf.Start.dtype)
        df["End"] = oe[a:b].astype(scdf.End.dtype)
        results.append(df)
    return results


def _b_count_overlaps(plan, kwargs):
    """pr.count_overlaps helper: pyranges/methods/intersection.py:86-99 for all keys -- overlaps per row according to
    `how` as an int32 column kwargs["name"].  The reference inserts into the caller's frame in place; here the key's
    frame is copied (Appendix B quirk 6 consciously not kept)."""
    how = kwargs.get("how")
    mode = E.IV_MODE_ALL if not how else (E.IV_MODE_CONTAIN if how == "containment" else E.IV_MODE_ANY)
    counts = E.overlap_counts(plan.s.iv, plan.group_ranges or None, plan.queries(), mode).cpu().numpy()
    results = []
    for k, scdf in enumerate(plan.q.dfs):
        if scdf.empty:
            results.append(None)
            continue
        c = counts[plan.q.seg_off[k]:plan.q.seg_off[k + 1]] if not plan.partner(k).empty else np.zeros(len(scdf), np.int64)
        df = scdf.copy(deep=False)  # new frame object, columns shared: only a column is added
        df.insert(df.shape[1], kwargs["name"], c.astype(np.int32))
        results.append(df)
    return results


def _b_coverage(plan, kwargs):
    """coverage fraction: pyranges/methods/coverage.py:48-74 for all keys."""
    col = kwargs["fraction_col"]
    _, frac = plan.index().coverage_bp(plan.queries())
    frac = frac.cpu().numpy()
    results = []
    for k, scdf in enumerate(plan.q.dfs):
        if scdf.empty:
            results.append(None)
            continue
        df = scdf.copy(deep=False)  # new frame object, columns shared: only a column is added
        f = frac[plan.q.seg_off[k]:plan.q.seg_off[k + 1]] if not plan.partner(k).empty else 0.0
        df.insert(df.shape[1], col, f)
        results.append(df)
    return results


def _distance_name(ocdf, suffix):
    """pyranges/methods/nearest.py:9-26"""
    if "Distance" not in ocdf:
        return "Distance"
    if "Distance" + suffix not in ocdf:
        return "Distance" + suffix
    i = 1
    while "Distance" + str(i) in ocdf:
        i += 1
    return "Distance" + str(i)


def _b_nearest(plan, kwargs):
    """nearest: pyranges/methods/nearest.py:77-132 for all keys."""
    overlap, how, suffix = kwargs["overlap"], kwargs["how"], kwargs["suffix"]
    modes = []
    for k, key in enumerate(plan.q.keys):
        h = how
        if how in ("upstream", "downstream"):
            strand = key[1] if isinstance(key, tuple) else plan.q.dfs[k].Strand.iloc[0]
            h = ({"+": "previous", "-": "next"} if how == "upstream" else {"+": "next", "-": "previous"})[strand]
        modes.append({"previous": E.IV_NEAREST_PREVIOUS, "next": E.IV_NEAREST_NEXT}.get(h, E.IV_NEAREST_BOTH))
    row, dist = plan.index(ends_perm=True).nearest(plan.queries(seg_mode=np.asarray(modes, np.int32)), overlap=overlap)
    row, dist = row.cpu().numpy(), dist.cpu().numpy()
    # the reference sorts the rows without overlap by (Start, End) (sort_one_by_one, nearest.py:100): the (key, Start,
    # End, row) order of all keys comes from the device sort of the cluster kernels
    by_start = E.cluster(plan.q.iv, None, 0)[0].cpu().numpy() if plan.q.iv.n else np.zeros(0, np.int64)
    results = []
    for k, scdf in enumerate(plan.q.dfs):
        ocdf = plan.partner(k)
        if scdf.empty or ocdf.empty:
            results.append(None)
            continue
        a, b = plan.q.seg_off[k], plan.q.seg_off[k + 1]
        sl = slice(a, b)
        r = row[sl] - plan.group_row_off[plan.seg_group[k]]
        d = dist[sl]
        hit = np.flatnonzero(d == 0)  # overlapping rows first, in self order (nearest.py:29-53)
        order = by_start[sl] - a
        rest = order[d[order] > 0]
        sel = np.concatenate([hit, rest])
        if not len(sel):
            results.append(None)
            continue
        results.append(_assemble(scdf, ocdf, sel, r[sel], suffix, extra={_distance_name(ocdf, suffix): d[sel]},
                                 extra_is_other=True))
    return results


def _b_k_nearest(plan, kwargs):
    """The per-key candidate frames of k_nearest (pyranges/methods/k_nearest.py:51-168) for all keys: rows of self
    (they carry the `__k__` column PyRanges.k_nearest adds) next to their non-overlapping neighbours in other, previous
    ones before next ones at equal distance, sorted by (self row, Distance); Distance = gap + 1, still unsigned."""
    how, suffix, ties = kwargs.get("how"), kwargs.get("suffix", "_b"), kwargs.get("ties")
    stranded = kwargs.get("stranded", False)
    modes = []
    for k, key in enumerate(plan.q.keys):
        h = None
        if how in ("upstream", "downstream"):
            if stranded:
                strand = key[1] if isinstance(key, tuple) else plan.q.dfs[k].Strand.iloc[0]
                h = ({"+": "previous", "-": "next"} if how == "upstream" else {"+": "next", "-": "previous"})[strand]
            else:
                h = "previous" if how == "upstream" else "next"
        modes.append({"previous": E.IV_NEAREST_PREVIOUS, "next": E.IV_NEAREST_NEXT}.get(h, E.IV_NEAREST_BOTH))
    if not plan.q.iv.n:
        return [None] * len(plan.q.dfs)
    ks = np.concatenate([np.asarray(df["__k__"].values, dtype=np.int64) for df in plan.q.dfs])
    kdev = E.to_device_i64(ks, getattr(plan.q.iv, "device", None))
    q, r, d = plan.index(ends_perm=True).k_nearest(plan.queries(seg_mode=np.asarray(modes, np.int32)), kdev, ties)
    q, r, d = q.cpu().numpy(), r.cpu().numpy(), d.cpu().numpy()
    cut = np.searchsorted(q, np.asarray(plan.q.seg_off))  # candidates come in query order
    results = []
    for k, scdf in enumerate(plan.q.dfs):
        ocdf = plan.partner(k)
        a, b = cut[k], cut[k + 1]
        if scdf.empty or ocdf.empty or a == b:
            results.append(None)
            continue
        lq = q[a:b] - plan.q.seg_off[k]
        ls = r[a:b] - plan.group_row_off[plan.seg_group[k]]
        dd = d[a:b]
        o = np.lexsort((dd, lq))  # stable: previous before next at equal distance (k_nearest.py:79)
        results.append(_assemble(scdf, ocdf, lq[o], ls[o], suffix, extra={"Distance": dd[o]}))
    return results


_BINARY = {
    "_k_nearest": _b_k_nearest,
    "_write_both": _b_write_both,
    "_number_overlapping": _b_number_overlapping,
    "_overlap": _b_overlap,
    "_intersection": _b_intersection,
    "_subtraction": _b_subtraction,
    "_count_overlaps": _b_count_overlaps,
    "_coverage": _b_coverage,
    "_nearest": _b_nearest,
}


def pyrange_apply(function, self, other, **kwargs):
    """Drop-in for pyranges.multithreaded.pyrange_apply (pyranges/multithreaded.py:182-303).

    `function` is one of the reference's per-key callables (or its name); `self` / `other` are
    PyRanges-like objects exposing ``.dfs``.  ``nb_cpu`` is accepted and ignored: all keys run in one
    batched launch sequence on the current CUDA device."""
    name = _name(function)
    if name == "_nearest" and getattr(function, "__module__", "").endswith("k_nearest"):
        name = "_k_nearest"  # pyranges.methods.k_nearest._nearest shares its name with pyranges.methods.nearest._nearest
    if name not in _BINARY:
        raise NotImplementedError(
            "pyranges_b200.pyrange_apply: %r is not on the B200 hot path (supported: %s); no CPU fallback"
            % (name, ", ".join(sorted(_BINARY))))
    plan = _Plan(self, other, kwargs["strandedness"], _device_of(kwargs))
    results = _BINARY[name](plan, kwargs)
    return process_results(results, plan.q.keys)


# ---- unary operations -----------------------------------------------------------------------------------
class _UnaryPlan:
    """Key iteration of pyrange_apply_single (pyranges/multithreaded.py:324-363)."""

    def __init__(self, self_gr, strand, device, upload=True):
        stranded = _is_stranded(self_gr)
        if strand:
            assert stranded, "Can only do stranded operation when PyRange contains strand info"
        self.f = _frames_of(self_gr, device) if upload else _Frames(natsorted(self_gr.dfs.items()), device, upload=False)
        keys = self.f.keys
        if strand or not stranded:
            self.group_ranges = [(i, i + 1) for i in range(len(keys))]
            self.out_keys = list(keys)
            self.frames = list(self.f.dfs)
            self.chrom = [k[0] if isinstance(k, tuple) else k for k in keys]
            self.strand = [k[1] if (strand and isinstance(k, tuple)) else None for k in keys]
        else:
            self.group_ranges, self.out_keys, self.frames, self.chrom, self.strand = [], [], [], [], []
            i = 0
            while i < len(keys):
                j = i
                while j < len(keys) and keys[j][0] == keys[i][0]:
                    j += 1
                self.group_ranges.append((i, j))
                self.out_keys.append(keys[i][0])
                dfs = self.f.dfs[i:j]
                self.frames.append(dfs[0] if len(dfs) == 1 else merge_dfs(dfs[0], dfs[1]))
                self.chrom.append(keys[i][0])
                self.strand.append(None)
                i = j
        self.row_off = np.asarray([self.f.seg_off[r[0]] for r in self.group_ranges] + [self.f.seg_off[-1]], dtype=np.int64)


def _u_cluster(plan, kwargs):
    """cluster: pyranges/methods/cluster.py:4-22 per key; ids are LOCAL (1-based per key) exactly like the
    reference's per-key result -- PyRanges.cluster adds the running offsets (pyranges_main.py:1210-1223)."""
    slack, count = kwargs.get("slack", 0), kwargs.get("count", False)
    row, cid, _ = E.cluster(plan.f.iv, plan.group_ranges or None, slack)
    row, cid = row.cpu().numpy(), cid.cpu().numpy()
    results = []
    for g, df in enumerate(plan.frames):
        a, b = plan.row_off[g], plan.row_off[g + 1]
        if a == b:
            results.append(None)
            continue
        if kwargs.get("sparse", {}).get("self"):
            df = make_sparse(df)
        cdf = df.take(row[a:b] - a)
        ids = cid[a:b] - cid[a] + 1
        cdf.insert(cdf.shape[1], "Cluster", ids)
        if count:
            cdf.insert(cdf.shape[1], "Count", np.bincount(ids)[ids])
        results.append(cdf)
    return results


def _u_merge(plan, kwargs):
    """merge: pyranges/methods/merge.py:5-38 per key."""
    slack = kwargs.get("slack", 0)
    st, en, cnt, grp = E.merge(plan.f.iv, plan.group_ranges or None, slack, want_count=bool(kwargs["count"]))
    st, en, grp = st.cpu().numpy(), en.cpu().numpy(), grp.cpu().numpy()
    cnt = cnt.cpu().numpy() if cnt is not None else None
    first = np.searchsorted(grp, np.arange(len(plan.frames) + 1))
    results = []
    for g in range(len(plan.frames)):
        a, b = first[g], first[g + 1]
        if a == b:
            results.append(None)
            continue
        nidx = pd.RangeIndex(b - a)
        d = {"Chromosome": pd.Series(plan.chrom[g], dtype="category", index=nidx), "Start": st[a:b], "End": en[a:b]}
        if plan.strand[g]:
            d["Strand"] = pd.Series(plan.strand[g], dtype="category", index=nidx)
        cdf = pd.DataFrame(d)
        if kwargs["count"]:
            cdf.insert(cdf.shape[1], kwargs["count_col"], cnt[a:b])
        results.append(cdf)
    return results


def _u_coverage(plan, kwargs):
    """to_rle: pyrle.methods.coverage (pyranges/methods/to_rle.py:18) per key; unit weights or value_col."""
    from .rle import Rle

    weights = None
    if kwargs.get("value_col"):
        import torch

        col = kwargs["value_col"]
        w = np.concatenate([np.asarray(d[col].values, dtype=np.float64) for d in plan.f.dfs]) if plan.f.dfs else np.zeros(0)
        weights = torch.from_numpy(np.ascontiguousarray(w))
        dev = getattr(plan.f.iv, "device", None)
        if dev is not None:
            weights = weights.to(dev)
    run_off, runs, vals = E.coverage_rle(plan.f.iv, plan.group_ranges or None, weights)
    runs, vals = runs.cpu().numpy(), vals.cpu().numpy()
    results = []
    for g in range(len(plan.frames)):
        if plan.row_off[g] == plan.row_off[g + 1]:
            results.append(None)
            continue
        a, b = run_off[g], run_off[g + 1]
        results.append(Rle(runs[a:b].copy(), vals[a:b].copy()))
    return results


class _ByPlan:
    """cluster(by=) / merge(by=) (pyranges/methods/cluster.py:25-58, merge.py:41-92): inside every key the rows are
    grouped by the `by` column(s); the reference sorts each frame by (by, Start) and lets sorted_nearest.cluster_by /
    merge_by restart at every change of `by`.  Here every (key, by-value) becomes its own row group of the SAME batched
    cluster / merge kernels: per frame a stable order by the by-codes, then one device upload of the permuted columns
    with one segment per (key, by-value)."""

    def __init__(self, self_gr, strand, by, device):
        base = _UnaryPlan(self_gr, strand, device, upload=False)
        self.out_keys, self.frames, self.chrom, self.strand = base.out_keys, base.frames, base.chrom, base.strand
        self.by = [by] if isinstance(by, str) else list(by)
        self.perm, self.first_rows = [], []
        starts, ends, seg_len, frame_sub = [], [], [], [0]
        for df in self.frames:
            # group numbers in sorted order of the by-values = the order of sort_values(by) (cluster.py:35-38)
            codes = df.groupby(self.by, sort=True, observed=True).ngroup().to_numpy()
            if (codes < 0).any():
                raise NotImplementedError("cluster/merge(by=...): missing values in the `by` columns are out of scope")
            perm = np.argsort(codes, kind="stable")
            cnt = np.bincount(codes)
            cnt = cnt[cnt > 0]
            self.perm.append(perm)
            self.first_rows.append(perm[np.concatenate([[0], np.cumsum(cnt)[:-1]])] if len(cnt) else perm[:0])
            starts.append(np.asarray(df["Start"].values, dtype=np.int64)[perm])
            ends.append(np.asarray(df["End"].values, dtype=np.int64)[perm])
            seg_len.append(cnt)
            frame_sub.append(frame_sub[-1] + len(cnt))
        self.frame_sub = np.asarray(frame_sub, dtype=np.int64)  # sub-groups of frame g: [frame_sub[g], frame_sub[g+1])
        seg_off = np.zeros(self.frame_sub[-1] + 1, np.int64)
        if len(seg_len):
            np.cumsum(np.concatenate(seg_len), out=seg_off[1:])
        s = np.concatenate(starts) if starts else np.zeros(0, np.int64)
        e = np.concatenate(ends) if ends else np.zeros(0, np.int64)
        self.iv = E.DeviceIntervals.from_numpy(s, e, seg_off, device)
        self.row_off = seg_off[self.frame_sub]


def _u_cluster_by(plan, kwargs):
    """cluster(by=): pyranges/methods/cluster.py:25-58 per key; rows come back in (by, Start) order, ids local."""
    slack, count = kwargs.get("slack", 0), kwargs.get("count", False)
    row, cid, _ = E.cluster(plan.iv, None, slack)
    row, cid = row.cpu().numpy(), cid.cpu().numpy()
    results = []
    for g, df in enumerate(plan.frames):
        a, b = plan.row_off[g], plan.row_off[g + 1]
        if a == b:
            results.append(None)
            continue
        cdf = df.take(plan.perm[g][row[a:b] - a])
        ids = cid[a:b] - cid[a] + 1
        cdf.insert(cdf.shape[1], "Cluster", ids)
        if count:
            cdf.insert(cdf.shape[1], "Count", np.bincount(ids)[ids])
        results.append(cdf)
    return results


def _u_merge_by(plan, kwargs):
    """merge(by=): pyranges/methods/merge.py:41-92 per key: Chromosome, Start, End, [Strand], the by column(s) of the
    cluster's group, [count]."""
    slack = kwargs.get("slack", 0)
    st, en, cnt, grp = E.merge(plan.iv, None, slack, want_count=bool(kwargs["count"]))
    st, en, grp = st.cpu().numpy(), en.cpu().numpy(), grp.cpu().numpy()
    cnt = cnt.cpu().numpy() if cnt is not None else None
    first = np.searchsorted(grp, plan.frame_sub)
    results = []
    for g, df in enumerate(plan.frames):
        a, b = first[g], first[g + 1]
        if a == b:
            results.append(None)
            continue
        nidx = pd.RangeIndex(b - a)
        d = {"Chromosome": pd.Series(plan.chrom[g], dtype="category", index=nidx), "Start": st[a:b], "End": en[a:b]}
        if plan.strand[g]:
            d["Strand"] = pd.Series(plan.strand[g], dtype="category", index=nidx)
        src = plan.first_rows[g][grp[a:b] - plan.frame_sub[g]]  # a row of the cluster's by-group
        for c in plan.by:
            d[c] = df[c].values.take(src)
        cdf = pd.DataFrame(d, copy=False)
        if kwargs["count"]:
            cdf.insert(cdf.shape[1], kwargs["count_col"], cnt[a:b])
        results.append(cdf)
    return results


def _u_split(plan, kwargs):
    """split(between=True): pyranges/methods/split.py:4-24 per key -- Chromosome, Start, End[, Strand] of every pair
    of consecutive distinct Start/End points."""
    off, st, en = E.split(plan.f.iv, plan.group_ranges or None)
    st, en = st.cpu().numpy(), en.cpu().numpy()
    strand = kwargs.get("strand", False)
    results = []
    for g, df in enumerate(plan.frames):
        a, b = off[g], off[g + 1]
        if a == b:
            results.append(None)
            continue
        dtype = df.Start.dtype
        d = {"Chromosome": np.full(b - a, df.Chromosome.iloc[0], dtype=object), "Start": st[a:b].astype(dtype, copy=False),
             "End": en[a:b].astype(dtype, copy=False)}
        if strand and "Strand" in df:
            d["Strand"] = np.full(b - a, df["Strand"].iloc[0], dtype=object)
        results.append(pd.DataFrame(d, copy=False))
    return results


_UNARY = {"_split": _u_split, "_cluster": _u_cluster, "_merge": _u_merge, "coverage": _u_coverage, "_cluster_by": _u_cluster_by,
          "_merge_by": _u_merge_by}


def pyrange_apply_single(function, self, **kwargs):
    """Drop-in for pyranges.multithreaded.pyrange_apply_single (pyranges/multithreaded.py:306-372)."""
    name = _name(function)
    if name not in _UNARY:
        raise NotImplementedError(
            "pyranges_b200.pyrange_apply_single: %r is not on the B200 hot path (supported: %s); no CPU fallback"
            % (name, ", ".join(sorted(_UNARY))))
    if name in ("_cluster_by", "_merge_by"):
        plan = _ByPlan(self, kwargs["strand"], kwargs["by"], _device_of(kwargs))
    else:
        plan = _UnaryPlan(self, kwargs["strand"], _device_of(kwargs))
    results = _UNARY[name](plan, kwargs)
    return process_results(results, plan.out_keys)
