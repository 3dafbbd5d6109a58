"""Loader / comparer of the golden fixtures under tests/golden/ (made by oracle/make_golden.py from the
UNMODIFIED reference Python layer running over the oracle shims)."""
import glob
import json
import os

import numpy as np
import pandas as pd

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def case_names():
    return sorted(os.path.basename(p)[:-4] for p in glob.glob(os.path.join(GOLDEN, "*.npz"))
                  if not os.path.basename(p).startswith("input_"))


def _load_df(z, prefix):
    key = prefix + ".__cols__"
    if key not in z.files:
        return pd.DataFrame()
    d = {}
    for c in z[key]:
        v = z["%s.%s" % (prefix, c)]
        d[str(c)] = v.astype(object) if v.dtype.kind == "U" else v
    return pd.DataFrame(d)


_inputs = {}


def load_input(tag):
    if tag not in _inputs:
        with np.load(os.path.join(GOLDEN, "input_%s.npz" % tag), allow_pickle=False) as z:
            _inputs[tag] = _load_df(z, "df")
    return _inputs[tag].copy()


def load_case(name):
    with np.load(os.path.join(GOLDEN, name + ".npz"), allow_pickle=False) as z:
        meta = json.loads(str(z["meta"]))
        if meta["op"] == "to_rle":
            res = {}
            for i, k in enumerate(meta["rle_keys"]):
                k = tuple(k) if isinstance(k, list) else k
                res[k] = (z["rle.%d.runs" % i], z["rle.%d.values" % i])
        else:
            res = _load_df(z, "result")
    return meta, res


BINARY_OPS = ("join", "overlap", "count_overlaps", "coverage", "nearest", "k_nearest", "intersect", "set_intersect",
              "set_union", "subtract")


def run_case(meta, PyRanges):
    if meta["op"] == "count_matrix":  # pr.count_overlaps(grs, features)
        from pyranges_b200.multioverlap import count_overlaps

        grs = {n: PyRanges(load_input(tag)) for n, tag in meta["samples"].items()}
        feats = PyRanges(load_input(meta["self"])) if meta["self"] is not None else None
        return count_overlaps(grs, feats, **meta["kwargs"]), feats, None
    s = PyRanges(load_input(meta["self"]))
    kw = dict(meta["kwargs"])
    if meta["other"] is not None:
        o = PyRanges(load_input(meta["other"]))
        return getattr(s, meta["op"])(o, **kw), s, o
    return getattr(s, meta["op"])(**kw), s, None


def _norm(df):
    """canonical form: categories/objects as str, numbers as float64 when they hold NaN-free ints too."""
    out = {}
    for c in df.columns:
        v = df[c]
        if str(v.dtype) == "category" or v.dtype == object or str(v.dtype) in ("str", "string"):
            out[c] = v.astype(str).values.astype(object)
        else:
            out[c] = np.asarray(v.values, dtype=np.float64)
    return pd.DataFrame(out)


def canon(df, cols=None):
    n = _norm(df if cols is None else df[cols])
    if len(n) == 0:
        return n
    return n.sort_values(list(n.columns), kind="stable").reset_index(drop=True)


def assert_frames_equal(got, want, what=""):
    assert list(got.columns) == list(want.columns), "%s: columns %s != %s" % (what, list(got.columns), list(want.columns))
    assert len(got) == len(want), "%s: %d rows != %d" % (what, len(got), len(want))
    g, w = canon(got), canon(want)
    for c in g.columns:
        if g[c].dtype.kind == "f":
            a, b = g[c].to_numpy(dtype=np.float64), w[c].to_numpy(dtype=np.float64)
            assert np.array_equal(a, b, equal_nan=True), "%s: column %s differs" % (what, c)
        else:
            a, b = g[c].to_numpy(dtype=object), w[c].to_numpy(dtype=object)
            assert (a == b).all(), "%s: column %s differs" % (what, c)


def check_case(name, PyRanges):
    meta, want = load_case(name)
    got, s, o = run_case(meta, PyRanges)
    op = meta["op"]
    if op == "to_rle":
        assert set(got.keys()) == set(want.keys()), name
        for k, (runs, vals) in want.items():
            assert np.array_equal(got[k].runs, runs), (name, k)
            if meta["kwargs"].get("value_col"):
                assert np.allclose(got[k].values, vals, rtol=1e-9, atol=1e-9), (name, k)
            else:
                assert np.array_equal(got[k].values, vals), (name, k)
        return
    gdf = got.df
    if len(want) == 0:
        assert len(gdf) == 0, name
        return
    if op == "nearest":
        # tie-breaks between equally near subjects are unpinned upstream: compare the
        # self columns + Distance exactly, then check that OUR partner really sits at that distance
        assert list(gdf.columns) == list(want.columns), (name, list(gdf.columns), list(want.columns))
        dcol = [c for c in want.columns if c.startswith("Distance")][-1]
        own = [c for c in want.columns if not c.endswith("_b") or c == dcol]
        own = [c for c in own if c in s.columns or c == dcol]
        assert_frames_equal(gdf[own], want[own], name)
        S, E, Sb, Eb, D = (gdf[c].values.astype(np.int64) for c in ("Start", "End", "Start_b", "End_b", dcol))
        gap = np.where(Eb <= S, S - Eb + 1, np.where(Sb >= E, Sb - E + 1, 0))
        assert np.array_equal(gap, D), name
        # where the distance pins the partner coordinate, it must match the reference's
        return
    if op == "k_nearest":
        assert list(gdf.columns) == list(want.columns), (name, list(gdf.columns), list(want.columns))
        # every row: the partner sits at |Distance| in front of / behind the interval, the sign follows
        # pyranges_main.py:2802-2813 (negative in front of the interval, mirrored on the - strand)
        S, E, Sb, Eb, D = (gdf[c].values.astype(np.int64) for c in ("Start", "End", "Start_b", "End_b", "Distance"))
        gap = np.where(Eb <= S, S - Eb + 1, np.where(Sb >= E, Sb - E + 1, 0))
        assert np.array_equal(gap, np.abs(D)), name
        minus = (gdf["Strand"].values.astype(str) == "-") if "Strand" in gdf.columns else np.zeros(len(gdf), bool)
        assert np.array_equal(D < 0, ((Eb < S) != minus) & (D != 0)), name
        if meta["kwargs"].get("ties") == "different":
            assert_frames_equal(gdf, want, name)  # every subject of the k nearest distinct distances: a set
            return
        # ties None / first / last keep ONE of several equally near subjects (the k-th row; one per distance): which one
        # follows the traversal order of ncls / an unstable sort upstream.  The interval's own columns
        # and the distances are pinned.
        own = [c for c in want.columns if c in s.columns or c == "Distance"]
        if "Strand" in o.columns and "Strand" not in s.columns:
            # an unstranded self picks up other's Strand and the reference then chooses the k nearest PER STRAND OF THE
            # NEIGHBOUR (PyRanges(dict) regroups the frames): which of two equally near neighbours is met first decides
            # the group a row lands in, i.e. even the number of rows.  Pinned here: the intervals that get neighbours and
            # the distance of each one's nearest neighbour.
            base = [c for c in own if c != "Distance"]
            g = gdf.assign(D=np.abs(gdf["Distance"].values)).groupby(base, sort=True).D.min().reset_index()
            w = want.assign(D=np.abs(want["Distance"].values)).groupby(base, sort=True).D.min().reset_index()
            assert_frames_equal(g, w, name)
            return
        assert_frames_equal(gdf[own], want[own], name)
        return
    if op == "cluster":
        # row order inside a key follows an unstable sort upstream; ids are pinned
        assert_frames_equal(gdf, want, name)
        return
    assert_frames_equal(gdf, want, name)
