"""Generate tests/golden/*.npz by running the UNMODIFIED reference Python layer (/root/reference)
over the oracle shims (oracle/shims -> oracle/iv_oracle.c).

Run in the authoring container only (the reference does not travel to the GPU box):

    python oracle/make_golden.py

Every fixture stores the constructor DataFrames (row order preserved), the call (op + kwargs) and
the reference's result, so that tests can (a) re-check the oracle and (b) check the CUDA path
against the reference's own glue semantics (strandedness dispatch, empty partners, -1 fillers,
cluster-id offsets ...).  The script also replays the known answers of the reference's doctests
and unit tests for this path and aborts if the oracle misses one.

pandas-3 work-arounds: future.infer_string off; PyRanges rebuilt from a dict
without the empty (chrom, '.') groups; importlib.metadata.version patched; for pr.count_overlaps only,
Series.value_counts of an int32 Series returns int32 counts (pandas >= 3 refuses the int64 -> int32 setitem of
pyranges/methods/intersection.py:95; pandas < 3 down-cast silently).
"""
import importlib.metadata as md
import json
import os
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(HERE, "shims"))
sys.path.insert(0, "/root/reference")
_v = md.version
md.version = lambda n: "0.1.4" if n == "pyranges" else _v(n)

import numpy as np  # noqa: E402
import pandas as pd  # noqa: E402

pd.set_option("future.infer_string", False)
import pyranges as pr  # noqa: E402

OUT = os.path.join(ROOT, "tests", "golden")


def fix(gr):
    return pr.PyRanges({k: v for k, v in gr.dfs.items() if not v.empty})


def mk(df):
    """reference PyRanges from a constructor frame (row labels = range)"""
    df = df.reset_index(drop=True).copy()
    for c in ("Chromosome", "Strand"):
        if c in df.columns:
            df[c] = df[c].astype(str).astype(object)
    return fix(pr.PyRanges(df))


def _col(a):
    a = np.asarray(a)
    if a.dtype == object or a.dtype.kind in "OUS":
        return np.asarray([str(x) for x in a], dtype="U")
    return a


def _store_df(d, prefix, df):
    d[prefix + ".__cols__"] = np.asarray(list(df.columns), dtype="U")
    for c in df.columns:
        v = df[c]
        if str(v.dtype) == "category":
            v = v.astype(str)
        d["%s.%s" % (prefix, c)] = _col(v.values)


_saved_inputs = {}


def save_input(tag, df):
    """constructor frames are stored once per dataset (tests/golden/input_<tag>.npz)"""
    if tag in _saved_inputs:
        return
    d = {}
    _store_df(d, "df", df.reset_index(drop=True))
    np.savez_compressed(os.path.join(OUT, "input_%s.npz" % tag), **d)
    _saved_inputs[tag] = True


def save(name, op, kwargs, self_tag, other_tag, result, note=""):
    d = {}
    meta = {"case": name, "op": op, "kwargs": kwargs, "note": note, "self": self_tag, "other": other_tag}
    if op == "to_rle":
        keys = list(result.rles.keys())
        meta["rle_keys"] = [list(k) if isinstance(k, tuple) else k for k in keys]
        for i, k in enumerate(keys):
            d["rle.%d.runs" % i] = np.asarray(result.rles[k].runs, dtype=np.int64)
            d["rle.%d.values" % i] = np.asarray(result.rles[k].values, dtype=np.float64)
    else:
        rdf = result.df if len(result) > 0 else pd.DataFrame()
        meta["n_rows"] = int(len(rdf))
        _store_df(d, "result", rdf)
    d["meta"] = np.asarray(json.dumps(meta))
    np.savez_compressed(os.path.join(OUT, name + ".npz"), **d)
    return meta


def synth(seed, n, chroms, stranded=True, maxpos=20000, maxlen=400, nested=False, extra_cols=True):
    rng = np.random.default_rng(seed)
    c = rng.choice(chroms, n)
    s = rng.integers(0, maxpos, n)
    ln = rng.integers(1, maxlen, n)
    if nested:
        # 10% long containers, duplicates and bookended neighbours
        big = rng.random(n) < 0.1
        ln[big] = rng.integers(2000, 9000, big.sum())
        dup = rng.random(n) < 0.1
        src = rng.integers(0, n, n)
        s[dup], ln[dup], c[dup] = s[src[dup]], ln[src[dup]], c[src[dup]]
        bk = rng.random(n) < 0.1
        s[bk] = (s[src[bk]] + ln[src[bk]])
        c[bk] = c[src[bk]]
    df = pd.DataFrame({"Chromosome": c, "Start": s.astype(np.int64), "End": (s + ln).astype(np.int64)})
    if stranded:
        df["Strand"] = rng.choice(["+", "-"], n)
    if extra_cols:
        df["Score"] = rng.integers(0, 1000, n)
        df["Name"] = ["r%d" % i for i in range(n)]
    return df


def known_answers():
    """doctest / unit-test known answers of the reference for this path."""
    f1, f2 = fix(pr.data.f1()), fix(pr.data.f2())
    # join inner doctest pyranges_main.py:2328-2362 (f1 x f2, strand ignored): 1 row interval2 x b
    j = f1.join(f2).df
    assert j[["Start", "End", "Start_b", "End_b"]].values.tolist() == [[5, 7, 6, 7]], j
    # count_overlaps doctest :1338-1372 (strandedness same) -> [0, 0, 1]
    assert f1.count_overlaps(f2, strandedness="same").df.NumberOverlaps.tolist() == [0, 0, 1]
    # coverage doctest :1438-1474
    cv = f1.coverage(f2).df
    assert cv.NumberOverlaps.tolist() == [0, 0, 1] and cv.FractionOverlaps.tolist() == [0.0, 0.0, 0.5]
    # nearest doctests :3258-3309
    n = f1.nearest(f2).df
    assert n.Distance.tolist() == [1, 2, 0] and n.Start_b.tolist() == [6, 6, 6]
    n = f1.nearest(f2, how="upstream").df
    assert n.Distance.tolist() == [2, 2, 0] and n.Start_b.tolist() == [1, 6, 6], n
    # overlap doctests :3536-3605
    gr = pr.from_dict({"Chromosome": ["chr1", "chr1", "chr2", "chr1", "chr3"], "Start": [1, 1, 4, 10, 0],
                       "End": [3, 3, 9, 11, 1], "ID": ["a", "a", "b", "c", "d"]})
    gr2 = pr.from_dict({"Chromosome": ["chr1", "chr1", "chr2"], "Start": [2, 2, 1], "End": [3, 9, 10]})
    assert gr.overlap(gr2).df.ID.tolist() == ["a", "a", "b"]
    assert gr.overlap(gr2, how=None).df.ID.tolist() == ["a", "a", "a", "a", "b"]
    assert gr.overlap(gr2, how="containment").df.ID.tolist() == ["b"]
    assert gr.overlap(gr2, invert=True).df.ID.tolist() == ["c", "d"]
    # subtract doctest :4819-4860: [1,3) a, [4,9) b, [10,11) c minus [2,3), [2,9), [9,10) -> [1,2) a, [10,11) c
    sa = pr.from_dict({"Chromosome": ["chr1"] * 3, "Start": [1, 4, 10], "End": [3, 9, 11], "ID": ["a", "b", "c"]})
    sb = pr.from_dict({"Chromosome": ["chr1"] * 3, "Start": [2, 2, 9], "End": [3, 9, 10]})
    sub = sa.subtract(sb).df
    assert sub[["Start", "End", "ID"]].values.tolist() == [[1, 2, "a"], [10, 11, "c"]], sub
    # cluster doctests :1111-1184
    g = pr.from_dict({"Chromosome": [1, 1, 1, 1], "Start": [1, 2, 3, 9], "End": [3, 3, 10, 12],
                      "Gene": [1, 2, 3, 3]})
    assert g.cluster().df.Cluster.tolist() == [1, 1, 1, 1]
    assert g.cluster(slack=-1).df.Cluster.tolist() == [1, 1, 2, 2]
    # cluster(by=) doctests :1139-1184 (cluster_by restatement)
    gb = pr.from_dict({"Chromosome": [1, 1, 1, 1], "Start": [1, 2, 3, 9], "End": [3, 3, 10, 12], "Gene": [1, 2, 3, 3]})
    cb = gb.cluster(by="Gene", count=True).df
    assert cb.Cluster.tolist() == [1, 2, 3, 3] and cb.Count.tolist() == [1, 1, 2, 2]
    cb2 = fix(pr.data.ensembl_gtf())[["Feature", "Source"]].cluster(by=["Feature", "Source"]).df
    assert len(cb2) == 2446 and int(cb2.Cluster.max()) == 1145 and cb2.Cluster.tolist()[:4] == [1, 2, 2, 2]
    # merge(by=) doctests :3087-3125 (merge_by restatement)
    mb = fix(pr.data.ensembl_gtf()).merge(by="Feature", count=True).df
    assert len(mb) == 748 and mb.Count.tolist()[:4] == [1, 2, 1, 11] and mb.Start.tolist()[:2] == [65564, 69036]
    mb2 = fix(pr.data.ensembl_gtf()).merge(by=["Feature", "gene_name"], count=True).df
    assert len(mb2) == 807 and mb2.gene_name.tolist()[-2:] == ["WASH7P", "WASH9P"] and mb2.Count.tolist()[:4] == [1, 2, 2, 4]
    # split doctests :4386-4460
    sg = fix(pr.from_dict({"Chromosome": ["chr1"] * 4, "Start": [3, 5, 5, 11], "End": [6, 9, 7, 12],
                           "Strand": ["+", "+", "-", "-"]}))
    assert sg.split().df.Start.tolist() == [3, 5, 6, 5, 11] and sg.split().df.End.tolist() == [5, 6, 9, 7, 12]
    assert sg.split(between=True).df.Start.tolist() == [3, 5, 6, 5, 7, 11]
    assert sg.split(strand=False).df.Start.tolist() == [3, 5, 6, 7, 11]
    assert sg.split(strand=False, between=True).df.End.tolist() == [5, 6, 7, 9, 11, 12]
    # pr.count_overlaps doctest multioverlap.py:38-134
    ga = pr.from_dict({"Chromosome": ["chr1"] * 4, "Start": [6, 10, 22, 24], "End": [12, 20, 27, 30]})
    gb2 = pr.from_dict({"Chromosome": ["chr1"] * 2, "Start": [12, 14], "End": [32, 30]})
    gc = pr.from_dict({"Chromosome": ["chr1"] * 3, "Start": [8, 10, 32], "End": [15, 14, 34]})
    feat = pr.PyRanges(chromosomes=["chr1"] * 4, starts=[0, 10, 20, 30], ends=[10, 20, 30, 40])
    with int32_value_counts():
        cm = pr.count_overlaps({"a": ga, "b": gb2, "c": gc}, feat).df
        cm2 = pr.count_overlaps({"a": ga, "b": gb2, "c": gc}).df
    assert cm.a.tolist() == [1, 2, 2, 0] and cm.b.tolist() == [0, 2, 2, 1] and cm.c.tolist() == [1, 2, 0, 1], cm
    assert len(cm2) == 12, len(cm2)
    # merge doctest :3069-3086 (ensembl_gtf: 62 merged rows... first 11868-14409 count 12)
    eg = fix(pr.data.ensembl_gtf())
    m = eg.merge(count=True, count_col="Count").df
    first = m.iloc[0]
    assert (int(first.Start), int(first.End), int(first.Count)) == (11868, 14409, 12), first
    # to_rle doctests :5800-5872
    d = {"Chromosome": ["chr1", "chr1", "chr1"], "Start": [3, 8, 5], "End": [6, 9, 7], "Score": [0.1, 5, 3.14],
         "Strand": ["+", "+", "-"]}
    g = fix(pr.from_dict(d))
    r = g.to_rle()
    assert r[("chr1", "+")].runs.tolist() == [3, 3, 2, 1] and r[("chr1", "+")].values.tolist() == [0, 1, 0, 1]
    assert r[("chr1", "-")].runs.tolist() == [5, 2] and r[("chr1", "-")].values.tolist() == [0, 1]
    r = g.to_rle(value_col="Score")
    assert r[("chr1", "+")].values.tolist() == [0.0, 0.1, 0.0, 5.0]
    r = g.to_rle(value_col="Score", strand=False)
    assert r["chr1"].runs.tolist() == [3, 2, 1, 1, 1, 1]
    assert np.allclose(r["chr1"].values, [0.0, 0.1, 3.24, 3.14, 0.0, 5.0])
    r = g.to_rle(rpm=True)
    assert np.allclose(r[("chr1", "+")].values, [0.0, 333333.3333333333, 0.0, 333333.3333333333])
    # bundled chipseq known answers (brute force)
    cs, bg = fix(pr.data.chipseq()), fix(pr.data.chipseq_background())
    j = cs.join(bg).df
    assert sorted(j.Start.tolist()) == sorted([226987592, 38747226, 26105515])
    assert cs.join(bg, strandedness="same").df.Start.tolist() == [26105515]
    co = cs.count_overlaps(bg).df
    assert len(co) == 10000 and int(co.NumberOverlaps.sum()) == 3 and str(co.NumberOverlaps.dtype) == "int64"
    assert len(cs.merge(count=True)) == 9915 and int(cs.cluster().df.Cluster.max()) == 9915
    # unit tests: tests/unit/test_stranded.py:8-25 (cpg x exons)
    cpg, exons = fix(pr.data.cpg()), fix(pr.data.exons())
    j = cpg.join(exons)
    assert j.stranded
    # tests/unit/join/test_join.py:17-36 preserve_order
    gr1 = pr.from_dict({"Chromosome": ["chr1", "chr1", "chr1"], "Start": [3, 8, 5], "End": [6, 9, 7]})
    gr2 = pr.from_dict({"Chromosome": ["chr1", "chr1"], "Start": [1, 6], "End": [2, 7]})
    assert gr1.join(gr2, how="left", preserve_order=True).df.Start.tolist() == [3, 8, 5]
    # k_nearest doctests pyranges_main.py:2560-2700
    kf1 = pr.from_dict({"Chromosome": ["chr1"] * 3, "Start": [3, 8, 5], "End": [6, 9, 7], "Strand": ["+", "+", "-"]})
    kf2 = pr.from_dict({"Chromosome": ["chr1"] * 2, "Start": [1, 6], "End": [2, 7], "Strand": ["+", "-"]})
    cols = ["Start", "End", "Start_b", "End_b", "Distance"]
    assert kf1.k_nearest(kf2, k=2).df[cols].values.tolist() == [
        [3, 6, 6, 7, 1], [3, 6, 1, 2, -2], [8, 9, 6, 7, -2], [8, 9, 1, 2, -7], [5, 7, 6, 7, 0], [5, 7, 1, 2, 4]]
    assert kf1.k_nearest(kf2, how="upstream", k=2).df[cols].values.tolist() == [
        [3, 6, 1, 2, -2], [8, 9, 6, 7, -2], [8, 9, 1, 2, -7], [5, 7, 6, 7, 0]]
    assert kf1.k_nearest(kf2, k=[1, 2, 1]).df[cols].values.tolist() == [
        [3, 6, 6, 7, 1], [8, 9, 6, 7, -2], [8, 9, 1, 2, -7], [5, 7, 6, 7, 0]]
    kg = pr.from_dict({"Chromosome": [1], "Start": [5], "End": [6]})
    kg2 = pr.from_dict({"Chromosome": 1, "Start": [1] * 2 + [5] * 2 + [9] * 2, "End": [3] * 2 + [7] * 2 + [11] * 2,
                        "ID": range(6)})
    idd = lambda r: r.df[["ID", "Distance"]].values.tolist()  # noqa: E731
    assert idd(kg.k_nearest(kg2, k=2)) == [[2, 0], [3, 0]]
    assert idd(kg.k_nearest(kg2, k=2, ties="different")) == [[2, 0], [3, 0], [1, -3], [0, -3]]
    assert idd(kg.k_nearest(kg2, k=3, ties="first")) == [[2, 0], [1, -3], [4, 4]]
    assert idd(kg.k_nearest(kg2, k=1, overlap=False)) == [[1, -3]]
    # tests/unit/k_nearest.py:35-200 (the frames of the fixtures, column by column)
    ug = pr.PyRanges(chromosomes="chr1", starts=[1, 15, 200], ends=[10, 20, 2000])
    ug2 = pr.PyRanges(chromosomes="chr1", starts=[11, 11, 20, 20, 50], ends=[16, 20, 21, 22, 100])
    assert ug.k_nearest(ug2, ties="different", k=[3, 2, 1]).df[cols].values.tolist() == [
        [1, 10, 11, 16, 2], [1, 10, 11, 20, 2], [1, 10, 20, 21, 11], [1, 10, 20, 22, 11], [1, 10, 50, 100, 41],
        [15, 20, 11, 20, 0], [15, 20, 11, 16, 0], [15, 20, 20, 21, 1], [15, 20, 20, 22, 1], [200, 2000, 50, 100, -101]]
    # (rows of one interval follow an unstable sort_values("__IX__") when k varies over the rows: compared as a set)
    assert sorted(ug2.k_nearest(ug, ties="different", k=[1, 2, 3, 4, 2]).df[cols].values.tolist()) == sorted([
        [11, 16, 15, 20, 0], [11, 20, 15, 20, 0], [11, 20, 1, 10, -2], [20, 21, 15, 20, 1], [20, 21, 1, 10, -11],
        [20, 21, 200, 2000, 180], [20, 22, 15, 20, 1], [20, 22, 1, 10, -11], [20, 22, 200, 2000, 179],
        [50, 100, 15, 20, -31], [50, 100, 1, 10, -41]])
    assert ug.k_nearest(ug2, ties="first", k=[3, 2, 1]).df[cols].values.tolist() == [
        [1, 10, 11, 16, 2], [1, 10, 20, 21, 11], [1, 10, 50, 100, 41], [15, 20, 11, 20, 0], [15, 20, 20, 21, 1],
        [200, 2000, 50, 100, -101]]
    first2 = [[11, 16, 15, 20, 0], [11, 20, 15, 20, 0], [11, 20, 1, 10, -2], [20, 21, 15, 20, 1], [20, 21, 1, 10, -11],
              [20, 21, 200, 2000, 180], [20, 22, 15, 20, 1], [20, 22, 1, 10, -11], [20, 22, 200, 2000, 179],
              [50, 100, 15, 20, -31], [50, 100, 1, 10, -41], [50, 100, 200, 2000, 101]]
    assert ug2.k_nearest(ug, ties="first", k=list(range(1, 6))).df[cols].values.tolist() == first2
    assert ug.k_nearest(ug2, ties="last", k=[3, 2, 1]).df[cols].values.tolist() == [
        [1, 10, 11, 20, 2], [1, 10, 20, 22, 11], [1, 10, 50, 100, 41], [15, 20, 11, 20, 0], [15, 20, 20, 22, 1],
        [200, 2000, 50, 100, -101]]
    assert ug2.k_nearest(ug, ties="last", k=list(range(1, 6))).df[cols].values.tolist() == first2
    print("known answers of the reference doctests/unit tests: all reproduced by reference+oracle")


def run_case(name, op, kwargs, sdf, odf=None, note=""):
    stag, sdf = sdf
    save_input(stag, sdf)
    otag = None
    if odf is not None:
        otag, odf = odf
        save_input(otag, odf)
    s = mk(sdf)
    o = mk(odf) if odf is not None else None
    kw = dict(kwargs)
    if op in ("join", "overlap", "count_overlaps", "coverage", "nearest", "k_nearest", "intersect", "set_intersect", "set_union",
              "subtract"):
        try:
            res = getattr(s, op)(o, **kw)
        except AssertionError as exc:
            if op != "set_union":
                raise
            # the reference's own pr.concat trips over pandas >= 3 for some stranded inputs: no fixture then
            print("%-44s %-15s SKIPPED (reference fails here: %s)" % (name, op, str(exc)[:60]))
            return
    else:
        res = getattr(s, op)(**kw)
    m = save(name, op, kwargs, stag, otag, res, note)
    print("%-44s %-15s rows=%s" % (name, op, m.get("n_rows", "rle")))


class int32_value_counts:
    """pandas >= 3 only: see the module docstring."""

    def __enter__(self):
        self._vc = vc = pd.Series.value_counts

        def value_counts(series, *a, **k):
            r = vc(series, *a, **k)
            return r.astype(np.int32) if series.dtype == np.int32 else r

        pd.Series.value_counts = value_counts

    def __exit__(self, *exc):
        pd.Series.value_counts = self._vc


def run_multi(name, samples, features, kwargs):
    """pr.count_overlaps(grs, features) (pyranges/multioverlap.py:6-165); samples: {column name: (tag, frame)}"""
    for tag, df in samples.values():
        save_input(tag, df)
    ftag = None
    if features is not None:
        ftag, fdf = features
        save_input(ftag, fdf)
    grs = {n: mk(df) for n, (tag, df) in samples.items()}
    with int32_value_counts():
        res = pr.count_overlaps(grs, mk(features[1]) if features is not None else None, **kwargs)
    d = {}
    meta = {"case": name, "op": "count_matrix", "kwargs": kwargs, "note": "", "self": ftag, "other": None,
            "samples": {n: tag for n, (tag, _) in samples.items()}}
    rdf = res.df
    meta["n_rows"] = int(len(rdf))
    _store_df(d, "result", rdf)
    d["meta"] = np.asarray(json.dumps(meta))
    np.savez_compressed(os.path.join(OUT, name + ".npz"), **d)
    print("%-44s %-15s rows=%s" % (name, "count_matrix", len(rdf)))


def main():
    os.makedirs(OUT, exist_ok=True)
    for f in os.listdir(OUT):
        if f.endswith(".npz"):
            os.remove(os.path.join(OUT, f))
    known_answers()

    f1 = pr.data.f1().df
    f2 = pr.data.f2().df
    cs = pr.data.chipseq().df
    bg = pr.data.chipseq_background().df
    chroms = ["chr1", "chr2", "chr10", "chrX"]
    A = synth(11, 3000, chroms, nested=True)
    B = synth(12, 700, chroms + ["chrY"], nested=True)
    Au = synth(13, 1500, chroms, stranded=False, nested=True)
    Bu = synth(14, 400, chroms[:3], stranded=False, nested=True)
    pile = synth(15, 300, ["chr1"], maxpos=200, maxlen=150)          # deep pile-up: > 100 hits per query
    pileq = synth(16, 200, ["chr1"], maxpos=200, maxlen=150)

    f1, f2, cs, bg = ("f1", f1), ("f2", f2), ("chipseq", cs), ("chipseq_background", bg)
    A, B, Au, Bu = ("A", A), ("B", B), ("Au", Au), ("Bu", Bu)
    pile, pileq = ("pile", pile), ("pileq", pileq)
    binary_sets = {
        "f1f2": (f1, f2),
        "chipseq": (cs, bg),
        "synth": (A, B),
        "synth_us": (Au, B),      # self unstranded, other stranded
        "synth_su": (A, Bu),      # self stranded, other unstranded
        "synth_uu": (Au, Bu),
        "pile": (pileq, pile),
    }
    for tag, (s, o) in binary_sets.items():
        both_stranded = "Strand" in s[1].columns and "Strand" in o[1].columns
        strands = [None, "same", "opposite"] if both_stranded else [None]
        for st in strands:
            sn = {None: "none", "same": "same", "opposite": "opp"}[st]
            run_case("join_%s_%s" % (tag, sn), "join", {"strandedness": st}, s, o)
            run_case("count_%s_%s" % (tag, sn), "count_overlaps", {"strandedness": st}, s, o)
            run_case("overlap_first_%s_%s" % (tag, sn), "overlap", {"strandedness": st}, s, o)
            run_case("nearest_%s_%s" % (tag, sn), "nearest", {"strandedness": st}, s, o)
            run_case("coverage_%s_%s" % (tag, sn), "coverage", {"strandedness": st}, s, o)
        if tag in ("f1f2", "synth", "synth_uu", "pile"):
            run_case("join_left_%s" % tag, "join", {"how": "left"}, s, o)
            run_case("join_right_%s" % tag, "join", {"how": "right"}, s, o)
            run_case("join_outer_%s" % tag, "join", {"how": "outer"}, s, o)
            run_case("join_left_po_%s" % tag, "join", {"how": "left", "preserve_order": True}, s, o)
            run_case("join_slack_%s" % tag, "join", {"slack": 5, "report_overlap": True}, s, o)
            run_case("join_contain_%s" % tag, "join", {"how": "containment"}, s, o)
            run_case("join_first_%s" % tag, "join", {"how": "first"}, s, o)
            run_case("join_last_%s" % tag, "join", {"how": "last"}, s, o)
            run_case("overlap_all_%s" % tag, "overlap", {"how": None}, s, o)
            run_case("overlap_contain_%s" % tag, "overlap", {"how": "containment"}, s, o)
            run_case("overlap_invert_%s" % tag, "overlap", {"invert": True}, s, o)
            run_case("count_drop_%s" % tag, "count_overlaps", {"keep_nonoverlapping": False, "overlap_col": "C"}, s, o)
            run_case("nearest_nooverlap_%s" % tag, "nearest", {"overlap": False}, s, o)
        # compositions of pairs / merge
        for st in strands:
            sn = {None: "none", "same": "same", "opposite": "opp"}[st]
            run_case("intersect_%s_%s" % (tag, sn), "intersect", {"strandedness": st}, s, o)
            run_case("set_intersect_%s_%s" % (tag, sn), "set_intersect", {"strandedness": st}, s, o)
            run_case("set_union_%s_%s" % (tag, sn), "set_union", {"strandedness": st}, s, o)
            run_case("subtract_%s_%s" % (tag, sn), "subtract", {"strandedness": st}, s, o)
        if tag in ("f1f2", "synth", "synth_uu", "pile"):
            run_case("intersect_first_%s" % tag, "intersect", {"how": "first"}, s, o)
            run_case("intersect_last_%s" % tag, "intersect", {"how": "last"}, s, o)
            run_case("intersect_contain_%s" % tag, "intersect", {"how": "containment"}, s, o)
            run_case("intersect_invert_%s" % tag, "intersect", {"invert": True}, s, o)
            run_case("set_intersect_contain_%s" % tag, "set_intersect", {"how": "containment"}, s, o)
        if both_stranded:
            for how in ("upstream", "downstream"):
                for st in ("same", None, "opposite"):
                    sn = {None: "none", "same": "same", "opposite": "opp"}[st]
                    run_case("nearest_%s_%s_%s" % (how, tag, sn), "nearest", {"how": how, "strandedness": st}, s, o)
                run_case("nearest_%s_nooverlap_%s" % (how, tag), "nearest",
                         {"how": how, "strandedness": "same", "overlap": False}, s, o)

    # k_nearest (small sets: the oracle's k-nearest walks are plain Python)
    Ks = ("Ks", synth(21, 400, chroms[:2], maxpos=4000, maxlen=120, nested=True))
    Ko = ("Ko", synth(22, 250, chroms[:3], maxpos=4000, maxlen=120, nested=True))
    Ksu = ("Ksu", synth(23, 300, chroms[:2], stranded=False, maxpos=4000, maxlen=120, nested=True))
    Kou = ("Kou", synth(24, 200, chroms[:2], stranded=False, maxpos=4000, maxlen=120, nested=True))
    for tag, (s, o) in {"f1f2": (f1, f2), "ksyn": (Ks, Ko), "ksyn_uu": (Ksu, Kou), "ksyn_us": (Ksu, Ko),
                        "pile": (pileq, pile)}.items():
        both_stranded = "Strand" in s[1].columns and "Strand" in o[1].columns
        per_row = [int(i % 4) + 1 for i in range(len(s[1]))]
        for ties in (None, "first", "last", "different"):
            tn = ties or "none"
            run_case("knear_k1_%s_%s" % (tn, tag), "k_nearest", {"k": 1, "ties": ties}, s, o)
            run_case("knear_k3_%s_%s" % (tn, tag), "k_nearest", {"k": 3, "ties": ties}, s, o)
            run_case("knear_rowk_%s_%s" % (tn, tag), "k_nearest", {"k": per_row, "ties": ties}, s, o)
            run_case("knear_k2_nooverlap_%s_%s" % (tn, tag), "k_nearest", {"k": 2, "ties": ties, "overlap": False}, s, o)
        if both_stranded:
            for st in ("same", "opposite"):
                run_case("knear_k2_%s_%s" % (st, tag), "k_nearest", {"k": 2, "strandedness": st}, s, o)
            for how in ("upstream", "downstream"):
                run_case("knear_k2_%s_%s" % (how, tag), "k_nearest", {"k": 2, "how": how, "ties": "different"}, s, o)
                run_case("knear_k2_%s_same_%s" % (how, tag), "k_nearest",
                         {"k": 2, "how": how, "strandedness": "same", "overlap": False}, s, o)

    unary_sets = {"f1": f1, "chipseq": cs, "synth": A, "synth_u": Au, "pile": pile}
    for tag, s in unary_sets.items():
        stranded = "Strand" in s[1].columns
        for strand in ([None, False] if stranded else [None]):
            sn = "dflt" if strand is None else "nostrand"
            kw = {} if strand is None else {"strand": strand}
            run_case("cluster_%s_%s" % (tag, sn), "cluster", dict(kw), s)
            run_case("cluster_count_%s_%s" % (tag, sn), "cluster", dict(kw, count=True), s)
            run_case("cluster_slack_%s_%s" % (tag, sn), "cluster", dict(kw, slack=10), s)
            run_case("merge_%s_%s" % (tag, sn), "merge", dict(kw), s)
            run_case("merge_count_%s_%s" % (tag, sn), "merge", dict(kw, count=True, count_col="N", slack=3), s)
            run_case("rle_%s_%s" % (tag, sn), "to_rle", dict(kw), s)
            run_case("rle_rpm_%s_%s" % (tag, sn), "to_rle", dict(kw, rpm=True), s)
    # negative slack: distinct starts only (tie order among equal starts is unpinned upstream)
    uniq = ("Auniq", A[1].drop_duplicates(["Chromosome", "Start"]))
    run_case("cluster_negslack_synth", "cluster", {"slack": -5}, uniq)
    run_case("merge_negslack_synth", "merge", {"slack": -5, "count": True}, uniq)
    # cluster / merge by=
    rng = np.random.default_rng(17)
    Gd = A[1].copy()
    Gd["Gene"] = rng.integers(0, 40, len(Gd))
    Gd["Kind"] = rng.choice(["exon", "cds", "utr"], len(Gd))
    G = ("G", Gd)
    Gu = ("Gu", Gd.drop("Strand", axis=1))
    for tag, s in (("G", G), ("Gu", Gu)):
        for strand in ([None, False] if tag == "G" else [None]):
            sn = "dflt" if strand is None else "nostrand"
            kw = {} if strand is None else {"strand": strand}
            run_case("cluster_by_%s_%s" % (tag, sn), "cluster", dict(kw, by="Gene"), s)
            run_case("cluster_by2_%s_%s" % (tag, sn), "cluster", dict(kw, by=["Kind", "Gene"], count=True, slack=10), s)
            run_case("merge_by_%s_%s" % (tag, sn), "merge", dict(kw, by="Gene"), s)
            run_case("merge_by2_%s_%s" % (tag, sn), "merge", dict(kw, by=["Kind", "Gene"], count=True, slack=3), s)
    # split and the multi-sample count matrix
    for tag, s in (("f1", f1), ("synth", A), ("synth_u", Au), ("pile", pile)):
        run_case("split_%s" % tag, "split", {}, s)
        run_case("split_between_%s" % tag, "split", {"between": True}, s)
        if "Strand" in s[1].columns:
            run_case("split_nostrand_%s" % tag, "split", {"strand": False, "between": True}, s)
    run_multi("matrix_synth_feat", {"a": A, "b": B, "c": Au}, Bu, {})
    run_multi("matrix_synth_same", {"a": A, "b": B}, A, {"strandedness": "same"})
    run_multi("matrix_synth_contain", {"a": A, "b": B}, B, {"how": "containment"})
    run_multi("matrix_synth_nofeat", {"a": Au, "b": Bu}, None, {})
    run_multi("matrix_f1f2_nofeat", {"x": f1, "y": f2}, None, {})
    run_case("rle_value_synth", "to_rle", {"value_col": "Score"}, A,
             note="float weights: compared with tolerance, not bit-exact (SURVEY.md section 8(c)(iv))")
    print("fixtures written to", OUT)


if __name__ == "__main__":
    main()
