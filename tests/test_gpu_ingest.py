"""GPU: read_bed with the parse on the device (pyranges_b200/readers.py, csrc/ingest.cu) against pandas.read_csv --
what the reference's read_bed calls (pyranges/readers.py:133-139) -- on generated BED files: BED3 / BED6 / BED12, with
and without header, with and without a final newline, CRLF line ends, nrows, gzip; and the seeded device columns
against the frames of the PyRanges."""
import gzip
import os

import numpy as np
import pandas as pd
import pytest

pytestmark = pytest.mark.gpu

COLS = ("Chromosome Start End Name Score Strand ThickStart ThickEnd ItemRGB BlockCount BlockSizes BlockStarts").split()


def _bed(tmp_path, n, ncols, header=False, final_newline=True, crlf=False, gz=False, seed=0, float_score=False):
    rng = np.random.default_rng(seed)
    chroms = rng.choice(["chr1", "chr2", "chr10", "chrX", "chrUn_gl000220"], n)
    st = rng.integers(0, 5_000_000, n)
    en = st + rng.integers(1, 4000, n)
    rows = []
    for i in range(n):
        f = [chroms[i], str(st[i]), str(en[i]), "name_%d" % (i % 97), ("%.2f" % (i / 7)) if float_score else str(i % 1000),
             "+-"[i % 2], str(st[i] + 1), str(en[i] - 1), "255,0,%d" % (i % 255), str(2), "10,20,", "0,%d," % (i % 50)]
        rows.append("\t".join(f[:ncols]))
    nl = "\r\n" if crlf else "\n"
    text = nl.join(rows) + (nl if final_newline else "")
    if header:
        text = "\t".join(COLS[:ncols]) + nl + text
    p = os.path.join(str(tmp_path), "t%d_%d.bed%s" % (n, ncols, ".gz" if gz else ""))
    with (gzip.open(p, "wt", newline="") if gz else open(p, "w", newline="")) as fh:
        fh.write(text)
    return p


def _want(path, header, nrows=None):
    """what the reference's read_bed hands to PyRanges (pyranges/readers.py:118-139)"""
    with (gzip.open(path, "rt") if path.endswith(".gz") else open(path)) as fh:
        ncols = len(fh.readline().rstrip("\r\n").split("\t"))
    return pd.read_csv(path, dtype={"Chromosome": "category", "Strand": "category"}, nrows=nrows,
                       header=0 if header else None, names=COLS[:ncols], sep="\t")


def _same(got, want):
    assert list(got.columns) == list(want.columns)
    assert len(got) == len(want)
    for c in want.columns:
        a, b = got[c], want[c]
        if str(b.dtype) == "category" or b.dtype == object or str(b.dtype) in ("str", "string"):
            assert str(a.dtype) == str(b.dtype), (c, a.dtype, b.dtype)
            txt = lambda x: np.asarray(x.astype(object).where(x.notna(), "<missing>"), dtype=str)  # noqa: E731
            assert (txt(a) == txt(b)).all(), c
        else:
            assert a.dtype == b.dtype, (c, a.dtype, b.dtype)
            assert np.array_equal(a.values, b.values), c


@pytest.mark.parametrize("n,ncols,kw", [
    (1, 3, {}), (1000, 3, {}), (70000, 6, {}), (70000, 6, {"header": True}), (5000, 6, {"final_newline": False}),
    (5000, 6, {"crlf": True}), (3000, 12, {}), (4000, 6, {"gz": True}), (2000, 5, {"float_score": True}),
    (300000, 6, {"seed": 5}),
])
def test_read_bed_frame(tmp_path, n, ncols, kw):
    from pyranges_b200.readers import read_bed

    path = _bed(tmp_path, n, ncols, **kw)
    header = kw.get("header", False)
    _same(read_bed(path, as_df=True), _want(path, header))
    _same(read_bed(path, as_df=True, nrows=7), _want(path, header, nrows=7))


def test_read_bed_pyranges_and_device_columns(tmp_path):
    """the PyRanges equals the one built from the pandas frame; its device columns are the cached ingest (no upload on
    the first operation) and hold exactly the rows of its frames, key by key"""
    from pyranges_b200 import multithreaded as M
    from pyranges_b200.pyranges_main import PyRanges
    from pyranges_b200.readers import read_bed

    path = _bed(tmp_path, 50000, 6, seed=3)
    M.clear_device_cache()
    gr = read_bed(path)
    want = PyRanges(_want(path, False))
    assert gr.keys() == want.keys()
    for k in gr.keys():
        _same(gr.dfs[k].reset_index(drop=True), want.dfs[k].reset_index(drop=True))
        # the frames are the constructor's: same dtypes (ordered Strand categories included), and the index holds the
        # rows' positions in the file
        assert gr.dfs[k].dtypes.equals(want.dfs[k].dtypes), (k, gr.dfs[k].dtypes, want.dfs[k].dtypes)
        assert np.array_equal(gr.dfs[k].index.values, want.dfs[k].index.values), k
    _same(gr.df, want.df)
    assert len(M._FRAME_CACHE) == 1
    f = M._frames_of(gr, next(iter(M._FRAME_CACHE.values()))[0].iv.device)
    assert f is next(iter(M._FRAME_CACHE.values()))[0]  # served from the cache: nothing was uploaded
    s, e = f.iv.start.cpu().numpy(), f.iv.end.cpu().numpy()
    for i, (k, d) in enumerate(zip(f.keys, f.dfs)):
        a, b = f.seg_off[i], f.seg_off[i + 1]
        assert np.array_equal(s[a:b], d.Start.values) and np.array_equal(e[a:b], d.End.values), k
    # and an operation on it agrees with the same operation on the pandas-built object
    got, ref = gr.cluster().df, want.cluster().df
    assert np.array_equal(np.sort(got.Cluster.values), np.sort(ref.Cluster.values))
    assert len(gr.join(want)) == len(want.join(want))


def test_read_bed_rejects_bad_rows(tmp_path):
    from pyranges_b200.readers import read_bed

    p = os.path.join(str(tmp_path), "bad.bed")
    with open(p, "w") as fh:
        fh.write("chr1\t10\t20\nchr1\t30\tforty\n")
    with pytest.raises(ValueError):
        read_bed(p)


@pytest.mark.parametrize("strands,stranded", [("+-", True), ("+-.", False), ("+-x", False)])
def test_read_bed_strand_values(tmp_path, strands, stranded):
    """'.' strands leave an unstranded PyRanges keyed by chromosome; characters outside the reference's Strand dtype go
    through the constructor -- either way the object equals PyRanges(pandas frame)"""
    from pyranges_b200.pyranges_main import PyRanges
    from pyranges_b200.readers import read_bed

    rng = np.random.default_rng(11)
    n = 20000
    st = rng.integers(0, 1_000_000, n)
    p = os.path.join(str(tmp_path), "s.bed")
    with open(p, "w") as fh:
        for i in range(n):
            fh.write("chr%d\t%d\t%d\tn%d\t%d\t%s\n" % (rng.integers(1, 30), st[i], st[i] + 50, i, i % 9,
                                                        strands[i % len(strands)]))
    gr, want = read_bed(p), PyRanges(_want(p, False))
    assert gr.stranded == stranded == want.stranded
    assert gr.keys() == want.keys()
    for k in gr.keys():
        _same(gr.dfs[k].reset_index(drop=True), want.dfs[k].reset_index(drop=True))
        assert np.array_equal(gr.dfs[k].index.values, want.dfs[k].index.values), k
    assert len(gr.join(want)) == len(want.join(want))


def test_read_bed_text_columns(tmp_path):
    """text columns of every shape the host assembly treats differently: short with few distinct values (dictionary on
    the device), short and all distinct, wider than 8 bytes, longer than 64 bytes, numeric text in a late column"""
    from pyranges_b200.readers import read_bed

    rng = np.random.default_rng(12)
    n = 30000
    st = rng.integers(0, 1_000_000, n)
    p = os.path.join(str(tmp_path), "t.bed")
    with open(p, "w") as fh:
        for i in range(n):
            f = ["chr%d" % (i % 3), str(st[i]), str(st[i] + 9), "g%d" % (i % 50), str(i % 7), "+-"[i % 2],
                 "u%06d" % i, "%d" % (i * 3), "a_rather_long_identifier_%d" % i, "2", "7," * (1 + i % 40), "0.5"]
            fh.write("\t".join(f) + "\n")
    _same(read_bed(p, as_df=True), _want(p, False))
    _same(read_bed(p).df, _want(p, False).sort_values(["Chromosome", "Strand"], kind="stable").reset_index(drop=True))
