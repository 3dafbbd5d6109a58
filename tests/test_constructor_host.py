"""PyRanges(df): the vectorised dtype setting and grouping of the constructor (pyranges_b200/pyranges_main.py
_codes_of / _set_dtypes / _group_frames) against the plain pandas statement of pyranges/methods/init.py:10-32,140-150
(astype(str) -> category, groupby(observed=True)) on every kind of Chromosome / Strand column."""
import numpy as np
import pandas as pd
import pytest
from pandas.api.types import CategoricalDtype

from pyranges_b200.pyranges_main import PyRanges


def _plain(df):
    """the reference's construction, stated with pandas only"""
    df = df.copy().reset_index(drop=True)
    stranded = "Strand" in df and bool(df["Strand"].astype(str).isin(["+", "-"]).all())
    for c in ("Chromosome", "Strand"):
        if c in df:
            df[c] = df[c].astype(str)
    df["Start"] = df["Start"].astype(np.int64)
    df["End"] = df["End"].astype(np.int64)
    df["Chromosome"] = df["Chromosome"].astype("category")
    if "Strand" in df:
        import warnings

        with warnings.catch_warnings():
            warnings.simplefilter("ignore")
            df["Strand"] = df["Strand"].astype(CategoricalDtype([".", "-", "+"], ordered=True))
    by = ["Chromosome", "Strand"] if stranded else "Chromosome"
    return {k: v for k, v in df.groupby(by, observed=True)}


def _frame(n, chrom, strand, seed=0):
    rng = np.random.default_rng(seed)
    st = rng.integers(0, 10_000, n)
    d = {"Chromosome": chrom(rng, n), "Start": st, "End": st + rng.integers(1, 50, n), "Score": rng.integers(0, 9, n),
         "Name": np.asarray(["n%d" % i for i in range(n)], dtype=object)}
    if strand is not None:
        d["Strand"] = strand(rng, n)
    return pd.DataFrame(d)


NAMES = np.asarray(["chr1", "chr2", "chr10", "chrX", "chrUn_1"], dtype=object)
CHROM = {
    "object": lambda r, n: NAMES[r.integers(0, 5, n)],
    "category": lambda r, n: pd.Categorical(NAMES[r.integers(0, 5, n)]),
    "category_unused": lambda r, n: pd.Categorical(NAMES[r.integers(0, 3, n)], categories=list(NAMES[::-1]) + ["chrZ"]),
    "str_dtype": lambda r, n: pd.array(NAMES[r.integers(0, 5, n)], dtype="str"),
    "int": lambda r, n: r.integers(1, 23, n),
    "int_category": lambda r, n: pd.Categorical(r.integers(1, 12, n)),
    "mixed": lambda r, n: np.asarray([1, "1", "chr2", 2], dtype=object)[r.integers(0, 4, n)],
    "one": lambda r, n: np.asarray(["chr7"] * n, dtype=object),
}
STRAND = {
    "none": None,
    "pm": lambda r, n: np.asarray(["+", "-"], dtype=object)[r.integers(0, 2, n)],
    "pm_category": lambda r, n: pd.Categorical(np.asarray(["+", "-"], dtype=object)[r.integers(0, 2, n)]),
    "pm_category_unused": lambda r, n: pd.Categorical(np.asarray(["+", "-"], dtype=object)[r.integers(0, 2, n)],
                                                      categories=[".", "-", "+", "?"]),
    "dot": lambda r, n: np.asarray(["+", "-", "."], dtype=object)[r.integers(0, 3, n)],
    "other": lambda r, n: np.asarray(["+", "-", "x"], dtype=object)[r.integers(0, 3, n)],
    "nan": lambda r, n: np.asarray(["+", "-", np.nan], dtype=object)[r.integers(0, 3, n)],
    "plus_only": lambda r, n: np.asarray(["+"] * n, dtype=object),
}


@pytest.mark.parametrize("strand", sorted(STRAND))
@pytest.mark.parametrize("chrom", sorted(CHROM))
def test_constructor_matches_plain_pandas(chrom, strand):
    df = _frame(3000, CHROM[chrom], STRAND[strand], seed=len(chrom) * 31 + len(strand))
    got, want = PyRanges(df).dfs, _plain(df)
    assert list(got.keys()) == list(want.keys())
    for k in want:
        pd.testing.assert_frame_equal(got[k], want[k])


def test_constructor_keeps_a_custom_index_out():
    """the incoming index is dropped (init.py:128), the groups are indexed by row position"""
    df = _frame(500, CHROM["object"], STRAND["pm"], seed=3)
    df.index = np.arange(500)[::-1] * 7
    got, want = PyRanges(df).dfs, _plain(df)
    for k in want:
        pd.testing.assert_frame_equal(got[k], want[k])


def test_constructor_many_keys():
    """more than 2^15 (chromosome, strand) combinations: the wide sort key"""
    n = 70_000
    rng = np.random.default_rng(5)
    df = pd.DataFrame({"Chromosome": np.asarray(["c%d" % i for i in rng.integers(0, 20_000, n)], dtype=object),
                       "Start": np.arange(n), "End": np.arange(n) + 5,
                       "Strand": np.asarray(["+", "-"], dtype=object)[rng.integers(0, 2, n)]})
    got, want = PyRanges(df).dfs, _plain(df)
    assert list(got.keys()) == list(want.keys())
    for k in list(want)[:200]:
        pd.testing.assert_frame_equal(got[k], want[k])
