"""``PyRanges`` -- the reference's container API for the interval hot path, on the B200 engine.

Same layout as pyranges v0 (pyranges/pyranges_main.py:42, pyranges/methods/init.py:35-45,105-169): a dict
``dfs`` {chromosome | (chromosome, strand) -> pandas.DataFrame} with int64 Start/End.  The hot-path
methods keep the reference's signatures, kwargs and result layout and route through
``pyranges_b200.multithreaded.pyrange_apply`` / ``pyrange_apply_single`` (batched CUDA kernels):

    join (:2248) overlap (:3490) count_overlaps (:1293) coverage (:1388) nearest (:3192)
    cluster (:1067) merge (:3005) to_rle (:5772)

Everything else of the reference class (printing, I/O, statistics, tiling ...) is out of scope
(SURVEY.md section 8); the few container helpers those eight methods need are provided.
"""
import sys

import numpy as np
import pandas as pd

from ._natsort import natsorted
from .multithreaded import pyrange_apply, pyrange_apply_single


def fill_kwargs(kwargs):
    """pyranges/pyranges_main.py:23-39"""
    defaults = {
        "strandedness": None,
        "overlap": True,
        "how": None,
        "invert": None,
        "new_pos": None,
        "suffixes": ["_a", "_b"],
        "suffix": "_b",
        "sparse": {"self": False, "other": False},
    }
    defaults.update(kwargs)
    return defaults


def _codes_of(col):
    """(codes, names): the column as small-integer codes into the str() of its distinct values, without stringifying
    row by row; None when str() would merge values or a missing value would become the text 'nan' (the callers then
    take the row-by-row route)."""
    if isinstance(col.dtype, pd.CategoricalDtype):
        codes, values = col.cat.codes.to_numpy(), list(col.cat.categories)
    else:
        if col.dtype.kind in "iufb":
            return None
        codes, values = pd.factorize(col)
        values = list(values)
    if len(codes) and codes.min() < 0:
        return None
    names = [str(v) for v in values]
    if len(set(names)) != len(names):
        return None
    return codes, names


def _strand_dtype():
    from pandas.api.types import CategoricalDtype

    return CategoricalDtype([".", "-", "+"], ordered=True)


def _set_dtypes(df, coded=None):
    """pyranges/methods/init.py:10-32: Start/End int64, Chromosome category, Strand ordered category.  `coded` holds
    the (codes, names) of Chromosome / Strand where _codes_of could produce them: the categoricals are then built
    from the codes (same result as astype(str).astype(dtype), which walks every row twice)."""
    coded = coded or {}
    df["Start"] = df["Start"].astype(np.int64)
    df["End"] = df["End"].astype(np.int64)
    c = coded.get("Chromosome")
    if c is not None:
        codes, names = c
        used = np.zeros(len(names), bool)
        used[codes] = True
        cats = sorted(n for n, u in zip(names, used) if u)
        where = {n: i for i, n in enumerate(cats)}
        lut = np.asarray([where.get(n, -1) for n in names] + [-1], dtype=np.int64)
        df["Chromosome"] = pd.Categorical.from_codes(lut[codes], categories=cats)
    else:
        df["Chromosome"] = df["Chromosome"].astype(str).astype("category")
    if "Strand" in df:
        c = coded.get("Strand")
        if c is not None:
            codes, names = c
            where = {".": 0, "-": 1, "+": 2}
            lut = np.asarray([where.get(n, -1) for n in names] + [-1], dtype=np.int64)
            df["Strand"] = pd.Categorical.from_codes(lut[codes], dtype=_strand_dtype())
        else:
            df["Strand"] = df["Strand"].astype(str).astype(_strand_dtype())
    return df


def _check_strandedness(df, coded=None):
    """pyranges/methods/init.py:73-102: stranded iff a Strand column holds only '+' / '-'."""
    if "Strand" not in df:
        return False
    c = (coded or {}).get("Strand")
    if c is not None:
        codes, names = c
        used = np.zeros(len(names), bool)
        used[codes] = True
        return all(n in ("+", "-") for n, u in zip(names, used) if u)
    return bool(df["Strand"].astype(str).isin(["+", "-"]).all())


def _group_frames(df, stranded):
    """{key: frame} of df.groupby(Chromosome[, Strand], observed=True) (pyranges/methods/init.py:140-150) for the
    categorical columns _set_dtypes leaves: one stable sort of the combined category codes, one gather per column, and
    the groups as row ranges of the gathered frame -- same keys in the same order, same rows with the same index."""
    ch = df["Chromosome"]
    st = df["Strand"] if stranded else None
    ok = isinstance(ch.dtype, pd.CategoricalDtype) and (st is None or isinstance(st.dtype, pd.CategoricalDtype))
    if ok:
        code = ch.cat.codes.to_numpy().astype(np.int64)
        ok = not len(code) or code.min() >= 0
        if ok and st is not None:
            sc = st.cat.codes.to_numpy().astype(np.int64)
            ok = sc.min() >= 0
            code = code * len(st.cat.categories) + sc
    if not ok:
        by = ["Chromosome", "Strand"] if stranded else "Chromosome"
        return {k: v for k, v in df.groupby(by, observed=True)}
    top = int(code.max()) + 1
    small = code.astype(np.int16) if top < (1 << 15) else code  # 16-bit keys: numpy's stable sort is a radix sort
    order = np.argsort(small, kind="stable")
    counts = np.bincount(code, minlength=top)
    big = df.take(order)
    chrom_cats = list(ch.cat.categories)
    strand_cats = list(st.cat.categories) if st is not None else None
    out, a = {}, 0
    for c in np.flatnonzero(counts).tolist():
        b = a + int(counts[c])
        key = (chrom_cats[c // len(strand_cats)], strand_cats[c % len(strand_cats)]) if st is not None else chrom_cats[c]
        out[key] = big.iloc[a:b]
        a = b
    return out


def _all_equal(col, value):
    """(col.astype(str) == value).all() without stringifying a categorical column row by row."""
    if isinstance(col.dtype, pd.CategoricalDtype):
        codes = col.cat.codes.to_numpy()
        if not len(codes):
            return True
        c = codes[0]
        return bool(c >= 0 and (codes == c).all() and str(col.cat.categories[c]) == value)
    return bool((col.astype(str) == value).all())


def concat(pyranges, strand=None):
    """pr.concat (pyranges/methods/concat.py:8-51): per-key concatenation of several PyRanges."""
    pyranges = [gr for gr in pyranges if not gr.empty]
    if not pyranges:
        return PyRanges()
    strand_info = [gr.stranded for gr in pyranges]
    if strand is None:
        strand = all(strand_info)
    per_key = {}
    if strand:
        assert all(strand_info), "Cannot do stranded concat, not all pyranges contain strand info."
        for gr in pyranges:
            for k, df in gr.dfs.items():
                per_key.setdefault(k, []).append(df)
    else:
        for gr in pyranges:
            for c in gr.chromosomes:
                per_key.setdefault(c, []).append(gr[c].df)
    dfs = {k: pd.concat(v, sort=False).reset_index(drop=True) for k, v in per_key.items()}
    if any(strand_info) and not all(strand_info):
        for k, v in dfs.items():
            if "Strand" in v:
                dfs[k] = v.assign(Strand=v.Strand.astype(object).where(v.Strand.notna(), "."))
    return PyRanges(dfs)


class PyRanges:
    def __init__(self, df=None, chromosomes=None, starts=None, ends=None, strands=None, int64=False, copy_df=True):
        if isinstance(df, PyRanges):
            raise Exception("Object is already a PyRange.")
        if df is None or df is False:
            if chromosomes is None:
                df = pd.DataFrame(columns="Chromosome Start End".split())
            else:
                d = {"Chromosome": chromosomes, "Start": starts, "End": ends}
                if strands is not None:
                    d["Strand"] = strands
                df = pd.DataFrame(d)
        if isinstance(df, pd.DataFrame):
            assert all(c in df for c in "Chromosome Start End".split()), \
                "The dataframe does not have all the columns Chromosome, Start and End."
            df = df.copy() if copy_df else df
            df = df.reset_index(drop=True)
            coded = {c: _codes_of(df[c]) for c in ("Chromosome", "Strand") if c in df}
            stranded = _check_strandedness(df, coded)
            if len(df):
                df = _set_dtypes(df, coded)
                # observed groups only: pandas >= 3 would otherwise create empty (chrom, '.') groups
                dfs = _group_frames(df, stranded)
            else:
                dfs = {}
        else:
            dfs = {k: (v.copy() if copy_df else v) for k, v in df.items() if not v.empty}
            ok = True
            for key, d in dfs.items():
                if isinstance(key, tuple):
                    ok = ok and ("Strand" in d) and key[1] in ("+", "-") and \
                        _all_equal(d["Chromosome"], str(key[0])) and _all_equal(d["Strand"], key[1])
                else:
                    # a frame that carries a Strand column belongs under a (chromosome, strand) key: the reference
                    # regroups such dicts (pyranges/methods/init.py:112-150), e.g. the result of unstranded.join(stranded)
                    ok = ok and "Strand" not in d and _all_equal(d["Chromosome"], str(key))
            if not ok:
                alldf = pd.concat(dfs.values()).reset_index(drop=True)
                stranded = _check_strandedness(alldf)
                dfs = _group_frames(alldf, stranded)
        self.__dict__["dfs"] = dfs

    @classmethod
    def _wrap(cls, dfs):
        """A PyRanges around a dict the dispatcher just produced: every frame is new, non-empty and belongs to its
        key, so the copy and the key validation of __init__ are skipped."""
        self = cls.__new__(cls)
        self.__dict__["dfs"] = dfs
        return self

    # ---- container helpers ---------------------------------------------------------------------------
    def __len__(self):
        return sum(len(d) for d in self.dfs.values())

    def __iter__(self):
        return iter(self.items())

    def keys(self):
        return natsorted(self.dfs.keys())

    def items(self):
        return natsorted([(k, df) for k, df in self.dfs.items()])

    def values(self):
        return [df for _, df in self.items()]

    @property
    def empty(self):
        return len(self) == 0

    @property
    def stranded(self):
        keys = self.keys()
        if not len(keys):
            return True
        return isinstance(keys[0], tuple)

    @property
    def chromosomes(self):
        if self.stranded:
            return natsorted(set(k[0] for k in self.keys()))
        return natsorted(set(self.keys()))

    @property
    def columns(self):
        v = self.values()
        return v[0].columns if v else pd.Index([])

    @property
    def df(self):
        if not len(self):
            return pd.DataFrame()
        return pd.concat(self.values()).reset_index(drop=True)

    as_df = lambda self: self.df  # noqa: E731

    def copy(self):
        return PyRanges({k: v.copy() for k, v in self.dfs.items()})

    def head(self, n=8):
        return PyRanges(self.df.head(n))

    def __getattr__(self, name):
        if name.startswith("__") and name.endswith("__") and name != "__ix__":
            raise AttributeError(name)
        if name in self.columns:
            return pd.concat([df[name] for df in self.values()])
        raise AttributeError("PyRanges object has no attribute", name)

    def __setattr__(self, column_name, column):
        """pyranges/methods/attr.py:7-58 (whole-column assignment in key order)."""
        if column_name == "columns":
            self.__dict__["dfs"] = {k: v.set_axis(column, axis=1) for k, v in self.dfs.items()}
            return
        if not len(self):
            return
        iterable = isinstance(column, (list, pd.Series, np.ndarray))
        if iterable and len(column) != len(self):
            raise Exception("DataFrame and column must be same length.")
        if isinstance(column, pd.Series):
            column = column.values
        exists = column_name in self.columns
        pos = list(self.values()[0].columns).index(column_name) if exists else self.values()[0].shape[1]
        a = 0
        dfs = {}
        for k, df in self.items():
            b = a + len(df)
            if exists:
                df = df.drop(column_name, axis=1)
            else:
                df = df.copy()
            df.insert(pos, column_name, column[a:b] if iterable else column)
            dfs[k] = df
            a = b
        if column_name in ("Chromosome", "Strand"):
            dfs = PyRanges(pd.concat(dfs.values())).dfs
        self.__dict__["dfs"] = dfs

    def __getitem__(self, val):
        """subset by chromosome / strand / (chromosome, strand) / boolean vector (pyranges/methods/getitem.py)."""
        if isinstance(val, str):
            if val in ("+", "-"):
                dfs = {k: v for k, v in self.dfs.items() if isinstance(k, tuple) and k[1] == val}
            else:
                dfs = {k: v for k, v in self.dfs.items() if (k[0] if isinstance(k, tuple) else k) == val}
        elif isinstance(val, tuple) and len(val) == 2 and all(isinstance(x, str) for x in val):
            dfs = {k: v for k, v in self.dfs.items() if k == val}
        elif isinstance(val, (pd.Series, np.ndarray)) and val.dtype == bool:
            assert len(val) == len(self), "Boolean indexer must be same length as pyrange!"
            val = val.values if isinstance(val, pd.Series) else val
            a, dfs = 0, {}
            for k, df in self.items():
                dfs[k] = df[val[a:a + len(df)]]
                a += len(df)
        else:
            raise NotImplementedError("pyranges_b200.PyRanges: subsetter %r is out of scope" % (val,))
        return PyRanges(dfs)

    def drop(self, drop=None, like=None):
        """pyranges/pyranges_main.py:1502 (columns by name or regex; the four core columns stay)."""
        import re

        core = {"Chromosome", "Start", "End", "Strand"}
        cols = list(self.columns)
        if drop is None and like is None:
            kill = [c for c in cols if c not in core]
        elif like is not None:
            rx = re.compile(like)
            kill = [c for c in cols if rx.search(c) and c not in core]
        else:
            kill = [drop] if isinstance(drop, str) else list(drop)
        if isinstance(drop, str) and drop == "Strand" or (not isinstance(drop, str) and drop and "Strand" in drop):
            return PyRanges(self.df.drop(kill, axis=1))
        return PyRanges({k: v.drop(kill, axis=1) for k, v in self.dfs.items()})

    def unstrand(self):
        """pyranges/pyranges_main.py:5882-5933"""
        if not self.stranded and "Strand" in self.columns:
            return self.drop("Strand")
        if not self.stranded:
            return self
        plus, minus = self["+"], self["-"]
        dfs = [*plus.values(), *minus.values()]
        return PyRanges(pd.concat(dfs).drop("Strand", axis=1).reset_index(drop=True))

    def extend(self, ext):
        """pyranges/multithreaded.py:411-443 (_extend) with an integer slack on both sides."""
        assert isinstance(ext, int), "pyranges_b200.PyRanges.extend: only integer extension is in scope"
        dfs = {}
        for k, df in self.dfs.items():
            df = df.copy()
            df["Start"] = np.maximum(df["Start"].values - ext, 0)
            df["End"] = df["End"].values + ext
            assert (df.Start < df.End).all(), "Some intervals are negative or zero length after applying extend!"
            dfs[k] = df
        return PyRanges(dfs)

    # ---- the hot path ------------------------------------------------------------------------------------
    def join(self, other, strandedness=None, how=None, report_overlap=False, slack=0, suffix="_b", nb_cpu=1,
             apply_strand_suffix=None, preserve_order=False):
        """pyranges/pyranges_main.py:2248-2451"""
        kwargs = {"strandedness": strandedness, "how": how, "report_overlap": report_overlap, "suffix": suffix,
                  "nb_cpu": nb_cpu, "apply_strand_suffix": apply_strand_suffix, "preserve_order": preserve_order}
        if slack:
            self = self.copy()
            self.Start__slack = self.Start
            self.End__slack = self.End
            self = self.extend(slack)
        if "suffix" in kwargs and isinstance(kwargs["suffix"], str):
            kwargs["suffixes"] = ("", kwargs["suffix"])
        kwargs = fill_kwargs(kwargs)
        how = kwargs.get("how")
        if how in ["left", "outer"]:
            kwargs["example_header_other"] = other.head(1).df
        if how in ["right", "outer"]:
            kwargs["example_header_self"] = self.head(1).df
        dfs = pyrange_apply("_write_both", self, other, **kwargs)
        gr = PyRanges._wrap(dfs)
        if slack and len(gr) > 0:
            gr.Start = gr.Start__slack
            gr.End = gr.End__slack
            gr = gr.drop(like="(Start|End).*__slack")
        if not self.stranded and other.stranded:
            if apply_strand_suffix is None:
                print("join: Strand data from other will be added as strand data to self.\nIf this is undesired use the "
                      "flag apply_strand_suffix=False.\nTo turn off the warning set apply_strand_suffix to True or False.",
                      file=sys.stderr)
            elif apply_strand_suffix:
                gr.columns = gr.columns.str.replace("Strand", "Strand" + kwargs["suffix"])
        return gr

    def overlap(self, other, strandedness=None, how="first", invert=False, nb_cpu=1):
        """pyranges/pyranges_main.py:3490-3628"""
        kwargs = {"strandedness": strandedness, "nb_cpu": nb_cpu}
        kwargs["sparse"] = {"self": False, "other": True}
        kwargs["how"] = how
        kwargs["invert"] = invert
        kwargs = fill_kwargs(kwargs)
        if len(self) == 0:
            return self
        if invert:
            self = self.copy()
            self.__ix__ = np.arange(len(self))
        dfs = pyrange_apply("_overlap", self, other, **kwargs)
        result = PyRanges._wrap(dfs)
        if invert:
            found_idxs = getattr(result, "__ix__", []) if len(result) else []
            result = self[~self.__ix__.isin(found_idxs)]
            result = result.drop("__ix__")
        return result

    def intersect(self, other, strandedness=None, how=None, invert=False, nb_cpu=1):
        """pyranges/pyranges_main.py:2092-2217 (pairs + elementwise clip)"""
        kwargs = fill_kwargs({"how": how, "strandedness": strandedness, "nb_cpu": nb_cpu})
        kwargs["sparse"] = {"self": False, "other": True}
        if len(self) == 0:
            return self
        if invert:
            self = self.copy()
            self.__ix__ = np.arange(len(self))
        result = PyRanges._wrap(pyrange_apply("_intersection", self, other, **kwargs))
        if invert:
            found_idxs = getattr(result, "__ix__", []) if len(result) else []
            result = self[~self.__ix__.isin(found_idxs)]
            result = result.drop("__ix__")
        return result

    def set_intersect(self, other, strandedness=None, how=None, new_pos=False, nb_cpu=1):
        """pyranges/pyranges_main.py:3865-3981: intersect of the two merged sets"""
        kwargs = fill_kwargs({"strandedness": strandedness, "how": how, "nb_cpu": nb_cpu, "new_pos": new_pos})
        strand = True if strandedness else False
        self_clusters = self.merge(strand=strand)
        other_clusters = other.merge(strand=strand)
        return PyRanges._wrap(pyrange_apply("_intersection", self_clusters, other_clusters, **kwargs))

    def set_union(self, other, strandedness=None, nb_cpu=1):
        """pyranges/pyranges_main.py:3983-4073: merge of the concatenation"""
        if self.empty and other.empty:
            return PyRanges()
        strand = True if strandedness else False
        if not strand:
            self = self.unstrand()
            other = other.unstrand()
        if strandedness == "opposite" and len(other):
            other = other.copy()
            other.Strand = other.Strand.astype(str).replace({"+": "-", "-": "+"})
        return concat([self, other], strand).merge(strand=strand)

    def subtract(self, other, strandedness=None, nb_cpu=1):
        """pyranges/pyranges_main.py:4802-4879.  (The reference first annotates self with a __num__ overlap count
        for NCLS.set_difference_helper; the batched kernel does not need it.)"""
        kwargs = fill_kwargs({"strandedness": strandedness})
        kwargs["sparse"] = {"self": False, "other": True}
        strand = True if strandedness else False
        other_clusters = other.merge(strand=strand)
        return PyRanges._wrap(pyrange_apply("_subtraction", self, other_clusters, **kwargs))

    def count_overlaps(self, other, strandedness=None, keep_nonoverlapping=True, overlap_col="NumberOverlaps"):
        """pyranges/pyranges_main.py:1293-1386"""
        kwargs = {"strandedness": strandedness, "keep_nonoverlapping": keep_nonoverlapping, "overlap_col": overlap_col}
        kwargs = fill_kwargs(kwargs)
        return PyRanges._wrap(pyrange_apply("_number_overlapping", self, other, **kwargs))

    def coverage(self, other, strandedness=None, keep_nonoverlapping=True, overlap_col="NumberOverlaps",
                 fraction_col="FractionOverlaps", nb_cpu=1):
        """pyranges/pyranges_main.py:1388-1500"""
        kwargs = {"strandedness": strandedness, "keep_nonoverlapping": keep_nonoverlapping, "overlap_col": overlap_col,
                  "fraction_col": fraction_col, "nb_cpu": nb_cpu}
        kwargs = fill_kwargs(kwargs)
        counts = self.count_overlaps(other, keep_nonoverlapping=True, overlap_col=overlap_col, strandedness=strandedness)
        strand = True if kwargs["strandedness"] else False
        other = other.merge(count=True, strand=strand)
        return PyRanges._wrap(pyrange_apply("_coverage", counts, other, **kwargs))

    def nearest(self, other, strandedness=None, overlap=True, how=None, suffix="_b", nb_cpu=1, apply_strand_suffix=None):
        """pyranges/pyranges_main.py:3192-3339"""
        kwargs = {"strandedness": strandedness, "how": how, "overlap": overlap, "nb_cpu": nb_cpu, "suffix": suffix,
                  "apply_strand_suffix": apply_strand_suffix}
        kwargs = fill_kwargs(kwargs)
        if kwargs.get("how") in "upstream downstream".split():
            assert other.stranded, "If doing upstream or downstream nearest, other pyranges must be stranded"
        dfs = pyrange_apply("_nearest", self, other, **kwargs)
        gr = PyRanges._wrap(dfs)
        if not self.stranded and other.stranded:
            if apply_strand_suffix is None:
                print("join: Strand data from other will be added as strand data to self.\nIf this is undesired use the "
                      "flag apply_strand_suffix=False.\nTo turn off the warning set apply_strand_suffix to True or False.",
                      file=sys.stderr)
            elif apply_strand_suffix:
                gr.columns = gr.columns.str.replace("Strand", "Strand" + kwargs["suffix"])
        return gr

    def k_nearest(self, other, k=1, ties=None, strandedness=None, overlap=True, how=None, suffix="_b", nb_cpu=1,
                  apply_strand_suffix=None):
        """pyranges/pyranges_main.py:2480-2834.  k: int or one int per row; ties in (None, "first", "last",
        "different").  Rows of one interval come out by increasing |Distance| (overlaps, then previous before next at
        equal distance); the reference leaves that order to an unstable sort when k varies over the rows."""
        kwargs = fill_kwargs({"strandedness": strandedness, "how": how, "overlap": overlap, "nb_cpu": nb_cpu, "k": k,
                              "ties": ties, "suffix": suffix})
        kwargs["stranded"] = self.stranded and other.stranded
        if ties not in (None, False, "first", "last", "different"):
            raise ValueError("ties must be None, 'first', 'last' or 'different'")
        ties = ties or None
        kwargs["ties"] = ties
        if not len(self):
            return PyRanges()
        me = self.copy()
        if isinstance(k, pd.Series):
            k = k.values
        me.__k__ = k if np.ndim(k) else int(k)   # how many to find, per row
        me.__IX__ = np.arange(len(me))            # row identity across the overlap / neighbour frames
        # (the validating constructor: frames of an unstranded self that picked up other's Strand column are regrouped by
        # (chromosome, strand) as in the reference -- the k nearest are then chosen per strand of the neighbour)
        near = PyRanges(pyrange_apply("_k_nearest", me, other, **kwargs))
        if overlap:
            hits = me.join(other, strandedness=strandedness, how={"first": "first", "last": "last"}.get(ties),
                           suffix=suffix, nb_cpu=nb_cpu, apply_strand_suffix=apply_strand_suffix)
            hits = PyRanges(hits.dfs)
            if len(hits):
                hits.Distance = 0
            result = concat([hits, near])
        else:
            result = near
        if not len(result):
            return PyRanges()
        out = {}
        for key, df in result.items():
            # by row, then distance; the stable sort keeps overlaps / previous / next in that order at equal distance
            ix, dist, kk = df["__IX__"].values, df["Distance"].values.astype(np.int64), df["__k__"].values.astype(np.int64)
            order = np.lexsort((dist, ix))
            ix, dist, kk = ix[order], dist[order], kk[order]
            first = np.r_[True, ix[1:] != ix[:-1]]
            start = np.maximum.accumulate(np.where(first, np.arange(len(ix)), 0))
            if ties == "different":   # every row of the k smallest distinct distances (get_different_ties)
                new_d = first | np.r_[True, dist[1:] != dist[:-1]]
                rank = np.cumsum(new_d)
                keep = rank - rank[start] < kk
            else:                     # the k first rows (get_all_ties + head(k); head(k) for first / last)
                keep = np.arange(len(ix)) - start < kk
            df = df.take(order[keep]).drop(["__IX__", "__k__"], axis=1)
            if not len(df):
                continue
            # neighbours in front of the interval (behind it on the - strand) get a negative distance (:2802-2813)
            strand = df.Strand.iloc[0] if "Strand" in df else "+"
            neg = (df["End" + suffix].values < df.Start.values)
            if strand != "+":
                neg = ~neg
            d = df["Distance"].values.astype(np.int64)
            df = df.assign(Distance=np.where(neg, -d, d)).reset_index(drop=True)
            out[key] = df
        gr = PyRanges._wrap(out)
        if not self.stranded and other.stranded:
            if apply_strand_suffix is None:
                print("join: Strand data from other will be added as strand data to self.\nIf this is undesired use the "
                      "flag apply_strand_suffix=False.\nTo turn off the warning set apply_strand_suffix to True or False.",
                      file=sys.stderr)
            elif apply_strand_suffix:
                gr.columns = gr.columns.str.replace("Strand", "Strand" + suffix)
        return gr

    def cluster(self, strand=None, by=None, slack=0, count=False, nb_cpu=1):
        """pyranges/pyranges_main.py:1067-1230"""
        if strand is None:
            strand = self.stranded
        kwargs = fill_kwargs({"strand": strand, "slack": slack, "count": count, "by": by})
        _stranded = self.stranded
        if not strand and _stranded:
            self.Strand2 = self.Strand
            self = self.unstrand()
        df = pyrange_apply_single("_cluster_by" if by else "_cluster", self, **kwargs)
        gr = PyRanges._wrap(df)
        # ids are per key; make them globally unique in natsorted key order (pyranges_main.py:1210-1223)
        new_dfs = {}
        max_id = 0
        for i, (k, v) in enumerate(gr.items()):
            if i:
                v = v.copy(deep=False)
                v["Cluster"] = v["Cluster"].values + max_id
            max_id = v.Cluster.max()
            new_dfs[k] = v
        if not strand and _stranded:
            new_dfs = {k: d.rename(columns={"Strand2": "Strand"}) for k, d in new_dfs.items()}
        return PyRanges._wrap(new_dfs)

    def merge(self, strand=None, count=False, count_col="Count", by=None, slack=0):
        """pyranges/pyranges_main.py:3005-3146"""
        if strand is None:
            strand = self.stranded
        kwargs = {"strand": strand, "count": count, "by": by, "count_col": count_col, "slack": slack,
                  "sparse": {"self": not by}}
        return PyRanges._wrap(pyrange_apply_single("_merge_by" if by else "_merge", self, **kwargs))

    def split(self, strand=None, between=False, nb_cpu=1):
        """pyranges/pyranges_main.py:4351-4475: the intervals cut at every Start / End of the key (SURVEY.md 8(f)
        rank 3; the feature table of the multi-sample count matrix)."""
        if strand is None:
            strand = self.stranded
        kwargs = fill_kwargs({"strand": strand})
        split = PyRanges(pyrange_apply_single("_split", self, **kwargs))
        if not between:
            split = split.overlap(self, strandedness="same" if strand else False)
        return split

    def to_rle(self, value_col=None, strand=None, rpm=False, nb_cpu=1):
        """pyranges/pyranges_main.py:5772-5880, pyranges/methods/to_rle.py:4-24"""
        from .rle import PyRles

        if strand is None:
            strand = self.stranded
        kwargs = {"strand": strand, "value_col": value_col, "sparse": {"self": False}, "nb_cpu": nb_cpu}
        result = pyrange_apply_single("coverage", self, **kwargs)
        if rpm:
            multiplier = 1e6 / len(self)
            result = {k: v * multiplier for k, v in result.items()}
        return PyRles(result)

    def __getstate__(self):
        return self.dfs

    def __setstate__(self, d):
        self.__dict__["dfs"] = d

    def __repr__(self):
        return "PyRanges(%d rows, %d keys, %s)" % (len(self), len(self.dfs), "stranded" if self.stranded else "unstranded")
