"""``pyrle.methods.coverage`` per-key entry as called through pyrange_apply_single (to_rle.py:18)."""
import numpy as np

from oracle.oracle import coverage_rle
from pyrle import Rle


def coverage(df, **kwargs):
    value_col = kwargs.get("value_col", None)
    w = None
    if value_col:
        w = df[value_col].astype(np.float64).values
    runs, vals = coverage_rle(df.Start.values, df.End.values, w)
    return Rle(runs, vals)
