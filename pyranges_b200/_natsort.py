"""Natural key order of PyRanges dicts (the reference uses natsort.natsorted: pyranges/multithreaded.py:213-214,
pyranges/pyranges_main.py:2246,2478): chr1 < chr2 < chr10 < chrX; tuple keys element-wise, '+' < '-'."""
import re

_digits = re.compile(r"(\d+)")


def natkey(x):
    if isinstance(x, str):
        parts = _digits.split(x)
        return tuple((1, int(p)) if i % 2 else (0, p) for i, p in enumerate(parts) if i % 2 or p != "" or i == 0)
    if isinstance(x, (tuple, list)):
        return tuple(natkey(y) for y in x)
    if isinstance(x, (int, float)):
        return ((0, ""), (1, x))
    return ()


def natsorted(seq, key=None):
    """sorts keys, or (key, value) items by their key only (values are never compared)."""
    if key is not None:
        return sorted(seq, key=lambda v: natkey(key(v)))
    seq = list(seq)
    if seq and isinstance(seq[0], tuple) and len(seq[0]) == 2 and not isinstance(seq[0][1], (str, int, float)):
        return sorted(seq, key=lambda kv: natkey(kv[0]))  # dict items
    return sorted(seq, key=natkey)
