"""shim of ``natsort.natsorted`` (test infrastructure): chr1 < chr2 < chr10 < chrX; tuples element-wise."""
import re

_num = re.compile(r"(\d+)")


def _key(x):
    if isinstance(x, str):
        parts = _num.split(x)
        return tuple(int(p) if i % 2 else p for i, p in enumerate(parts))
    if isinstance(x, (tuple, list)):
        return tuple(_key(y) for y in x)
    if isinstance(x, (int, float)):
        return ("", x)
    return ()  # DataFrames etc. (second element of dict items): never decisive, keys are unique


def natsorted(seq, key=None, reverse=False, alg=None):
    if key is None:
        return sorted(seq, key=_key, reverse=reverse)
    return sorted(seq, key=lambda v: _key(key(v)), reverse=reverse)


def natsort_keygen(key=None, alg=None):
    if key is None:
        return _key
    return lambda v: _key(key(v))


class ns:  # placeholder for natsort.ns flags
    pass
