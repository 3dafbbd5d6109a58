"""Minimal run-length containers returned by PyRanges.to_rle: the reference returns pyrle.Rle / pyrle.PyRles
(pyranges/methods/to_rle.py:18-24); only the fields the hot path produces are mirrored (runs int64,
values float64, scalar multiplication for rpm)."""
import numpy as np


class Rle:
    def __init__(self, runs, values):
        self.runs = np.asarray(runs, dtype=np.int64)
        self.values = np.asarray(values, dtype=np.float64)

    def __mul__(self, other):
        return Rle(self.runs, self.values * other)

    __rmul__ = __mul__

    def __len__(self):
        return len(self.runs)

    def __eq__(self, other):
        return np.array_equal(self.runs, other.runs) and np.array_equal(self.values, other.values)

    def __repr__(self):
        return "Rle(runs=%r, values=%r)" % (self.runs.tolist()[:8], self.values.tolist()[:8])


class PyRles:
    def __init__(self, rles):
        self.rles = dict(rles)

    def __getitem__(self, key):
        return self.rles[key]

    def keys(self):
        return list(self.rles.keys())

    def items(self):
        return list(self.rles.items())

    @property
    def stranded(self):
        k = self.keys()
        return bool(k) and isinstance(k[0], tuple)

    def __len__(self):
        return len(self.rles)
