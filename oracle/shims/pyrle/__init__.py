"""shim of ``pyrle`` (test infrastructure): just enough of Rle / PyRles for methods/to_rle.py:4-24."""
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

    def __repr__(self):
        return "Rle(runs=%r, values=%r)" % (self.runs.tolist(), self.values.tolist())


class PyRles:
    def __init__(self, rles):
        self.rles = dict(rles)

    def __getitem__(self, k):
        return self.rles[k]

    def keys(self):
        return list(self.rles.keys())

    def items(self):
        return list(self.rles.items())

    def __len__(self):
        return len(self.rles)
