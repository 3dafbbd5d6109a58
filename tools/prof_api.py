"""cProfile of one public-API call (join by default) at a moderate size: where the host time goes."""
import cProfile
import pstats
import sys

import numpy as np
import pandas as pd

sys.path.insert(0, ".")
from pyranges_b200 import synthetic as S
from pyranges_b200.pyranges_main import PyRanges


def frame(k):
    key = np.repeat(np.arange(k.n_keys), np.diff(k.seg_off))
    return pd.DataFrame({"Chromosome": np.asarray([c for c, _ in S.HG38], dtype=object)[key // 2], "Start": k.starts,
                         "End": k.ends, "Strand": np.asarray(["+", "-"], dtype=object)[key % 2],
                         "Score": np.arange(k.n) % 1000})


n = int(sys.argv[1]) if len(sys.argv) > 1 else 2_000_000
op = sys.argv[2] if len(sys.argv) > 2 else "join"
gr, an = PyRanges(frame(S.reads(n, seed=1))), PyRanges(frame(S.annotations(n // 50, seed=2)))
fn = {"join": lambda: gr.join(an), "count_overlaps": lambda: gr.count_overlaps(an), "overlap": lambda: gr.overlap(an),
      "nearest": lambda: gr.nearest(an), "cluster": lambda: gr.cluster(), "merge": lambda: gr.merge(),
      "to_rle": lambda: gr.to_rle(), "subtract": lambda: gr.subtract(an)}[op]
fn()
pr = cProfile.Profile()
pr.enable()
fn()
pr.disable()
pstats.Stats(pr).sort_stats("cumulative").print_stats(35)
