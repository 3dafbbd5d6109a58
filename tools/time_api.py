"""Wall-clock of the public PyRanges API (DataFrames in, DataFrames out) at a moderate size."""
import sys
import time

import numpy as np
import pandas as pd

sys.path.insert(0, ".")
from pyranges_b200 import synthetic as S
from pyranges_b200.pyranges_main import PyRanges


def frame(k, names):
    key = np.repeat(np.arange(k.n_keys), np.diff(k.seg_off))
    df = pd.DataFrame({"Chromosome": np.asarray([c for c, _ in S.HG38], dtype=object)[key // 2], "Start": k.starts,
                       "End": k.ends, "Strand": np.asarray(["+", "-"], dtype=object)[key % 2]})
    if names:
        df["Score"] = np.arange(len(df)) % 1000
    return df


n = int(sys.argv[1]) if len(sys.argv) > 1 else 2_000_000
reads, annot = S.reads(n, seed=1), S.annotations(n // 50, seed=2)
t = time.time(); gr = PyRanges(frame(reads, True)); an = PyRanges(frame(annot, True)); print("construct %.2f s" % (time.time() - t))
for name, fn in [("count_overlaps", lambda: gr.count_overlaps(an)), ("join", lambda: gr.join(an)),
                 ("overlap", lambda: gr.overlap(an)), ("nearest", lambda: gr.nearest(an)),
                 ("cluster", lambda: gr.cluster()), ("merge", lambda: gr.merge()), ("to_rle", lambda: gr.to_rle()),
                 ("coverage", lambda: gr.coverage(an))]:
    fn()
    dt = 1e9
    for _ in range(3):  # best of three: a call now and then pays for a fresh device allocation
        t = time.time(); r = fn(); dt = min(dt, time.time() - t)
    print("%-16s %.3f s   (%d rows out)" % (name, dt, len(r) if not hasattr(r, "rles") else sum(len(v) for v in r.rles.values())), flush=True)
