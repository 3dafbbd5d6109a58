"""read_bed: the parse on the device (pyranges_b200.readers) against pandas.read_csv + PyRanges on a generated BED6 file.
Usage: python tools/time_ingest.py [rows]"""
import os
import sys
import tempfile
import time

import numpy as np
import pandas as pd
import torch

sys.path.insert(0, ".")
from pyranges_b200 import multithreaded as M
from pyranges_b200.pyranges_main import PyRanges
from pyranges_b200.readers import read_bed

n = int(sys.argv[1]) if len(sys.argv) > 1 else 5_000_000
rng = np.random.default_rng(0)
chrom = np.asarray(["chr%d" % i for i in range(1, 23)] + ["chrX", "chrY"])[rng.integers(0, 24, n)]
st = rng.integers(0, 200_000_000, n)
df = pd.DataFrame({"c": chrom, "s": st, "e": st + rng.integers(50, 500, n), "n": ["r%d" % (i % 1000) for i in range(n)],
                   "sc": rng.integers(0, 1000, n), "st": np.asarray(["+", "-"])[rng.integers(0, 2, n)]})
with tempfile.TemporaryDirectory() as d:
    path = os.path.join(d, "big.bed")
    df.to_csv(path, sep="\t", header=False, index=False)
    size = os.path.getsize(path)
    read_bed(path)  # untimed first pass: library load, lazily loaded kernels, page cache of the file
    pd.read_csv(path, dtype={0: "category", 5: "category"}, header=None, sep="\t", nrows=1000)
    M.clear_device_cache()
    torch.cuda.synchronize()
    t = time.perf_counter()
    gr = read_bed(path)
    torch.cuda.synchronize()
    t_dev = time.perf_counter() - t
    t = time.perf_counter()
    fr = read_bed(path, as_df=True)
    torch.cuda.synchronize()
    t_df = time.perf_counter() - t
    t = time.perf_counter()
    ref = pd.read_csv(path, dtype={0: "category", 5: "category"}, header=None, sep="\t")
    t_csv = time.perf_counter() - t
    ref.columns = "Chromosome Start End Name Score Strand".split()
    t = time.perf_counter()
    gr2 = PyRanges(ref)
    t_ctor = time.perf_counter() - t
    M.clear_device_cache()
    t = time.perf_counter()
    gr2.cluster()
    torch.cuda.synchronize()
    t_first_ref = time.perf_counter() - t
    t = time.perf_counter()
    gr.cluster()
    torch.cuda.synchronize()
    t_first_dev = time.perf_counter() - t
print("%d rows, %.0f MB BED6" % (n, size / 1e6))
print("read_bed (device parse) -> PyRanges        %.2f s   (frame only: %.2f s)" % (t_dev, t_df))
print("pandas.read_csv                            %.2f s   + PyRanges(df) %.2f s" % (t_csv, t_ctor))
print("first cluster() on the pandas-built object %.3f s, on the read_bed object %.3f s (columns already on the device)"
      % (t_first_ref, t_first_dev))
