"""coverage_rle / cluster / split on n reads (for ncu launch lists): python tools/prof_rle.py [n]"""
import sys

import torch

sys.path.insert(0, ".")
from pyranges_b200 import engine as E
from pyranges_b200 import synthetic as S

n = int(sys.argv[1]) if len(sys.argv) > 1 else 50_000_000
iv = S.reads(n, seed=4)
r = E.DeviceIntervals.from_numpy(iv.starts, iv.ends, iv.seg_off)
r.bounds
for name, fn in (("coverage_rle", lambda: E.coverage_rle(r)), ("cluster", lambda: E.cluster(r)), ("split", lambda: E.split(r))):
    for _ in range(2):
        fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(3):
        fn()
    b.record()
    torch.cuda.synchronize()
    print("%s %d rows: %.3f ms" % (name, n, a.elapsed_time(b) / 3))
