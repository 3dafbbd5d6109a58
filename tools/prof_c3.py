"""The join at the north-star size (100 M reads x 1 M annotations), for ncu captures: python tools/prof_c3.py [scale]"""
import sys

import numpy as np
import torch

sys.path.insert(0, ".")
from pyranges_b200 import engine as E
from pyranges_b200 import synthetic as S

scale = float(sys.argv[1]) if len(sys.argv) > 1 else 1.0
r3, a3 = S.reads(int(100_000_000 * scale), seed=1), S.annotations(int(1_000_000 * scale), seed=2)
q3 = E.DeviceIntervals.from_numpy(r3.starts, r3.ends, r3.seg_off)
s3 = E.DeviceIntervals.from_numpy(a3.starts, a3.ends, a3.seg_off)
s3.bounds
qq = E.Queries(q3, (np.arange(r3.n_keys) // 2).astype(np.int32))
ix = E.SubjectIndex(s3, a3.chrom_ranges())
_, a, _ = ix.pairs(qq, E.IV_MODE_ALL, want_counts=False)
p = a.numel()
del a
for _ in range(3):
    ix.pairs(qq, E.IV_MODE_ALL, want_counts=False, capacity=p + 4096)
torch.cuda.synchronize()
x, y = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
x.record()
for _ in range(5):
    ix.pairs(qq, E.IV_MODE_ALL, want_counts=False, capacity=p + 4096)
y.record()
torch.cuda.synchronize()
print("C3 x %.2f: count + emit of %d pairs: %.3f ms (shift %d, %d bins)" % (scale, p, x.elapsed_time(y) / 5, ix.c.shift, ix.c.n_bins))
