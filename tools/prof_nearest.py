"""iv_nearest on a C4-shaped input (for ncu captures): python tools/prof_nearest.py [queries] [subjects]"""
import sys

import numpy as np
import torch

sys.path.insert(0, ".")
from pyranges_b200 import engine as E
from pyranges_b200 import synthetic as S

nq = int(sys.argv[1]) if len(sys.argv) > 1 else 20_000_000
ns = int(sys.argv[2]) if len(sys.argv) > 2 else 1_000_000
iv, ann = S.reads(nq, seed=3, min_len=50, max_len=500), S.annotations(ns, seed=2)
q = E.DeviceIntervals.from_numpy(iv.starts, iv.ends, iv.seg_off)
s = E.DeviceIntervals.from_numpy(ann.starts, ann.ends, ann.seg_off)
sg = np.arange(iv.n_keys, dtype=np.int32)
modes = np.asarray([E.IV_NEAREST_PREVIOUS if k % 2 == 0 else E.IV_NEAREST_NEXT for k in range(iv.n_keys)], np.int32)
ix = E.SubjectIndex(s, None, ends_perm=True)
for name, qq in (("upstream/same", E.Queries(q, sg, seg_mode=modes)), ("both", E.Queries(q, sg))):
    for _ in range(2):
        ix.nearest(qq, overlap=True)
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(5):
        ix.nearest(qq, overlap=True)
    b.record()
    torch.cuda.synchronize()
    ms = a.elapsed_time(b) / 5
    print("iv_nearest %s %d x %d: %.3f ms = %.1f G rows/s" % (name, nq, ns, ms, nq / ms / 1e6))
