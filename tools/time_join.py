"""Device times of the join / count paths at a BASELINE config (CUDA events, steady state).
Usage: python tools/time_join.py [c2|c3] [reps]"""
import sys
import time

import numpy as np
import torch

sys.path.insert(0, ".")
from pyranges_b200 import engine as E
from pyranges_b200 import synthetic as S

wl = sys.argv[1] if len(sys.argv) > 1 else "c3"
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 10
nr, na = {"c2": (10_000_000, 200_000), "c3": (100_000_000, 1_000_000)}[wl]
t0 = time.time()
reads, annot = S.reads_by_chrom(nr, seed=1), S.annotations_by_chrom(na, seed=2)
print("generated %s in %.1f s" % (wl, time.time() - t0), flush=True)
q_iv = E.DeviceIntervals.from_numpy(reads.starts, reads.ends, reads.seg_off)
s_iv = E.DeviceIntervals.from_numpy(annot.starts, annot.ends, annot.seg_off)
s_iv.bounds
groups = annot.chrom_ranges()
sg = (np.arange(reads.n_keys) // 2).astype(np.int32)
q = E.Queries(q_iv, sg)


def timeit(name, fn, alg_bytes=None):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(reps):
        fn()
    b.record()
    torch.cuda.synchronize()
    ms = a.elapsed_time(b) / reps
    print("%-52s %9.4f ms%s" % (name, ms, "  %.0f GB/s" % (alg_bytes / ms / 1e6) if alg_bytes else ""), flush=True)
    return ms


reach = E._max_len(q)
lx = E.ListIndex(s_iv, groups, reach=reach)
print("shift", lx.c.shift, "reach", lx.c.reach, "bins", lx.c.n_bins, "cand_cap", lx.c.cand_cap)
a, b = lx.pairs(q)
p = a.numel()
print("pairs", p)
cap = p + 4096
oq = torch.empty(cap, dtype=torch.int64, device="cuda")
os_ = torch.empty(cap, dtype=torch.int64, device="cuda")
del a, b
timeit("lists build", lambda: E.ListIndex(s_iv, groups, reach=reach), 24 * annot.n)
timeit("fused join kernel (+ memset)", lambda: lx.join_pairs(q, oq, os_, cap), 16 * reads.n + 16 * p)
timeit("fused join, out_q only", lambda: lx.join_pairs(q, oq, None, cap), 16 * reads.n + 8 * p)
timeit("list count ALL", lambda: lx.count(q, E.IV_MODE_ALL), 24 * reads.n)
timeit("list count ANY", lambda: lx.count(q, E.IV_MODE_ANY), 24 * reads.n)
ix = E.SubjectIndex(s_iv, groups)
timeit("rank index build", lambda: E.SubjectIndex(s_iv, groups), 24 * annot.n)
timeit("rank count ALL (counts only)", lambda: ix.count(q, E.IV_MODE_ALL, for_emit=False), 24 * reads.n)
print("max memory allocated: %.1f GB" % (torch.cuda.max_memory_allocated() / 1e9))
