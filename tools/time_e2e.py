"""Where the end-to-end (host arrays in, host arrays out) join step spends its time: raw PCIe copies next to the
chunked three-stream pipeline of hostapi.NCLSBatch.all_overlaps_both at several chunk counts."""
import sys
import time

import numpy as np
import torch

sys.path.insert(0, ".")
from pyranges_b200 import engine as E
from pyranges_b200 import hostapi
from pyranges_b200 import synthetic as S

nr = int(sys.argv[1]) if len(sys.argv) > 1 else 10_000_000
reads, annot = S.reads(nr), S.annotations(200_000)
sg = (np.arange(reads.n_keys) // 2).astype(np.int32)
dev = torch.device("cuda", 0)
hq_s, hq_e = torch.from_numpy(reads.starts).pin_memory(), torch.from_numpy(reads.ends).pin_memory()
hs_s, hs_e = torch.from_numpy(annot.starts).pin_memory(), torch.from_numpy(annot.ends).pin_memory()
s_iv = E.DeviceIntervals.from_numpy(annot.starts, annot.ends, annot.seg_off)
bounds = s_iv.bounds
groups = annot.chrom_ranges()


def wall(fn, reps=10):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    t = time.perf_counter()
    for _ in range(reps):
        fn()
    torch.cuda.synchronize()
    return (time.perf_counter() - t) / reps * 1e3


d = torch.empty(2 * nr, dtype=torch.int64, device=dev)
ms = wall(lambda: (d[:nr].copy_(hq_s, non_blocking=True), d[nr:].copy_(hq_e, non_blocking=True)))
print("H2D %d MB pinned: %.2f ms = %.1f GB/s" % (16 * nr / 1e6, ms, 16 * nr / ms / 1e6))
nb = hostapi.NCLSBatch(hs_s, hs_e, annot.seg_off, groups, device=dev, bounds=bounds)
a, b = nb.all_overlaps_both(reads.starts, reads.ends, reads.seg_off, sg)
p = len(a)
cap = int(p * 1.1) + 1024
out = (torch.empty(cap, dtype=torch.int64).pin_memory(), torch.empty(cap, dtype=torch.int64).pin_memory())
dp = torch.empty(2 * p, dtype=torch.int64, device=dev)
ms = wall(lambda: (out[0][:p].copy_(dp[:p], non_blocking=True), out[1][:p].copy_(dp[p:], non_blocking=True)))
print("D2H %d MB pinned: %.2f ms = %.1f GB/s" % (16 * p / 1e6, ms, 16 * p / ms / 1e6))


def step(chunks):
    nb = hostapi.NCLSBatch(hs_s, hs_e, annot.seg_off, groups, device=dev, bounds=bounds)
    return nb.all_overlaps_both(hq_s, hq_e, reads.seg_off, sg, out=out, chunks=chunks)


for c in (1, 2, 4, 8, 12, 16, 24, 32):
    print("e2e chunks=%2d: %.2f ms" % (c, wall(lambda: step(c))))
