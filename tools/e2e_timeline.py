"""Where the end-to-end join (host arrays in, host pair arrays out: bench.py's `e2e`) spends its time: the PCIe rates of
this box alone and in both directions at once, and the step with 4 / 8 / 16 / 32 row chunks.
Usage: python tools/e2e_timeline.py [reads] [annotations]"""
import sys
import time

import numpy as np
import torch

sys.path.insert(0, ".")
from pyranges_b200 import hostapi
from pyranges_b200 import synthetic as S

nq = int(sys.argv[1]) if len(sys.argv) > 1 else 100_000_000
ns = int(sys.argv[2]) if len(sys.argv) > 2 else 1_000_000
dev = torch.device("cuda", 0)
reads, annot = S.reads_by_chrom(nq, seed=1), S.annotations_by_chrom(ns, seed=2)
seg_group = (np.arange(reads.n_keys) // 2).astype(np.int32)
groups = [(i, i + 2) for i in range(0, annot.n_keys, 2)]
hq_s, hq_e = torch.from_numpy(reads.starts).pin_memory(), torch.from_numpy(reads.ends).pin_memory()
hs_s, hs_e = torch.from_numpy(annot.starts).pin_memory(), torch.from_numpy(annot.ends).pin_memory()


def wall(f, reps=3):
    f()
    torch.cuda.synchronize()
    t = time.perf_counter()
    for _ in range(reps):
        f()
    torch.cuda.synchronize()
    return (time.perf_counter() - t) / reps


# PCIe alone and both ways at once (1.6 GB up, 2.8 GB down: the sizes of the step)
up_h = torch.empty(200_000_000, dtype=torch.int64).pin_memory()
up_d = torch.empty_like(up_h, device=dev)
dn_d = torch.empty(355_000_000, dtype=torch.int64, device=dev)
dn_h = torch.empty(355_000_000, dtype=torch.int64).pin_memory()
s1, s2 = torch.cuda.Stream(dev), torch.cuda.Stream(dev)
t_up = wall(lambda: up_d.copy_(up_h, non_blocking=True))
t_dn = wall(lambda: dn_h.copy_(dn_d, non_blocking=True))


def both():
    with torch.cuda.stream(s1):
        up_d.copy_(up_h, non_blocking=True)
    with torch.cuda.stream(s2):
        dn_h.copy_(dn_d, non_blocking=True)


t_both = wall(both)
print("H2D 1.6 GB alone  %.1f ms (%.1f GB/s)" % (1e3 * t_up, 1.6 / t_up))
print("D2H 2.84 GB alone %.1f ms (%.1f GB/s)" % (1e3 * t_dn, 2.84 / t_dn))
print("both at once      %.1f ms" % (1e3 * t_both))
del up_h, up_d, dn_d, dn_h

nb0 = hostapi.NCLSBatch(hs_s, hs_e, annot.seg_off, groups, device=dev)
x, _ = nb0.all_overlaps_both(hq_s, hq_e, reads.seg_off, seg_group,
                             out=(torch.empty(200_000_000, dtype=torch.int64).pin_memory(),
                                  torch.empty(200_000_000, dtype=torch.int64).pin_memory()))
pairs = x.numel()
cap = int(pairs * 1.05) + 4096
out = (torch.empty(cap, dtype=torch.int64).pin_memory(), torch.empty(cap, dtype=torch.int64).pin_memory())
for chunks in (4, 8, 16, 32):
    def step():
        nb = hostapi.NCLSBatch(hs_s, hs_e, annot.seg_off, groups, device=dev)
        nb.all_overlaps_both(hq_s, hq_e, reads.seg_off, seg_group, out=out, chunks=chunks)
    t = wall(step, 5)
    print("e2e step, %2d chunks: %.1f ms (%.2f G pairs/s)" % (chunks, 1e3 * t, pairs / t / 1e9))
