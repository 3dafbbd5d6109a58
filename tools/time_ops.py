"""Device times of the other hot-path operations at the BASELINE config sizes (CUDA events, steady state).
Usage: python tools/time_ops.py [scale]   scale 1.0 = C4 (50 M intervals) / C5 (200 M reads)"""
import sys
import time

import numpy as np
import torch

sys.path.insert(0, ".")
from pyranges_b200 import engine as E
from pyranges_b200 import synthetic as S

scale = float(sys.argv[1]) if len(sys.argv) > 1 else 1.0


def timeit(name, fn, reps=5, alg_bytes=None, units=None):
    for _ in range(2):
        fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(reps):
        fn()
    b.record()
    torch.cuda.synchronize()
    ms = a.elapsed_time(b) / reps
    extra = ""
    if alg_bytes:
        extra += "  %.0f GB/s algorithmic" % (alg_bytes / ms / 1e6)
    if units:
        extra += "  %.2f G rows/s" % (units / ms / 1e6)
    print("%-44s %9.3f ms%s" % (name, ms, extra), flush=True)


# C3: join / count_overlaps of 100 M reads against 1 M annotations (strand ignored)
n3 = int(100_000_000 * scale)
t0 = time.time()
r3, a3 = S.reads(n3, seed=1), S.annotations(int(1_000_000 * scale), seed=2)
print("generated C3 in %.1f s" % (time.time() - t0), flush=True)
q3 = E.DeviceIntervals.from_numpy(r3.starts, r3.ends, r3.seg_off)
s3 = E.DeviceIntervals.from_numpy(a3.starts, a3.ends, a3.seg_off)
s3.bounds
g3 = a3.chrom_ranges()
sg3 = (np.arange(r3.n_keys) // 2).astype(np.int32)
qq3 = E.Queries(q3, sg3)
ix3 = E.SubjectIndex(s3, g3)
_, pa, _ = ix3.pairs(qq3, E.IV_MODE_ALL, want_counts=False)
p3 = pa.numel()
del pa
timeit("C3 index build (%d subjects)" % a3.n, lambda: E.SubjectIndex(s3, g3), alg_bytes=24 * a3.n)
timeit("C3 iv_count_overlaps counts only", lambda: ix3.count(qq3, E.IV_MODE_ALL, for_emit=False), alg_bytes=24 * n3, units=n3)
timeit("C3 count (hit lists, no counts)", lambda: ix3.count(qq3, E.IV_MODE_ALL, want_counts=False), alg_bytes=16 * n3, units=n3)
timeit("C3 join: count + emit (%d pairs)" % p3, lambda: ix3.pairs(qq3, E.IV_MODE_ALL, want_counts=False, capacity=p3 + 4096),
       alg_bytes=16 * n3 + 16 * p3, units=p3)
del q3, qq3, ix3, s3, r3, a3
torch.cuda.empty_cache()

n4 = int(50_000_000 * scale)
t0 = time.time()
iv4 = S.reads(n4, seed=3, min_len=50, max_len=500)
ann = S.annotations(int(1_000_000 * scale), seed=2)
print("generated C4 in %.1f s" % (time.time() - t0), flush=True)
q = E.DeviceIntervals.from_numpy(iv4.starts, iv4.ends, iv4.seg_off)
s = E.DeviceIntervals.from_numpy(ann.starts, ann.ends, ann.seg_off)
s.bounds
q.bounds
# nearest(strandedness="same", how="upstream"): partner = same key; '+' keys look at previous, '-' at next
sg = np.arange(iv4.n_keys, dtype=np.int32)
modes = np.asarray([E.IV_NEAREST_PREVIOUS if k % 2 == 0 else E.IV_NEAREST_NEXT for k in range(iv4.n_keys)], np.int32)
qq = E.Queries(q, sg, seg_mode=modes)
timeit("nearest index build (1M subjects, ends perm)", lambda: E.SubjectIndex(s, None, ends_perm=True), alg_bytes=24 * ann.n)
ix = E.SubjectIndex(s, None, ends_perm=True)
timeit("iv_nearest %d x %d upstream/same" % (n4, ann.n), lambda: ix.nearest(qq, overlap=True), alg_bytes=40 * n4, units=n4)
timeit("iv_count_overlaps %d x %d" % (n4, ann.n), lambda: ix.count(qq, E.IV_MODE_ALL, for_emit=False), alg_bytes=24 * n4, units=n4)
timeit("cluster %d unsorted (48 keys)" % n4, lambda: E.cluster(q), alg_bytes=40 * n4, units=n4)
timeit("merge %d unsorted (48 keys)" % n4, lambda: E.merge(q), alg_bytes=16 * n4, units=n4)
o = np.lexsort([iv4.ends, iv4.starts, np.repeat(np.arange(iv4.n_keys), np.diff(iv4.seg_off))])
qs = E.DeviceIntervals.from_numpy(iv4.starts[o], iv4.ends[o], iv4.seg_off)
qs.bounds
timeit("split %d (48 keys)" % n4, lambda: E.split(q), alg_bytes=16 * n4 + 16 * 2 * n4, units=n4)
timeit("cluster %d sorted" % n4, lambda: E.cluster(qs), alg_bytes=40 * n4, units=n4)
timeit("merge %d sorted" % n4, lambda: E.merge(qs), alg_bytes=16 * n4, units=n4)
del q, qs, qq, ix
torch.cuda.empty_cache()
n5 = int(200_000_000 * scale)
t0 = time.time()
iv5 = S.reads(n5, seed=4)
print("generated C5 in %.1f s" % (time.time() - t0), flush=True)
r = E.DeviceIntervals.from_numpy(iv5.starts, iv5.ends, iv5.seg_off)
r.bounds
off, runs, vals = E.coverage_rle(r)
print("runs:", len(runs))
timeit("coverage_rle %d reads (48 keys)" % n5, lambda: E.coverage_rle(r), reps=3, alg_bytes=16 * n5 + 16 * len(runs), units=n5)
print("max memory allocated: %.1f GB" % (torch.cuda.max_memory_allocated() / 1e9))
