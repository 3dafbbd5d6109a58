"""Steady-state device time of the individual C-ABI calls on the bench workload (CUDA events, back-to-back
launches so that host overhead is hidden).  Usage: python tools/time_kernels.py [reads] [annot]"""
import ctypes as C
import sys

import numpy as np
import torch

sys.path.insert(0, ".")
from pyranges_b200 import _cabi as cabi
from pyranges_b200 import engine as E
from pyranges_b200 import synthetic as S

_pos = [a for a in sys.argv[1:] if not a.startswith("--")]
nr = int(_pos[0]) if len(_pos) > 0 else 10_000_000
na = int(_pos[1]) if len(_pos) > 1 else 200_000
reads, annot = S.reads(nr), S.annotations(na)
if "--sorted" in sys.argv:  # position-sorted reads inside every key (the natural order of an aligned BAM)
    key = np.repeat(np.arange(reads.n_keys), np.diff(reads.seg_off))
    o = np.lexsort([reads.starts, key])
    reads.starts[:], reads.ends[:] = reads.starts[o], reads.ends[o]
sg = (np.arange(reads.n_keys) // 2).astype(np.int32)
q_iv = E.DeviceIntervals.from_numpy(reads.starts, reads.ends, reads.seg_off)
s_iv = E.DeviceIntervals.from_numpy(annot.starts, annot.ends, annot.seg_off)
L = cabi.lib()
ix = E.SubjectIndex(s_iv, annot.chrom_ranges())
q = E.Queries(q_iv, sg)
counts, ws, total = ix.count(q)
p = int(total.item())
out_q = torch.empty(p, dtype=torch.int64, device="cuda")
out_s = torch.empty(p, dtype=torch.int64, device="cuda")
st = E._stream()


def timeit(name, fn, reps=20, alg_bytes=None):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(reps):
        fn()
    b.record()
    torch.cuda.synchronize()
    us = a.elapsed_time(b) * 1e3 / reps
    extra = "  %.0f GB/s algorithmic" % (alg_bytes / us / 1e3) if alg_bytes else ""
    print("%-32s %9.1f us%s" % (name, us, extra))


def f_count():
    cabi.check(L.iv_count_overlaps(C.byref(ix.c), C.byref(q.c), 0, E._ptr(counts), E._ptr(ws), ws.numel(), E._ptr(total), st))


def f_count_lists():
    cabi.check(L.iv_count_overlaps(C.byref(ix.c), C.byref(q.c), 0, None, E._ptr(ws), ws.numel(), E._ptr(total), st))


def f_count_only():
    cabi.check(L.iv_count_overlaps(C.byref(ix.c), C.byref(q.c), 0, E._ptr(counts), None, 0, None, st))


def f_emit():
    cabi.check(L.iv_emit_pairs(C.byref(ix.c), C.byref(q.c), 0, None, E._ptr(ws), E._ptr(out_q), E._ptr(out_s), None, out_q.numel(), st))


def f_build():
    E.SubjectIndex(s_iv, annot.chrom_ranges())


print("reads %d annotations %d pairs %d" % (nr, na, p))
timeit("iv_count_overlaps (+lists)", f_count, alg_bytes=24 * nr)
timeit("iv_count_overlaps (lists only)", f_count_lists, alg_bytes=16 * nr)
timeit("iv_count_overlaps (no sums)", f_count_only, alg_bytes=24 * nr)
timeit("iv_emit_pairs", f_emit, alg_bytes=16 * p)
timeit("SubjectIndex() python", f_build, reps=10)
timeit("iv_index_build (C call)", ix.build, reps=10, alg_bytes=24 * na)
# subtract against the merged annotations (strand ignored -> chromosome groups)
ms, me, _, mg = E.merge(s_iv, annot.chrom_ranges())
moff = np.concatenate([[0], np.cumsum(np.bincount(mg.cpu().numpy(), minlength=len(annot.chrom_ranges())))])
m_iv = E.DeviceIntervals(ms, me, moff)
m_iv.bounds
mix = E.SubjectIndex(m_iv)
oq, _, _ = mix.subtract(q)
timeit("merge annotations (%d -> %d)" % (na, ms.numel()), lambda: E.merge(s_iv, annot.chrom_ranges()), reps=10)
timeit("subtract (count+scan+emit, %d pieces)" % oq.numel(), lambda: mix.subtract(q), reps=10, alg_bytes=16 * nr + 24 * oq.numel())
x = torch.empty(nr * 3, dtype=torch.int64, device="cuda")
y = torch.empty_like(x)
timeit("torch copy 240 MB (rd+wr)", lambda: y.copy_(x), alg_bytes=2 * 24 * nr)
