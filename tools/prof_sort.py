"""Radix sort of n 64-bit keys (+ 4-byte payload) on `bits` bits, for ncu captures and throughput numbers.
Usage: python tools/prof_sort.py [n] [bits] [value_bytes] [pile]   pile = fraction of the keys inside ONE bucket"""
import sys

import torch

sys.path.insert(0, ".")
from pyranges_b200 import engine as E

n = int(sys.argv[1]) if len(sys.argv) > 1 else 100_000_000
bits = int(sys.argv[2]) if len(sys.argv) > 2 else 32
vb = int(sys.argv[3]) if len(sys.argv) > 3 else 4
g = torch.Generator(device="cuda").manual_seed(1)
src = torch.randint(0, 2 ** min(bits, 62), (n,), dtype=torch.int64, device="cuda", generator=g)
pile = float(sys.argv[4]) if len(sys.argv) > 4 else 0.0
if pile > 0:  # a pile-up: the first pile * n keys share their top 22 bits
    m = int(n * pile)
    src[:m] = (src[:m] & ((1 << max(bits - 22, 1)) - 1)) | (5 << max(bits - 4, 1))
val = None if vb == 0 else torch.arange(n, dtype=torch.int32 if vb == 4 else torch.int64, device="cuda")
k = src.clone()
for _ in range(2):
    k.copy_(src)
    E.radix_sort_u64(k, val, 0, bits)
torch.cuda.synchronize()
a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
ms = 0.0
for _ in range(3):
    k.copy_(src)
    a.record()
    E.radix_sort_u64(k, val, 0, bits)
    b.record()
    torch.cuda.synchronize()
    ms += a.elapsed_time(b) / 3
passes = (bits + 7) // 8
print("sort %d keys, %d bits (%d LSD passes), pile %.3f, %d-byte payload: %.3f ms = %.1f G keys/s per pass, %.0f GB/s of pass traffic"
      % (n, bits, passes, pile, vb, ms, n * passes / ms / 1e6, n * passes * (24 + 2 * vb) / ms / 1e6))
assert bool((k[1:] >= k[:-1]).all())
