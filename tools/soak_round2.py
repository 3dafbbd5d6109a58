"""Randomised soak of the round-2 kernels against independent references (torch's stable sort, the oracle):
the hybrid radix sort (key widths, payloads, pile-ups of every size class), the dense to_rle path (random key
structure, pile-ups, zero-length intervals, keys without rows), k-nearest candidates, the pipelined join kernels against
the rank-index join.  Usage: python tools/soak_round2.py [first_seed] [last_seed]"""
import sys
import time

import numpy as np
import torch

sys.path.insert(0, ".")
from oracle import oracle as orc
from pyranges_b200 import engine as E
from tests import fake_engine as fe

a = int(sys.argv[1]) if len(sys.argv) > 1 else 0
b = int(sys.argv[2]) if len(sys.argv) > 2 else 40


def soak_sort(seed):
    rng = np.random.default_rng(seed)
    n = int(rng.integers(270_000, 6_000_000))
    bits = int(rng.choice([9, 17, 26, 33, 42, 53, 64]))
    k = rng.integers(0, 2 ** min(bits, 63), n, dtype=np.int64)
    if bits == 64:
        k |= rng.integers(0, 2, n).astype(np.int64) << 63
    for _ in range(int(rng.integers(0, 4))):  # pile-ups: one value of the top bits, or one exact key
        m = int(rng.choice([3000, 20_000, 90_000, 600_000]))
        m = min(m, n // 2)
        at = int(rng.integers(0, n - m))
        top = int(rng.integers(0, 2 ** min(bits, 62)))
        low = max(0, bits - int(rng.integers(8, 30)))
        k[at:at + m] = (top >> low << low) | (rng.integers(0, 2 ** low, m, dtype=np.int64) if low else 0)
        if rng.random() < 0.3:
            k[at:at + m] = top
    vb = rng.choice([0, 4, 8])
    kt = torch.from_numpy(k).cuda()
    vt = None if vb == 0 else torch.arange(n, dtype=torch.int32 if vb == 4 else torch.int64, device="cuda")
    ref = torch.from_numpy(k).cuda()
    if bits < 64:
        order = torch.argsort(ref, stable=True)
    else:  # unsigned order of the 64-bit pattern
        order = torch.argsort(ref ^ (-2 ** 63), stable=True)
    E.radix_sort_u64(kt, vt, 0, bits)
    assert torch.equal(kt, ref[order]), ("sort keys", seed, n, bits, vb)
    if vt is not None:
        assert torch.equal(vt.to(torch.int64), order), ("sort payload", seed, n, bits, vb)
    return n


def soak_rle(seed):
    rng = np.random.default_rng(seed)
    nk = int(rng.integers(1, 9))
    s, e, off = [], [], [0]
    for k in range(nk):
        n = 0 if rng.random() < 0.15 else int(rng.integers(1, 700_000))
        span = int(rng.choice([2_000, 300_000, 30_000_000, 120_000_000]))
        st = rng.integers(0 if rng.random() < 0.5 else int(rng.integers(0, span)), span + 1, n, dtype=np.int64)
        ln = rng.integers(0 if rng.random() < 0.3 else 1, int(rng.choice([2, 150, 20_000])), n, dtype=np.int64)
        if n > 10 and rng.random() < 0.5:  # a pile-up
            m = int(rng.integers(2, n))
            st[:m] = st[0]
            ln[:m] = ln[0] if rng.random() < 0.5 else ln[:m]
        s.append(st)
        e.append(st + ln)
        off.append(off[-1] + n)
    # one large key so that the bucket path applies (span >= 2^27, >= 1 event per 64 coordinates)
    n = int(rng.integers(1_100_000, 1_600_000))
    st = rng.integers(0, 135_000_000, n, dtype=np.int64)
    s.append(st)
    e.append(st + rng.integers(1, 400, n, dtype=np.int64))
    off.append(off[-1] + n)
    o = rng.permutation(len(s))
    s, e = [s[i] for i in o], [e[i] for i in o]
    off = np.concatenate([[0], np.cumsum([len(x) for x in s])]).astype(np.int64)
    s, e = np.concatenate(s), np.concatenate(e)
    iv = E.DeviceIntervals.from_numpy(s, e, off)
    run_off, runs, vals = E.coverage_rle(iv)
    runs, vals = runs.cpu().numpy(), vals.cpu().numpy()
    for k in range(len(off) - 1):
        x, y = off[k], off[k + 1]
        wr, wv = orc.coverage_rle(s[x:y], e[x:y]) if y > x else (np.zeros(0, np.int64), np.zeros(0))
        sl = slice(run_off[k], run_off[k + 1])
        assert np.array_equal(runs[sl], wr) and np.array_equal(vals[sl], wv), ("rle", seed, k)
    return len(s)


def soak_join(seed):
    rng = np.random.default_rng(seed)
    from tests.util import gen_segments

    kq = int(rng.integers(1, 6))
    nq = int(rng.integers(5_000, 400_000))
    ns = int(rng.integers(100, 60_000))
    maxpos = int(rng.choice([5_000, 2_000_000, 300_000_000]))
    qs, qe, qoff = gen_segments(seed, kq, nq, maxpos, int(rng.choice([2, 150, 3000])))
    ss, se, soff = gen_segments(seed + 7, kq, ns, maxpos, int(rng.choice([50, 5000, 200_000])))
    if rng.random() < 0.5:  # position-ordered rows: the other join kernel
        for k in range(kq):
            x, y = qoff[k], qoff[k + 1]
            o = np.argsort(qs[x:y], kind="stable")
            qs[x:y], qe[x:y] = qs[x:y][o], qe[x:y][o]
    sg = np.arange(kq, dtype=np.int32)
    q = E.Queries(E.DeviceIntervals.from_numpy(qs, qe, qoff), sg)
    siv = E.DeviceIntervals.from_numpy(ss, se, soff)
    if not E.ListIndex.supported(siv, None, float((qe - qs).mean())):
        return 0
    lx = E.ListIndex(siv, reach=E._max_len(q))
    if int(lx.count(q, E.IV_MODE_ALL).sum().item()) > 30_000_000:
        return 0  # (bounded memory: skip the cases where nearly everything overlaps everything)
    a1, b1 = lx.pairs(q)
    _, a2, b2 = E.SubjectIndex(siv).pairs(q, E.IV_MODE_ALL)
    k1 = torch.sort(a1 * (len(ss) + 1) + b1).values
    k2 = torch.sort(a2 * (len(ss) + 1) + b2).values
    assert torch.equal(k1, k2), ("join", seed)
    assert bool((a1[1:] >= a1[:-1]).all()), ("join order", seed)
    cnt = lx.count(q, E.IV_MODE_ANY)
    oq = torch.empty(len(qs), dtype=torch.int64, device="cuda")
    t = int(lx.join_pairs(q, oq, None, len(qs), mode=E.IV_MODE_ANY).item())
    assert torch.equal(oq[:t], torch.nonzero(cnt).flatten()), ("overlap rows", seed)
    return int(a1.numel())


t0 = time.time()
tot = [0, 0, 0]
for seed in range(a, b):
    tot[0] += soak_sort(seed)
    tot[1] += soak_rle(seed)
    tot[2] += soak_join(seed)
print("seeds %d..%d: %d keys sorted, %d intervals run-length encoded, %d pairs joined: all equal to the references (%.0f s)"
      % (a, b - 1, tot[0], tot[1], tot[2], time.time() - t0))
