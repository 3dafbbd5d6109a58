"""GPU: randomised parity sweep against the oracle -- small inputs, many shapes: empty / single-row keys, negative
and > 2^32 coordinates (64-bit flattened axis), very long subjects (coarse bins), dense pile-ups (overflowing
bins), bookended and duplicate intervals, random slack."""
import numpy as np
import pytest

from tests.util import canon_pairs, oracle_counts, oracle_pairs

pytestmark = pytest.mark.gpu


def _rand_segments(rng, n_seg, max_n, lo, span, maxlen):
    S, E, off = [], [], [0]
    for _ in range(n_seg):
        n = int(rng.integers(0, max_n + 1)) if rng.random() > 0.15 else int(rng.integers(0, 2))
        s = lo + rng.integers(0, max(span, 1), n)
        ln = rng.integers(1, max(maxlen, 2), n)
        if n > 2:
            k = rng.integers(0, n, n // 4)
            s[k] = s[rng.integers(0, n, len(k))]                      # duplicate starts
            j = rng.integers(0, n, n // 4)
            src = rng.integers(0, n, len(j))
            s[j] = s[src] + ln[src]                                   # bookended neighbours
            big = rng.random(n) < 0.05
            ln[big] = rng.integers(1, max(span, 2), big.sum())        # very long intervals
        S.append(s.astype(np.int64))
        E.append((s + ln).astype(np.int64))
        off.append(off[-1] + n)
    return np.concatenate(S), np.concatenate(E), np.asarray(off, np.int64)


@pytest.mark.parametrize("seed", range(40))
def test_fuzz_binary_ops(seed):
    from oracle import oracle as orc
    from pyranges_b200 import engine as E

    rng = np.random.default_rng(1000 + seed)
    lo = int(rng.choice([0, 0, -50, 7_000_000_000, -3_000_000_000_000]))
    span = int(rng.choice([60, 3000, 2_000_000, 6_000_000_000, 900_000_000_000]))
    maxlen = int(rng.choice([3, 200, 50_000]))
    kq, ks = int(rng.integers(1, 6)), int(rng.integers(1, 6))
    qs, qe, qoff = _rand_segments(rng, kq, 300, lo, span, maxlen)
    ss, se, soff = _rand_segments(rng, ks, 300, lo, span, maxlen)
    sg = rng.integers(-1, ks, kq).astype(np.int32)
    slack = int(rng.choice([0, 0, 3, 1000]))
    if lo < 0:
        slack = 0  # extend() clamps at 0 (pyranges/multithreaded.py:420-423): only meaningful for non-negative data
    q = E.Queries(E.DeviceIntervals.from_numpy(qs, qe, qoff), sg, slack=slack)
    ix = E.SubjectIndex(E.DeviceIntervals.from_numpy(ss, se, soff), ends_perm=True)
    for mode, name in [(E.IV_MODE_ALL, "all"), (E.IV_MODE_ANY, "any"), (E.IV_MODE_CONTAIN, "contain")]:
        c, _, total = ix.count(q, mode)
        want = oracle_counts(qs, qe, qoff, sg, ss, se, soff, name, slack)
        assert np.array_equal(c.cpu().numpy(), want), (seed, name)
        assert int(total.item()) == int(want.sum())
    for mode, name in [(E.IV_MODE_ALL, "all"), (E.IV_MODE_CONTAIN, "contain")]:
        cnt, a, b = ix.pairs(q, mode, want_counts=bool(seed % 2))  # with and without the per-row counts
        assert (cnt is not None) == bool(seed % 2)
        wa, wb = canon_pairs(*oracle_pairs(qs, qe, qoff, sg, ss, se, soff, name, slack))
        ga, gb = canon_pairs(a.cpu().numpy(), b.cpu().numpy())
        assert np.array_equal(ga, wa) and np.array_equal(gb, wb), (seed, name)
    for mode, name in [(E.IV_MODE_ANY, "first"), (E.IV_MODE_LAST, "last")]:
        _, a, b = ix.pairs(q, mode)
        wa, wb = oracle_pairs(qs, qe, qoff, sg, ss, se, soff, name, slack)
        assert np.array_equal(a.cpu().numpy(), wa) and np.array_equal(b.cpu().numpy(), wb), (seed, name)
    if slack == 0:
        bp, _ = ix.coverage_bp(q)
        want = np.zeros(len(qs), np.int64)
        for k in range(kq):
            g = sg[k]
            if g < 0 or soff[g] == soff[g + 1] or qoff[k] == qoff[k + 1]:
                continue
            t = orc.NCLS(ss[soff[g]:soff[g + 1]], se[soff[g]:soff[g + 1]], np.arange(soff[g], soff[g + 1]))
            sl = slice(qoff[k], qoff[k + 1])
            want[sl] = t.coverage(qs[sl], qe[sl], np.arange(qoff[k], qoff[k + 1]))
        assert np.array_equal(bp.cpu().numpy(), want), seed
        # nearest: distances + validity of the chosen subject
        from tests.test_gpu_engine import _oracle_nearest

        modes = rng.integers(0, 3, kq).astype(np.int32)
        for overlap in (True, False):
            row, dist = ix.nearest(E.Queries(E.DeviceIntervals.from_numpy(qs, qe, qoff), sg, seg_mode=modes), overlap=overlap)
            row, dist = row.cpu().numpy(), dist.cpu().numpy()
            for k in range(kq):
                sl = slice(qoff[k], qoff[k + 1])
                g = sg[k]
                if g < 0 or soff[g] == soff[g + 1]:
                    assert np.all(row[sl] == -1) and np.all(dist[sl] == -1)
                    continue
                a0, b0 = soff[g], soff[g + 1]
                want = _oracle_nearest(qs[sl], qe[sl], ss[a0:b0], se[a0:b0], {0: "both", 1: "previous", 2: "next"}[modes[k]], overlap)
                assert np.array_equal(dist[sl], want), (seed, k, overlap)
                ok = row[sl] >= 0
                r = row[sl][ok]
                assert np.all((r >= a0) & (r < b0))
                S, En, s, e = ss[r], se[r], qs[sl][ok], qe[sl][ok]
                gap = np.where(En <= s, s - En + 1, np.where(S >= e, S - e + 1, 0))
                assert np.array_equal(gap, dist[sl][ok])


@pytest.mark.parametrize("seed", range(30))
def test_fuzz_unary_ops(seed):
    from oracle import oracle as orc
    from pyranges_b200 import engine as E

    rng = np.random.default_rng(5000 + seed)
    lo = int(rng.choice([0, 0, 11, 4_000_000_000]))
    span = int(rng.choice([40, 5000, 3_000_000, 70_000_000_000]))
    maxlen = int(rng.choice([3, 200, 20_000]))
    k = int(rng.integers(1, 7))
    s, e, off = _rand_segments(rng, k, 500, lo, span, maxlen)
    if rng.random() < 0.3 and len(s):  # sorted input exercises the skip-the-sort path
        key = np.repeat(np.arange(k), np.diff(off))
        o = np.lexsort([e, s, key])
        s, e = s[o], e[o]
    slack = int(rng.choice([0, 0, 5, -2, 100000]))
    iv = E.DeviceIntervals.from_numpy(s, e, off)
    row, cid, nc = E.cluster(iv, slack=slack)
    ms, me, mc, mg = (t.cpu().numpy() for t in E.merge(iv, slack=slack))
    want_row, want_id, W, base = [np.zeros(0, np.int64)], [np.zeros(0, np.int64)], [], 0
    for g in range(k):
        a, b = off[g], off[g + 1]
        if a == b:
            continue
        o = np.lexsort([np.arange(a, b), e[a:b], s[a:b]])
        ids = orc.annotate_clusters(s[a:b][o], e[a:b][o], slack)
        want_row.append(o + a)
        want_id.append(ids + base)
        base += ids.max()
        cs, ce, cn = orc.find_clusters(s[a:b][o], e[a:b][o], slack)
        W.append((cs, ce, cn, np.full(len(cs), g)))
    assert np.array_equal(row.cpu().numpy(), np.concatenate(want_row)), seed
    assert np.array_equal(cid.cpu().numpy(), np.concatenate(want_id)), seed
    assert int(nc.item()) == base
    for i, got in enumerate((ms, me, mc, mg)):
        assert np.array_equal(got, np.concatenate([w[i] for w in W]) if W else np.zeros(0)), (seed, i)
    if lo >= 0:
        run_off, runs, vals = E.coverage_rle(iv)
        runs, vals = runs.cpu().numpy(), vals.cpu().numpy()
        for g in range(k):
            a, b = off[g], off[g + 1]
            wr, wv = orc.coverage_rle(s[a:b], e[a:b]) if b > a else (np.zeros(0, np.int64), np.zeros(0))
            sl = slice(run_off[g], run_off[g + 1])
            assert np.array_equal(runs[sl], wr) and np.array_equal(vals[sl], wv), (seed, g)
