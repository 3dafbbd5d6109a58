"""Shared helpers of the test-suite: seeded interval generators and per-group oracle drivers.

The oracle (oracle/) is the CHECKER here; the thing under test is always pyranges_b200."""
import numpy as np

from oracle import oracle as orc


def gen_segments(seed, n_seg, n_per_seg, maxpos=5000, maxlen=300, nested=True, empty_every=0, sort=False):
    """-> (starts, ends, seg_off): n_seg key segments with nested / duplicate / bookended intervals."""
    rng = np.random.default_rng(seed)
    S, E, off = [], [], [0]
    for k in range(n_seg):
        n = 0 if (empty_every and k % empty_every == empty_every - 1) else int(rng.integers(max(1, n_per_seg // 2), n_per_seg + 1))
        s = rng.integers(0, maxpos, n)
        ln = rng.integers(1, maxlen, n)
        if nested and n:
            big = rng.random(n) < 0.1
            ln[big] = rng.integers(maxlen * 5, maxlen * 20, big.sum())
            src = rng.integers(0, n, n)
            dup = rng.random(n) < 0.1
            s[dup], ln[dup] = s[src[dup]], ln[src[dup]]
            bk = rng.random(n) < 0.1
            s[bk] = s[src[bk]] + ln[src[bk]]
        e = s + ln
        if sort and n:
            o = np.lexsort([e, s])
            s, e = s[o], e[o]
        S.append(s.astype(np.int64))
        E.append(e.astype(np.int64))
        off.append(off[-1] + n)
    return (np.concatenate(S) if S else np.zeros(0, np.int64), np.concatenate(E) if E else np.zeros(0, np.int64),
            np.asarray(off, np.int64))


def canon_pairs(a, b):
    a, b = np.asarray(a, np.int64), np.asarray(b, np.int64)
    o = np.lexsort([b, a])
    return a[o], b[o]


def oracle_pairs(qs, qe, q_off, seg_group, ss, se, s_off, mode="all", slack=0):
    """Per query segment: NCLS of the partner group (rows of the concatenated subject layout as ids)."""
    A, B = [], []
    for k in range(len(q_off) - 1):
        g = seg_group[k]
        if g < 0:
            continue
        a, b = s_off[g], s_off[g + 1]
        if a == b:
            continue
        t = orc.NCLS(ss[a:b], se[a:b], np.arange(a, b, dtype=np.int64))
        s, e = qs[q_off[k]:q_off[k + 1]].copy(), qe[q_off[k]:q_off[k + 1]].copy()
        if slack:
            s = np.maximum(s - slack, 0)
            e = e + slack
        ids = np.arange(q_off[k], q_off[k + 1], dtype=np.int64)
        if mode == "all":
            x, y = t.all_overlaps_both(s, e, ids)
        elif mode == "contain":
            x, y = t.all_containments_both(s, e, ids)
        elif mode == "last":
            x, y = t.last_overlap_both(s, e, ids)
        else:
            x, y = t.first_overlap_both(s, e, ids)
        A.append(x)
        B.append(y)
    if not A:
        return np.zeros(0, np.int64), np.zeros(0, np.int64)
    return np.concatenate(A), np.concatenate(B)


def oracle_counts(qs, qe, q_off, seg_group, ss, se, s_off, mode="all", slack=0):
    a, _ = oracle_pairs(qs, qe, q_off, seg_group, ss, se, s_off, "contain" if mode == "contain" else "all", slack)
    c = np.bincount(a, minlength=len(qs)).astype(np.int64)
    return (c > 0).astype(np.int64) if mode == "any" else c
