"""GPU parity of the C-ABI kernels (through pyranges_b200.engine) against the CPU oracle on seeded inputs.
Bit-exact: every quantity on this path is int64 (the only float, FractionOverlaps, is one IEEE division)."""
import numpy as np
import pytest

from tests.util import canon_pairs, gen_segments, oracle_counts, oracle_pairs

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def eng():
    import torch

    from pyranges_b200 import engine

    assert torch.cuda.is_available()
    return engine


def _dev(eng, s, e, off):
    return eng.DeviceIntervals.from_numpy(s, e, off)


CASES = [
    # (seed, n_seg_q, nq, n_seg_s, ns, maxpos, maxlen)
    (1, 1, 50, 1, 30, 500, 60),
    (2, 4, 2000, 4, 700, 20000, 300),
    (3, 6, 5000, 6, 3000, 200000, 2000),
    (4, 3, 3000, 3, 300, 300, 150),  # deep pile-up: hundreds of hits per query
    (5, 5, 20000, 5, 9000, 3000000, 500),
]


def _setup(eng, case, empty_every=0):
    seed, kq, nq, ks, ns, maxpos, maxlen = case
    qs, qe, qoff = gen_segments(seed, kq, nq, maxpos, maxlen, empty_every=empty_every)
    ss, se, soff = gen_segments(seed + 100, ks, ns, maxpos, maxlen)
    seg_group = np.arange(kq, dtype=np.int32) % ks
    if kq > 2:
        seg_group[1] = -1  # a query key without partner frame
    return qs, qe, qoff, ss, se, soff, seg_group


def test_radix_sort(eng):
    import torch

    rng = np.random.default_rng(0)
    for n, bits in [(1, 64), (1000, 20), (70000, 64), (300001, 37)]:
        k = rng.integers(0, 2 ** min(bits, 63), n, dtype=np.int64)
        v8 = np.arange(n, dtype=np.int64)
        kt, vt = torch.from_numpy(k.copy()).cuda(), torch.from_numpy(v8.copy()).cuda()
        eng.radix_sort_u64(kt, vt, 0, bits)
        o = np.argsort(k, kind="stable")
        assert np.array_equal(kt.cpu().numpy(), k[o])
        assert np.array_equal(vt.cpu().numpy(), v8[o])
        v4 = torch.arange(n, dtype=torch.int32).cuda()
        kt = torch.from_numpy(k.copy()).cuda()
        eng.radix_sort_u64(kt, v4, 0, bits)
        assert np.array_equal(v4.cpu().numpy(), o.astype(np.int32))
        kt = torch.from_numpy(k.copy()).cuda()
        eng.radix_sort_u64(kt, None, 0, bits)
        assert np.array_equal(kt.cpu().numpy(), np.sort(k))


def test_radix_sort_small_input_path(eng):
    """4096 <= n <= 2^18 keys take one partition pass + one bucket pass (sort.cu): uniform keys (bitonic buckets),
    heavily skewed keys (a bucket beyond the shared-memory capacity: block-local chunked radix), duplicate keys
    (stability), every payload width, a begin bit."""
    import torch

    rng = np.random.default_rng(5)
    cases = []
    for n, bits in [(4096, 32), (70000, 40), (200000, 32), (262144, 24), (100000, 9), (50000, 59)]:
        cases.append((rng.integers(0, 2 ** bits, n, dtype=np.int64), 0, bits))
    skew = rng.integers(0, 2 ** 32, 150000, dtype=np.int64)
    skew[:140000] = (7 << 24) | rng.integers(0, 2 ** 20, 140000)          # 93 % of the keys in one bucket
    cases.append((skew, 0, 32))
    dup = rng.integers(0, 50, 60000, dtype=np.int64) << 20                 # 50 distinct keys: stability
    cases.append((dup, 0, 32))
    allsame = np.full(30000, 12345 << 3, dtype=np.int64)
    cases.append((allsame, 0, 24))
    cases.append(((rng.integers(0, 2 ** 30, 90000, dtype=np.int64) << 1) | rng.integers(0, 2, 90000), 1, 31))
    for k, b0, b1 in cases:
        n = len(k)
        sk = (k >> b0) & ((1 << (b1 - b0)) - 1)
        o = np.argsort(sk, kind="stable")
        for vdtype in (None, torch.int32, torch.int64):
            kt = torch.from_numpy(k.copy()).cuda()
            vt = torch.arange(n, dtype=vdtype).cuda() if vdtype is not None else None
            eng.radix_sort_u64(kt, vt, b0, b1)
            assert np.array_equal(kt.cpu().numpy(), k[o]), (n, b0, b1, vdtype)
            if vt is not None:
                assert np.array_equal(vt.cpu().numpy(), o), (n, b0, b1, vdtype)


def test_radix_sort_large_input_path(eng):
    """n > 2^18 keys: one-sweep partition passes on the top bits + one shared-memory pass per bucket (sort.cu).
    Uniform keys, keys with fewer bits than the partition covers (no bucket pass), 64-bit keys (the partition widens
    so that the composite words fit), a pile-up beyond the bucket capacity (global fallback), few distinct keys
    (stability; bins as large as the bucket), identical keys, a begin bit, a ragged last tile."""
    import torch

    rng = np.random.default_rng(11)
    cases = []
    for n, bits in [(262145, 32), (300001, 64), (1 << 20, 42), (700003, 6), (1234567, 17), (3000000, 40)]:
        cases.append((rng.integers(0, 2 ** min(bits, 63), n, dtype=np.int64) | (rng.integers(0, 2, n) << (bits - 1)), 0, bits))
    skew = rng.integers(0, 2 ** 40, 900000, dtype=np.int64)
    skew[100000:700000] = (5 << 36) | rng.integers(0, 2 ** 18, 600000)       # 600 k keys inside one bucket
    cases.append((skew, 0, 40))
    dup = rng.integers(0, 37, 500000, dtype=np.int64) << 13                   # 37 distinct keys
    cases.append((dup, 0, 30))
    two = (rng.integers(0, 2 ** 10, 400000, dtype=np.int64) << 24) | (rng.integers(0, 2, 400000) * 0xFFFFF)
    cases.append((two, 0, 34))                                                # two values of the low bits per bucket
    cases.append((np.full(300000, 99 << 5, dtype=np.int64), 0, 20))
    cases.append(((rng.integers(0, 2 ** 33, 800000, dtype=np.int64) << 1) | rng.integers(0, 2, 800000), 1, 34))
    for k, b0, b1 in cases:
        n = len(k)
        sk = (k.astype(np.uint64) >> np.uint64(b0)) & np.uint64((1 << (b1 - b0)) - 1 if b1 - b0 < 64 else 2 ** 64 - 1)
        o = np.argsort(sk, kind="stable")
        for vdtype in (None, torch.int32, torch.int64):
            kt = torch.from_numpy(k.copy()).cuda()
            vt = torch.arange(n, dtype=vdtype).cuda() if vdtype is not None else None
            eng.radix_sort_u64(kt, vt, b0, b1)
            assert np.array_equal(kt.cpu().numpy(), k[o]), (n, b0, b1, vdtype)
            if vt is not None:
                assert np.array_equal(vt.cpu().numpy(), o), (n, b0, b1, vdtype)


@pytest.mark.parametrize("case", CASES)
@pytest.mark.parametrize("slack", [0, 7])
def test_count_modes(eng, case, slack):
    qs, qe, qoff, ss, se, soff, sg = _setup(eng, case, empty_every=4)
    q = eng.Queries(_dev(eng, qs, qe, qoff), sg, slack=slack)
    ix = eng.SubjectIndex(_dev(eng, ss, se, soff))
    for mode, name in [(eng.IV_MODE_ALL, "all"), (eng.IV_MODE_ANY, "any"), (eng.IV_MODE_CONTAIN, "contain")]:
        c, _, total = ix.count(q, mode)
        want = oracle_counts(qs, qe, qoff, sg, ss, se, soff, name, slack)
        got = c.cpu().numpy()
        assert np.array_equal(got, want), name
        assert int(total.item()) == int(want.sum())


@pytest.mark.parametrize("case", CASES)
def test_pairs(eng, case):
    qs, qe, qoff, ss, se, soff, sg = _setup(eng, case)
    q = eng.Queries(_dev(eng, qs, qe, qoff), sg)
    ix = eng.SubjectIndex(_dev(eng, ss, se, soff))
    for mode, name in [(eng.IV_MODE_ALL, "all"), (eng.IV_MODE_CONTAIN, "contain")]:
        _, a, b = ix.pairs(q, mode)
        wa, wb = canon_pairs(*oracle_pairs(qs, qe, qoff, sg, ss, se, soff, name))
        a, b = a.cpu().numpy(), b.cpu().numpy()
        assert np.all(np.diff(a) >= 0), "pairs must come out in query order"
        ga, gb = canon_pairs(a, b)
        assert np.array_equal(ga, wa) and np.array_equal(gb, wb), name
        # pairs per key, known right after the count pass (before any pair is written)
        _, ws, total = ix.count(q, mode, want_counts=False)
        per_key = ix.segment_pair_counts(q, ws, total).cpu().numpy()
        assert np.array_equal(per_key, np.diff(np.searchsorted(a, qoff))), name
    # first_overlap_both: one pair per hit query, subject = head of the nested-list traversal
    _, a, b = ix.pairs(q, eng.IV_MODE_ANY)
    wa, wb = oracle_pairs(qs, qe, qoff, sg, ss, se, soff, "first")
    assert np.array_equal(a.cpu().numpy(), wa) and np.array_equal(b.cpu().numpy(), wb)
    # last_overlap_both: the tail of the same traversal (largest start, smallest end, largest row)
    _, a, b = ix.pairs(q, eng.IV_MODE_LAST)
    wa, wb = oracle_pairs(qs, qe, qoff, sg, ss, se, soff, "last")
    assert np.array_equal(a.cpu().numpy(), wa) and np.array_equal(b.cpu().numpy(), wb)


def test_pairs_labels(eng):
    import torch

    qs, qe, qoff, ss, se, soff, sg = _setup(eng, CASES[1])
    ql = torch.arange(len(qs), dtype=torch.int64).cuda() * 3 + 5
    sl = torch.arange(len(ss), dtype=torch.int64).cuda() * 7 + 1
    q = eng.Queries(_dev(eng, qs, qe, qoff), sg, label=ql)
    ix = eng.SubjectIndex(_dev(eng, ss, se, soff))
    _, a, b = ix.pairs(q, eng.IV_MODE_ALL, s_label=sl)
    wa, wb = canon_pairs(*oracle_pairs(qs, qe, qoff, sg, ss, se, soff, "all"))
    ga, gb = canon_pairs(a.cpu().numpy(), b.cpu().numpy())
    assert np.array_equal(ga, wa * 3 + 5) and np.array_equal(gb, wb * 7 + 1)


@pytest.mark.parametrize("case", CASES[:4])
def test_coverage_bp(eng, case):
    from oracle import oracle as orc

    qs, qe, qoff, ss, se, soff, sg = _setup(eng, case)
    q = eng.Queries(_dev(eng, qs, qe, qoff), sg)
    ix = eng.SubjectIndex(_dev(eng, ss, se, soff))
    bp, fr = ix.coverage_bp(q)
    want = np.zeros(len(qs), np.int64)
    for k in range(len(qoff) - 1):
        g = sg[k]
        if g < 0 or soff[g] == soff[g + 1]:
            continue
        t = orc.NCLS(ss[soff[g]:soff[g + 1]], se[soff[g]:soff[g + 1]], np.arange(soff[g], soff[g + 1]))
        sl = slice(qoff[k], qoff[k + 1])
        want[sl] = t.coverage(qs[sl], qe[sl], np.arange(qoff[k], qoff[k + 1]))
    assert np.array_equal(bp.cpu().numpy(), want)
    assert np.array_equal(fr.cpu().numpy(), np.nan_to_num(want / (qe - qs)))


def _oracle_nearest(qs, qe, ss, se, mode, overlap):
    """Composition of the reference's _nearest (pyranges/methods/nearest.py:77-132) over the oracle entry
    points, for ONE key; returns (distance, candidate-validity helpers)."""
    from oracle import oracle as orc

    n = len(qs)
    dist = np.full(n, -1, np.int64)
    hit = np.zeros(n, bool)
    if overlap:
        t = orc.NCLS(ss, se, np.arange(len(ss)))
        hit = t.count(qs, qe) > 0
        dist[hit] = 0
    rest = np.flatnonzero(~hit)
    o_e = np.argsort(se, kind="stable")
    o_s = np.argsort(ss, kind="stable")
    ls = np.argsort(qs[rest], kind="stable")
    le = np.argsort(qe[rest], kind="stable")
    pi, pd_ = orc.nearest_previous_nonoverlapping(qs[rest][ls], se[o_e] - 1, o_e)
    ni, nd = orc.nearest_next_nonoverlapping(qe[rest][le] - 1, ss[o_s], o_s)
    pi2, pd2, ni2, nd2 = (np.empty(len(rest), np.int64) for _ in range(4))
    pi2[ls], pd2[ls], ni2[le], nd2[le] = pi, pd_, ni, nd
    if mode == "previous":
        ri, rd = pi2, pd2
    elif mode == "next":
        ri, rd = ni2, nd2
    else:
        ri, rd = orc.nearest_nonoverlapping(pi2, pd2, ni2, nd2)
    dist[rest] = rd
    return dist


@pytest.mark.parametrize("case", CASES[:4])
@pytest.mark.parametrize("overlap", [True, False])
def test_nearest(eng, case, overlap):
    qs, qe, qoff, ss, se, soff, sg = _setup(eng, case)
    kq = len(qoff) - 1
    modes = np.asarray([k % 3 for k in range(kq)], np.int32)  # BOTH / PREVIOUS / NEXT per key
    q = eng.Queries(_dev(eng, qs, qe, qoff), sg, seg_mode=modes)
    ix = eng.SubjectIndex(_dev(eng, ss, se, soff), ends_perm=True)
    row, dist = ix.nearest(q, overlap=overlap)
    row, dist = row.cpu().numpy(), dist.cpu().numpy()
    names = {0: "both", 1: "previous", 2: "next"}
    for k in range(kq):
        sl = slice(qoff[k], qoff[k + 1])
        g = sg[k]
        if g < 0 or soff[g] == soff[g + 1]:
            assert np.all(row[sl] == -1) and np.all(dist[sl] == -1)
            continue
        a, b = soff[g], soff[g + 1]
        want = _oracle_nearest(qs[sl], qe[sl], ss[a:b], se[a:b], names[modes[k]], overlap)
        assert np.array_equal(dist[sl], want), (k, names[modes[k]])
        # the chosen subject must realise the distance and lie in the partner group
        r, d, s, e = row[sl], dist[sl], qs[sl], qe[sl]
        ok = r >= 0
        assert np.array_equal(ok, d >= 0)
        assert np.all((r[ok] >= a) & (r[ok] < b))
        S, E = ss[r[ok]], se[r[ok]]
        gap = np.where(E <= s[ok], s[ok] - E + 1, np.where(S >= e[ok], S - e[ok] + 1, 0))
        assert np.array_equal(gap, d[ok])


@pytest.mark.parametrize("case", CASES[:3] + [(9, 2, 400, 2, 300, 300, 40)])  # the last one: many equal ends / starts
@pytest.mark.parametrize("ties", [None, "first", "last", "different"])
def test_k_nearest_candidates(eng, case, ties):
    """iv_k_nearest_count / _emit against the restatement of sorted_nearest.k_nearest_previous_nonoverlapping /
    k_nearest_next_nonoverlapping (oracle.py) driven the way pyranges/methods/k_nearest.py:6-48 drives them: the same
    (query row, subject row, distance) triples in the same order -- previous candidates (equal ends: descending row)
    before next ones (equal starts: ascending row); BOTH / PREVIOUS / NEXT per key, one k per row (0 .. 4)."""
    import torch

    from tests import fake_engine as fe

    qs, qe, qoff, ss, se, soff, sg = _setup(eng, case)
    kq = len(qoff) - 1
    modes = np.asarray([k % 3 for k in range(kq)], np.int32)
    kk = (np.arange(len(qs)) * 7 % 5).astype(np.int64)
    q = eng.Queries(_dev(eng, qs, qe, qoff), sg, seg_mode=modes)
    ix = eng.SubjectIndex(_dev(eng, ss, se, soff), ends_perm=True)
    a, b, d = (t.cpu().numpy() for t in ix.k_nearest(q, torch.from_numpy(kk).cuda(), ties))
    fq = fe.Queries(fe.DeviceIntervals(qs, qe, qoff), sg, seg_mode=modes)
    wa, wb, wd = (t.numpy() for t in fe.SubjectIndex(fe.DeviceIntervals(ss, se, soff), ends_perm=True).k_nearest(fq, kk, ties))
    assert len(a) == len(wa) and len(a) > 0
    assert np.array_equal(a, wa) and np.array_equal(b, wb) and np.array_equal(d, wd)


@pytest.mark.parametrize("seed,n_seg,n,maxpos,maxlen", [(1, 1, 40, 300, 40), (2, 5, 3000, 50000, 300),
                                                         (3, 7, 20000, 3000000, 2000), (4, 2, 5000, 400, 100)])
@pytest.mark.parametrize("slack", [0, 10, -5])
def test_cluster_merge(eng, seed, n_seg, n, maxpos, maxlen, slack):
    from oracle import oracle as orc

    s, e, off = gen_segments(seed, n_seg, n, maxpos, maxlen, empty_every=3)
    iv = _dev(eng, s, e, off)
    row, cid, nc = eng.cluster(iv, slack=slack)
    row, cid = row.cpu().numpy(), cid.cpu().numpy()
    ms, me, mc, mg = (t.cpu().numpy() for t in eng.merge(iv, slack=slack))
    want_row, want_id, W = [], [], []
    base = 0
    for k in range(n_seg):
        a, b = off[k], off[k + 1]
        if a == b:
            continue
        o = np.lexsort([np.arange(a, b), e[a:b], s[a:b]])  # canonical (Start, End, row) order
        ids = orc.annotate_clusters(s[a:b][o], e[a:b][o], slack)
        want_row.append(o + a)
        want_id.append(ids + base)
        base += ids.max()
        cs, ce, cn = orc.find_clusters(s[a:b][o], e[a:b][o], slack)
        W.append((cs, ce, cn, np.full(len(cs), k)))
    assert np.array_equal(row, np.concatenate(want_row))
    assert np.array_equal(cid, np.concatenate(want_id))
    assert int(nc.item()) == base
    assert np.array_equal(ms, np.concatenate([w[0] for w in W]))
    assert np.array_equal(me, np.concatenate([w[1] for w in W]))
    assert np.array_equal(mc, np.concatenate([w[2] for w in W]))
    assert np.array_equal(mg, np.concatenate([w[3] for w in W]))


@pytest.mark.parametrize("seed,n_seg,n,maxpos,maxlen", [(1, 1, 5, 30, 10), (2, 1, 40, 300, 40), (3, 5, 3000, 50000, 300),
                                                         (4, 7, 20000, 3000000, 2000), (5, 2, 5000, 400, 100)])
def test_coverage_rle(eng, seed, n_seg, n, maxpos, maxlen):
    from oracle import oracle as orc

    s, e, off = gen_segments(seed, n_seg, n, maxpos, maxlen, empty_every=3)
    if seed == 2:
        s[:3] = 0  # a key whose coverage starts at coordinate 0
    iv = _dev(eng, s, e, off)
    run_off, runs, vals = eng.coverage_rle(iv)
    runs, vals = runs.cpu().numpy(), vals.cpu().numpy()
    for k in range(n_seg):
        a, b = off[k], off[k + 1]
        wr, wv = orc.coverage_rle(s[a:b], e[a:b]) if b > a else (np.zeros(0, np.int64), np.zeros(0))
        sl = slice(run_off[k], run_off[k + 1])
        assert np.array_equal(runs[sl], wr), k
        assert np.array_equal(vals[sl], wv), k


def test_coverage_rle_dense_path(eng):
    """Read sets dense enough for the bucket path of to_rle (rle_dense.cu: >= 1 event per 64 coordinates, span >= 2^27):
    every key against pyrle's restatement.  Keys: two large ones, one without rows, one whose coverage starts at
    coordinate 0, one whose first Start is far from 0, zero-length intervals (at 0, in the middle, at the last End),
    a pile-up of identical reads, a tiny key sharing its bucket with the neighbours' origins, reads touching
    bucket boundaries (multiples of 8192)."""
    from oracle import oracle as orc

    rng = np.random.default_rng(17)
    keys = []

    def reads(n, span, lo=0, maxlen=300):
        st = rng.integers(lo, span, n, dtype=np.int64)
        return st, st + rng.integers(1, maxlen, n, dtype=np.int64)

    keys.append(reads(900_000, 90_000_000))
    keys.append((np.zeros(0, np.int64), np.zeros(0, np.int64)))                      # no rows
    st, en = reads(700_000, 70_000_000)
    st[:5] = 0                                                                        # coverage from coordinate 0
    en[:5] = [1, 7, 7, 8192, 8193]
    keys.append((st, en))
    keys.append((np.array([0, 0, 5], np.int64), np.array([0, 0, 9], np.int64)))       # zero-length at coordinate 0
    keys.append((np.array([4], np.int64), np.array([4], np.int64)))                   # only a zero-length interval
    keys.append(reads(30_000, 3_000_000, lo=2_500_000))                               # first Start far from 0
    st, en = reads(60_000, 200_000)
    st[:40_000] = 81_920                                                               # pile-up on a bucket boundary
    en[:40_000] = 81_920 + 100
    st[40_000:40_010] = en[40_000:40_010] = 150_000                                    # zero-length in the middle
    st[-3:] = en[-3:] = int(en.max()) + 17                                             # zero-length at the last End
    keys.append((st, en))
    keys.append((np.array([8192, 16384, 16383], np.int64), np.array([16384, 16385, 16384], np.int64)))
    s = np.concatenate([k[0] for k in keys])
    e = np.concatenate([k[1] for k in keys])
    off = np.concatenate([[0], np.cumsum([len(k[0]) for k in keys])]).astype(np.int64)
    iv = _dev(eng, s, e, off)
    run_off, runs, vals = eng.coverage_rle(iv)
    runs, vals = runs.cpu().numpy(), vals.cpu().numpy()
    assert run_off[-1] == len(runs)
    for k in range(len(keys)):
        a, b = off[k], off[k + 1]
        wr, wv = orc.coverage_rle(s[a:b], e[a:b]) if b > a else (np.zeros(0, np.int64), np.zeros(0))
        sl = slice(run_off[k], run_off[k + 1])
        assert np.array_equal(runs[sl], wr), (k, runs[sl][:8], wr[:8])
        assert np.array_equal(vals[sl], wv), (k, vals[sl][:8], wv[:8])


def test_large_staged_paths(eng):
    """sizes above the staged-kernel threshold (>= 64 tiles of 2048 rows), ragged tail, several key segments,
    a key without partner: counts / pairs must equal the oracle."""
    rng = np.random.default_rng(5)
    kq, ks = 6, 5
    n_per = [40000, 1, 70001, 0, 35000, 20011]
    qs, qe, qoff = [], [], [0]
    for n in n_per:
        s = rng.integers(0, 2_000_000, n)
        qs.append(s)
        qe.append(s + rng.integers(1, 300, n))
        qoff.append(qoff[-1] + n)
    qs, qe, qoff = np.concatenate(qs).astype(np.int64), np.concatenate(qe).astype(np.int64), np.asarray(qoff, np.int64)
    ss, se, soff = gen_segments(77, ks, 6000, 2_000_000, 3000)
    sg = np.asarray([0, 1, 2, 3, -1, 4], np.int32)
    q = eng.Queries(_dev(eng, qs, qe, qoff), sg)
    ix = eng.SubjectIndex(_dev(eng, ss, se, soff))
    for mode, name in [(eng.IV_MODE_ALL, "all"), (eng.IV_MODE_ANY, "any")]:
        c, _, total = ix.count(q, mode)
        want = oracle_counts(qs, qe, qoff, sg, ss, se, soff, name)
        assert np.array_equal(c.cpu().numpy(), want), name
        assert int(total.item()) == int(want.sum())
    wa, wb = canon_pairs(*oracle_pairs(qs, qe, qoff, sg, ss, se, soff, "all"))
    for want_counts in (True, False):  # join does not need the per-row counts: hit lists only
        cnt, a, b = ix.pairs(q, eng.IV_MODE_ALL, want_counts=want_counts)
        assert (cnt is not None) == want_counts
        ga, gb = canon_pairs(a.cpu().numpy(), b.cpu().numpy())
        assert np.array_equal(ga, wa) and np.array_equal(gb, wb)
    _, ws, total = ix.count(q, eng.IV_MODE_ALL, want_counts=False)
    per_key = ix.segment_pair_counts(q, ws, total).cpu().numpy()
    assert np.array_equal(per_key, np.diff(np.searchsorted(np.sort(wa), qoff)))
    # join(slack=k) on the lean kernel's slack instantiation
    qk = eng.Queries(_dev(eng, qs, qe, qoff), sg, slack=25)
    c, _, _ = ix.count(qk, eng.IV_MODE_ALL)
    assert np.array_equal(c.cpu().numpy(), oracle_counts(qs, qe, qoff, sg, ss, se, soff, "all", 25))
    _, a, b = ix.pairs(qk, eng.IV_MODE_ALL, want_counts=False)
    wa, wb = canon_pairs(*oracle_pairs(qs, qe, qoff, sg, ss, se, soff, "all", 25))
    ga, gb = canon_pairs(a.cpu().numpy(), b.cpu().numpy())
    assert np.array_equal(ga, wa) and np.array_equal(gb, wb)
    # the same rows far beyond 32 bits (64-bit instantiation of the lean kernel, 32-byte list entries)
    big = 5_000_000_000_000
    q64 = eng.Queries(_dev(eng, qs + big, qe + big, qoff), sg)
    ss64 = ss + big
    ss64[0] = 0  # one subject at the origin stretches the flattened span of its group beyond 2^32
    se64 = np.maximum(se + big, ss64 + 1)
    ix64 = eng.SubjectIndex(_dev(eng, ss64, se64, soff))
    _, a, b = ix64.pairs(q64, eng.IV_MODE_ALL, want_counts=False)
    wa, wb = canon_pairs(*oracle_pairs(qs + big, qe + big, qoff, sg, ss64, se64, soff, "all"))
    ga, gb = canon_pairs(a.cpu().numpy(), b.cpu().numpy())
    assert np.array_equal(ga, wa) and np.array_equal(gb, wb)


def test_hostapi_chunked_pipeline(eng):
    """numpy/pinned host facade: the chunked three-stream pipeline returns the same pairs as one shot."""
    import torch

    from pyranges_b200 import hostapi

    rng = np.random.default_rng(9)
    n = 300_000
    s = rng.integers(0, 5_000_000, n).astype(np.int64)
    e = s + rng.integers(1, 400, n)
    qoff = np.asarray([0, 100_000, 100_000, 220_001, n], np.int64)
    ss, se, soff = gen_segments(91, 3, 20000, 5_000_000, 5000)
    sg = np.asarray([0, 1, 2, -1], np.int32)
    nb = hostapi.NCLSBatch(ss, se, soff)
    a1, b1 = nb.all_overlaps_both(s, e, qoff, sg)
    hs, he = torch.from_numpy(s).pin_memory(), torch.from_numpy(e).pin_memory()
    out = (torch.empty(len(a1) + 10, dtype=torch.int64).pin_memory(), torch.empty(len(a1) + 10, dtype=torch.int64).pin_memory())
    a2, b2 = nb.all_overlaps_both(hs, he, qoff, sg, out=out, chunks=5)
    assert np.array_equal(a1, a2.numpy()) and np.array_equal(b1, b2.numpy())
    wa, wb = canon_pairs(*oracle_pairs(s, e, qoff, sg, ss, se, soff, "all"))
    ga, gb = canon_pairs(a1, b1)
    assert np.array_equal(ga, wa) and np.array_equal(gb, wb)
    c = nb.count(s, e, qoff, sg)
    assert np.array_equal(c, np.bincount(wa, minlength=n))


def test_pairs_capacity_hint(eng):
    """speculative emission: any capacity guess gives the exact result (too small -> repeated)."""
    qs, qe, qoff, ss, se, soff, sg = _setup(eng, CASES[2])
    q = eng.Queries(_dev(eng, qs, qe, qoff), sg)
    ix = eng.SubjectIndex(_dev(eng, ss, se, soff))
    _, a0, b0 = ix.pairs(q, eng.IV_MODE_ALL)
    for cap in (1, a0.numel() // 2, a0.numel(), a0.numel() + 1000):
        _, a, b = ix.pairs(q, eng.IV_MODE_ALL, capacity=cap)
        assert a.numel() == a0.numel()
        assert bool((a == a0).all()) and bool((b.sort().values == b0.sort().values).all())


@pytest.mark.parametrize("seed,n_seg,n,maxpos,maxlen", [(1, 1, 5, 30, 10), (3, 5, 3000, 50000, 300), (5, 2, 5000, 400, 100)])
@pytest.mark.parametrize("integer_weights", [True, False])
def test_coverage_rle_weighted(eng, seed, n_seg, n, maxpos, maxlen, integer_weights):
    """to_rle(value_col): integer-valued weights are exact; general float weights within 1e-9 (a parallel float64
    sum is not bit-identical to pyrle's sequential one)."""
    import torch

    from oracle import oracle as orc

    s, e, off = gen_segments(seed, n_seg, n, maxpos, maxlen, empty_every=3)
    rng = np.random.default_rng(seed)
    w = rng.integers(0, 1000, len(s)).astype(np.float64) if integer_weights else rng.random(len(s)) * 10 + 0.5
    iv = _dev(eng, s, e, off)
    run_off, runs, vals = eng.coverage_rle(iv, None, torch.from_numpy(w).cuda())
    runs, vals = runs.cpu().numpy(), vals.cpu().numpy()
    for k in range(n_seg):
        a, b = off[k], off[k + 1]
        wr, wv = orc.coverage_rle(s[a:b], e[a:b], w[a:b]) if b > a else (np.zeros(0, np.int64), np.zeros(0))
        sl = slice(run_off[k], run_off[k + 1])
        if integer_weights:
            assert np.array_equal(runs[sl], wr) and np.array_equal(vals[sl], wv), k
        else:
            # run boundaries can only differ where a level changes by rounding noise: compare the step functions
            def expand(r, v):
                return np.repeat(v, r)
            assert int(runs[sl].sum()) == int(wr.sum())
            assert np.allclose(expand(runs[sl], vals[sl]), expand(wr, wv), rtol=1e-9, atol=1e-9), k


@pytest.mark.parametrize("seed", [1, 2, 3, 4, 5, 6])
def test_subtract(eng, seed):
    """iv_subtract_count/_emit vs the oracle's set_difference_helper on merged (disjoint) subjects; seeds 5 and 6 are
    large enough for the lean count kernel (most rows untouched / most rows cut)."""
    from oracle import oracle as orc

    rng = np.random.default_rng(seed)
    maxpos = int(rng.choice([500, 100_000, 8_000_000_000]))
    nq, ns = (2000, 800) if seed < 5 else (90_000, 3000 if seed == 5 else 60_000)
    if seed >= 5:
        maxpos = 30_000_000
    qs, qe, qoff = gen_segments(seed, 4, nq, min(maxpos, 10**9), 300)
    ss, se, soff = gen_segments(seed + 9, 3, ns, min(maxpos, 10**9), 300)
    if maxpos > 2**32:
        qs, qe, ss, se = qs * 9, qe * 9, ss * 9, se * 9
    # merge the subjects per key on the device first (PyRanges.subtract does the same)
    ms, me, _, mg = eng.merge(_dev(eng, ss, se, soff))
    moff = np.concatenate([[0], np.cumsum(np.bincount(mg.cpu().numpy(), minlength=3))])
    ix = eng.SubjectIndex(eng.DeviceIntervals(ms, me, moff))
    sg = np.asarray([0, 1, -1, 2], np.int32)
    oq, os_, oe = ix.subtract(eng.Queries(_dev(eng, qs, qe, qoff), sg))
    oq, os_, oe = oq.cpu().numpy(), os_.cpu().numpy(), oe.cpu().numpy()
    ms, me = ms.cpu().numpy(), me.cpu().numpy()
    WQ, WS, WE = [], [], []
    for k in range(4):
        a, b = qoff[k], qoff[k + 1]
        rows = np.arange(a, b)
        g = sg[k]
        if g < 0 or moff[g] == moff[g + 1]:
            WQ.append(rows), WS.append(qs[a:b]), WE.append(qe[a:b])
            continue
        t = orc.NCLS(ms[moff[g]:moff[g + 1]], me[moff[g]:moff[g + 1]], np.arange(moff[g], moff[g + 1]))
        idx, ns, ne = t.set_difference_helper(qs[a:b], qe[a:b], rows)
        untouched = np.setdiff1d(rows, idx)
        keep = ns != -1
        q2 = np.concatenate([untouched, idx[keep]])
        s2 = np.concatenate([qs[untouched], ns[keep]])
        e2 = np.concatenate([qe[untouched], ne[keep]])
        o = np.lexsort([s2, q2])
        WQ.append(q2[o]), WS.append(s2[o]), WE.append(e2[o])
    assert np.array_equal(oq, np.concatenate(WQ))
    assert np.array_equal(os_, np.concatenate(WS)) and np.array_equal(oe, np.concatenate(WE))


@pytest.mark.parametrize("case", CASES[1:4])
def test_materialize_pairs_and_gather(eng, case):
    """Device-side join tail: coordinate columns, Overlap, the clip of intersect and payload
    gathers equal numpy fancy indexing on the same pair lists."""
    import torch

    qs, qe, qoff, ss, se, soff, sg = _setup(eng, case)
    qiv, siv = _dev(eng, qs, qe, qoff), _dev(eng, ss, se, soff)
    _, a, b = eng.SubjectIndex(siv).pairs(eng.Queries(qiv, sg), eng.IV_MODE_ALL)
    an, bn = a.cpu().numpy(), b.cpu().numpy()
    assert len(an) > 0
    got = eng.materialize_pairs(qiv, siv, a, b, want=("q_start", "q_end", "s_start", "s_end", "overlap"))
    got = {k: v.cpu().numpy() for k, v in got.items()}
    assert np.array_equal(got["q_start"], qs[an]) and np.array_equal(got["q_end"], qe[an])
    assert np.array_equal(got["s_start"], ss[bn]) and np.array_equal(got["s_end"], se[bn])
    assert np.array_equal(got["overlap"], np.minimum(qe[an], se[bn]) - np.maximum(qs[an], ss[bn]))
    assert (got["overlap"] > 0).all()
    clip = eng.materialize_pairs(qiv, siv, a, b, clip=True, want=("q_start", "q_end"))
    assert np.array_equal(clip["q_start"].cpu().numpy(), np.maximum(qs[an], ss[bn]))
    assert np.array_equal(clip["q_end"].cpu().numpy(), np.minimum(qe[an], se[bn]))
    # payload columns of every element width, with the -1 filler rows of outer joins
    rng = np.random.default_rng(case[0])
    rows = bn.copy()
    rows[rng.random(len(rows)) < 0.1] = -1
    rt = torch.from_numpy(rows).cuda()
    for dtype, fill in [(np.int8, -1), (np.int16, -1), (np.int32, -1), (np.int64, -1), (np.float32, -1.0), (np.float64, -1.0)]:
        col = (rng.integers(-100, 100, len(ss))).astype(dtype)
        out = eng.gather(torch.from_numpy(col).cuda(), rt, fill=fill).cpu().numpy()
        want = np.where(rows >= 0, col[np.maximum(rows, 0)], np.asarray(fill, dtype))
        assert out.dtype == col.dtype and np.array_equal(out, want), dtype
    out = eng.gather(torch.from_numpy(ss).cuda(), rt).cpu().numpy()  # default filler: zero
    assert np.array_equal(out, np.where(rows >= 0, ss[np.maximum(rows, 0)], 0))
    # no pairs at all
    e = torch.zeros(0, dtype=torch.int64).cuda()
    assert all(v.numel() == 0 for v in eng.materialize_pairs(qiv, siv, e, e).values())


@pytest.mark.parametrize("seed,n_seg,n,maxpos,maxlen,shift", [
    (1, 1, 5, 30, 10, 0), (2, 5, 3000, 50000, 300, 0), (3, 7, 20000, 3000000, 2000, 0), (4, 2, 5000, 400, 100, 0),
    (5, 300, 40, 100000, 500, 0),              # many small groups (cluster / merge / split by=)
    (6, 4, 2000, 10 ** 6, 5000, -(10 ** 12)),  # negative coordinates beyond 32 bits
    (7, 3, 300000, 2 * 10 ** 8, 1000, 0),      # several radix passes, multi-tile scans
])
def test_split(eng, seed, n_seg, n, maxpos, maxlen, shift):
    """iv_split_count/_emit = consecutive distinct Start/End points per group (pyranges/methods/split.py:4-24), also
    with segments paired into groups (strand=False on stranded keys)."""
    from oracle import oracle as orc

    s, e, off = gen_segments(seed, n_seg, n, maxpos, maxlen, empty_every=3)
    s, e = s + shift, e + shift
    iv = _dev(eng, s, e, off)
    pairs = [(i, min(i + 2, n_seg)) for i in range(0, n_seg, 2)]
    for seg_ranges in (None, pairs):
        goff, st, en = eng.split(iv, seg_ranges)
        st, en = st.cpu().numpy(), en.cpu().numpy()
        ranges = seg_ranges or [(i, i + 1) for i in range(n_seg)]
        ws, we, woff = [], [], [0]
        for a, b in ranges:
            ra, rb = off[a], off[b]
            fs, fe = orc.split_points(s[ra:rb], e[ra:rb]) if rb > ra else (np.zeros(0, np.int64),) * 2
            ws.append(fs)
            we.append(fe)
            woff.append(woff[-1] + len(fs))
        assert np.array_equal(goff, np.asarray(woff))
        assert np.array_equal(st, np.concatenate(ws)) and np.array_equal(en, np.concatenate(we))


def test_split_empty(eng):
    z = np.zeros(0, np.int64)
    goff, st, en = eng.split(_dev(eng, z, z, np.zeros(3, np.int64)))
    assert goff.tolist() == [0, 0, 0] and st.numel() == 0 and en.numel() == 0


def test_index_build_graph_replay_with_changed_data(eng):
    """IV_INDEX_GRAPH: builds 2, 3, ... over the same buffers on a non-default stream capture / replay one launch
    sequence; the DATA at those addresses changes between builds (rows permuted inside their segments: same bounds,
    same launch sequence) and every result must follow it (oracle check); the last build drops the flag: plain
    launches on the same buffers."""
    import torch

    qs, qe, qoff = gen_segments(41, 2, 3000, 40000, 200)
    sg = np.arange(2, dtype=np.int32)
    st = torch.cuda.Stream()
    s0, e0, soff = gen_segments(42, 2, 1000, 40000, 300)
    ds = torch.from_numpy(s0).cuda()
    de = torch.from_numpy(e0).cuda()
    rng = np.random.default_rng(7)
    with torch.cuda.stream(st):
        q = eng.Queries(_dev(eng, qs, qe, qoff), sg)
        iv = eng.DeviceIntervals(ds, de, soff)
        ix = eng.SubjectIndex(iv, graph=True)
        for rep in range(5):
            perm = np.concatenate([soff[k] + rng.permutation(soff[k + 1] - soff[k]) for k in range(2)])
            s, e = s0[perm], e0[perm]
            ds.copy_(torch.from_numpy(s))
            de.copy_(torch.from_numpy(e))
            if rep == 4:
                ix.flags &= ~eng.cabi.IV_INDEX_GRAPH
            ix.build()
            counts, a, b = ix.pairs(q, eng.IV_MODE_ALL)
            st.synchronize()
            wa, wb = canon_pairs(*oracle_pairs(qs, qe, qoff, sg, s, e, soff, "all"))
            ga, gb = canon_pairs(a.cpu().numpy(), b.cpu().numpy())
            assert np.array_equal(ga, wa) and np.array_equal(gb, wb), rep
