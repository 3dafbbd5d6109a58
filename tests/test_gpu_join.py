"""GPU parity of the list index + fused join (csrc/lists.cu, csrc/join.cu; engine.ListIndex) against the CPU oracle
and against brute force: pairs after a canonical sort, counts, query order, labels, slack, capacity handling, the
staging-overflow and long-list paths, multi-bin queries, 64-bit / negative coordinates, zero-length intervals."""
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


def brute_pairs(qs, qe, qoff, sg, ss, se, soff):
    """the reference predicate S < e and s < E, key by key (pyranges_main.py:3563-3572)"""
    A, B = [], []
    for k in range(len(qoff) - 1):
        g = sg[k]
        if g < 0:
            continue
        a, b = qoff[k], qoff[k + 1]
        sa, sb = soff[g], soff[g + 1]
        hit = (ss[None, sa:sb] < qe[a:b, None]) & (qs[a:b, None] < se[None, sa:sb])
        i, j = np.nonzero(hit)
        A.append(i + a)
        B.append(j + sa)
    if not A:
        return np.zeros(0, np.int64), np.zeros(0, np.int64)
    return canon_pairs(np.concatenate(A), np.concatenate(B))


CASES = [
    # (seed, n_seg_q, nq, n_seg_s, ns, maxpos, maxlen)
    (1, 1, 50, 1, 30, 500, 60),
    (2, 4, 2000, 4, 700, 20000, 300),
    (3, 6, 5000, 6, 3000, 200000, 2000),
    (4, 3, 3000, 3, 300, 300, 150),        # deep pile-up: hundreds of hits per query, long lists, staging overflow
    (5, 5, 20000, 5, 9000, 3000000, 500),
    (6, 2, 9000, 2, 40000, 30000000, 3000),  # many subjects: sparse lists, queries reach several bins
    (7, 8, 700, 3, 5000, 100000, 40),
]


def _setup(case, empty_every=0):
    seed, kq, nq, ks, ns, maxpos, maxlen = case
    qs, qe, qoff = gen_segments(seed, kq, nq, maxpos, maxlen, empty_every=empty_every)
    ss, se, soff = gen_segments(seed + 100, ks, ns, maxpos, maxlen)
    seg_group = np.arange(kq, dtype=np.int32) % ks
    if kq > 2:
        seg_group[1] = -1  # a query key without partner frame
    return qs, qe, qoff, ss, se, soff, seg_group


@pytest.mark.parametrize("case", CASES)
@pytest.mark.parametrize("slack", [0, 7])
@pytest.mark.parametrize("reach", ["none", "max", "half"])
def test_join_pairs_vs_oracle(eng, case, slack, reach):
    """reach: the index also lists subjects that start just behind a bin, so that queries up to that length stay inside
    their first bin; queries longer than the hint (reach = half the longest) must stay exact"""
    qs, qe, qoff, ss, se, soff, sg = _setup(case, empty_every=3 if case[0] == 3 else 0)
    q = eng.Queries(_dev(eng, qs, qe, qoff), sg, slack=slack)
    r = {"none": 0, "max": eng._max_len(q), "half": eng._max_len(q) // 2}[reach]
    ix = eng.ListIndex(_dev(eng, ss, se, soff), reach=r)
    a, b = ix.pairs(q)
    a, b = a.cpu().numpy(), b.cpu().numpy()
    assert np.all(a[1:] >= a[:-1]), "pairs come out in query order"
    wa, wb = canon_pairs(*oracle_pairs(qs, qe, qoff, sg, ss, se, soff, "all", slack))
    ga, gb = canon_pairs(a, b)
    assert np.array_equal(ga, wa) and np.array_equal(gb, wb)
    # deterministic: a second build + join gives the identical (not just set-equal) pair list
    a2, b2 = eng.ListIndex(_dev(eng, ss, se, soff), reach=r).pairs(q)
    assert np.array_equal(a2.cpu().numpy(), a) and np.array_equal(b2.cpu().numpy(), b)
    # counts, all / any
    c = ix.count(q, eng.IV_MODE_ALL).cpu().numpy()
    assert np.array_equal(c, oracle_counts(qs, qe, qoff, sg, ss, se, soff, "all", slack))
    c1 = ix.count(q, eng.IV_MODE_ANY).cpu().numpy()
    assert np.array_equal(c1, (c > 0).astype(np.int64))
    # per-key pair counts from the sorted pair list
    import torch

    tot = torch.tensor([len(a)], dtype=torch.int64, device="cuda")
    per_key = ix.segment_pair_counts(q, torch.from_numpy(a).cuda(), tot, len(a)).cpu().numpy()
    assert np.array_equal(per_key, [c[qoff[k]:qoff[k + 1]].sum() for k in range(len(qoff) - 1)])


def test_join_groups_of_segments(eng):
    """strand-ignoring join on stranded data: a partner group = both strand segments of a chromosome"""
    qs, qe, qoff = gen_segments(11, 6, 4000, 50000, 200)
    ss, se, soff = gen_segments(12, 6, 900, 50000, 400)
    ranges = [(0, 2), (2, 4), (4, 6)]
    sg = (np.arange(6) // 2).astype(np.int32)
    goff = soff[::2]
    ix = eng.ListIndex(_dev(eng, ss, se, soff), ranges)
    a, b = ix.pairs(eng.Queries(_dev(eng, qs, qe, qoff), sg))
    wa, wb = canon_pairs(*oracle_pairs(qs, qe, qoff, sg, ss, se, goff, "all"))
    ga, gb = canon_pairs(a.cpu().numpy(), b.cpu().numpy())
    assert np.array_equal(ga, wa) and np.array_equal(gb, wb)


def test_join_labels_and_capacity(eng):
    import torch

    qs, qe, qoff, ss, se, soff, sg = _setup(CASES[2])
    rng = np.random.default_rng(0)
    ql = rng.permutation(len(qs)).astype(np.int64) + 1000
    sl = rng.permutation(len(ss)).astype(np.int64) + 5_000_000_000
    ix = eng.ListIndex(_dev(eng, ss, se, soff))
    q = eng.Queries(_dev(eng, qs, qe, qoff), sg, label=torch.from_numpy(ql).cuda())
    a, b = ix.pairs(q, s_label=torch.from_numpy(sl).cuda())
    wa, wb = oracle_pairs(qs, qe, qoff, sg, ss, se, soff, "all")
    ea, eb = canon_pairs(ql[wa], sl[wb])
    ga, gb = canon_pairs(a.cpu().numpy(), b.cpu().numpy())
    assert np.array_equal(ga, ea) and np.array_equal(gb, eb)
    # a capacity that is too small: the total is still exact, the pairs below the capacity are the true prefix,
    # nothing is written beyond it
    p = len(wa)
    cap = p // 3
    q0 = eng.Queries(_dev(eng, qs, qe, qoff), sg)
    full_a, full_b = ix.pairs(q0)
    oq = torch.full((cap + 64,), -7, dtype=torch.int64, device="cuda")
    os_ = torch.full((cap + 64,), -7, dtype=torch.int64, device="cuda")
    total = ix.join_pairs(q0, oq, os_, cap)
    assert int(total.item()) == p
    assert torch.equal(oq[:cap], full_a[:cap]) and torch.equal(os_[:cap], full_b[:cap])
    assert bool((oq[cap:] == -7).all()) and bool((os_[cap:] == -7).all())
    # only one of the two outputs
    a_only, none = ix.pairs(q0, want_s=False)
    assert none is None and torch.equal(a_only, full_a)


def test_join_many_hits_per_query(eng):
    """> 2000 hits per query (the ncls buffer boundary of 1024, CHANGELOG.txt:108, is far behind) and more pairs per
    warp than the staging area holds: the direct-write path"""
    rng = np.random.default_rng(3)
    ns, nq = 2500, 5000
    ss = rng.integers(1000, 1100, ns).astype(np.int64)
    se = ss + rng.integers(1, 50, ns)
    qs = rng.integers(900, 1200, nq).astype(np.int64)
    qe = qs + rng.integers(1, 300, nq)
    off_q, off_s = np.asarray([0, nq]), np.asarray([0, ns])
    sg = np.zeros(1, np.int32)
    ix = eng.ListIndex(_dev(eng, ss, se, off_s))
    q = eng.Queries(_dev(eng, qs, qe, off_q), sg)
    a, b = ix.pairs(q)
    c = ix.count(q).cpu().numpy()
    assert c.max() > 2000
    wa, wb = brute_pairs(qs, qe, off_q, sg, ss, se, off_s)
    ga, gb = canon_pairs(a.cpu().numpy(), b.cpu().numpy())
    assert np.array_equal(ga, wa) and np.array_equal(gb, wb)
    assert np.array_equal(c, np.bincount(wa, minlength=nq))


def test_join_identical_subjects_pile_up(eng):
    """thousands of identical subjects in ONE bin: the block-sorted long lists (shared memory and global)"""
    for ns in (100, 3000, 9000):
        ss = np.full(ns, 5000, np.int64)
        se = np.full(ns, 5100, np.int64)
        ss[::7] += 3  # a few different starts inside the same bin
        qs = np.asarray([4000, 5000, 5002, 5099, 5100, 5050], np.int64)
        qe = np.asarray([5001, 5001, 5003, 5200, 5300, 5050], np.int64)
        off_q, off_s = np.asarray([0, len(qs)]), np.asarray([0, ns])
        sg = np.zeros(1, np.int32)
        # a wide second group keeps the bins coarse enough for everything to share one
        ix = eng.ListIndex(_dev(eng, ss, se, off_s))
        q = eng.Queries(_dev(eng, qs, qe, off_q), sg)
        a, b = ix.pairs(q)
        wa, wb = brute_pairs(qs, qe, off_q, sg, ss, se, off_s)
        ga, gb = canon_pairs(a.cpu().numpy(), b.cpu().numpy())
        assert np.array_equal(ga, wa) and np.array_equal(gb, wb)
        # canonical list order: rows ascending inside equal starts
        bb = b.cpu().numpy()[a.cpu().numpy() == 1]
        assert len(bb) == len(np.unique(bb))


def test_join_wide_and_negative_coordinates(eng):
    """flattened span beyond 2^32 (64-bit instantiation) and negative coordinates"""
    rng = np.random.default_rng(9)
    S, E, Q, QE, offs, offq = [], [], [], [], [0], [0]
    for base in (-3_000_000_000, 0, 5_000_000_000_000):
        s = base + rng.integers(0, 4_000_000_000, 3000)
        S.append(s)
        E.append(s + rng.integers(1, 100_000_000, 3000))
        q = base + rng.integers(-1000, 4_000_100_000, 4000)
        Q.append(q)
        QE.append(q + rng.integers(0, 50_000_000, 4000))
        offs.append(offs[-1] + 3000)
        offq.append(offq[-1] + 4000)
    ss, se, qs, qe = (np.concatenate(x).astype(np.int64) for x in (S, E, Q, QE))
    soff, qoff = np.asarray(offs), np.asarray(offq)
    sg = np.arange(3, dtype=np.int32)
    s_iv = _dev(eng, ss, se, soff)
    if not eng.ListIndex.supported(s_iv, None):
        pytest.skip("span beyond the list index")
    ix = eng.ListIndex(s_iv)
    q = eng.Queries(_dev(eng, qs, qe, qoff), sg)
    a, b = ix.pairs(q)
    wa, wb = brute_pairs(qs, qe, qoff, sg, ss, se, soff)
    ga, gb = canon_pairs(a.cpu().numpy(), b.cpu().numpy())
    assert np.array_equal(ga, wa) and np.array_equal(gb, wb)


def test_join_zero_length_intervals(eng):
    """Start == End on either side, at coincident coordinates, at the window edges and at bin origins: the
    reference predicate S < e and s < E decides (a zero-length subject is hit iff s < S < e)"""
    rng = np.random.default_rng(21)
    for maxpos in (64, 5000, 3_000_000):
        ns, nq = 1500, 4000
        ss = rng.integers(0, maxpos, ns).astype(np.int64)
        ln = rng.integers(0, 40, ns)
        ln[rng.random(ns) < 0.4] = 0
        se = ss + ln
        qs = rng.integers(0, maxpos, nq).astype(np.int64)
        qln = rng.integers(0, 60, nq)
        qln[rng.random(nq) < 0.3] = 0
        qe = qs + qln
        # coincident: queries sitting exactly on zero-length subjects, starting / ending at them
        z = ss[ln == 0][:300]
        qs[:300], qe[:300] = z, z
        qs[300:600], qe[300:600] = z, z + 5
        qs[600:900], qe[600:900] = np.maximum(z - 5, 0), z
        qs[900:1200], qe[900:1200] = np.maximum(z - 1, 0), z + 1
        # the extreme rows of the subjects are zero-length too (window edges)
        ss[0], se[0] = ss.min(), ss.min()
        se[1] = se.max() + 1
        ss[1] = se[1]
        off_q, off_s = np.asarray([0, nq // 2, nq]), np.asarray([0, ns // 2, ns])
        sg = np.asarray([0, 1], np.int32)
        s_iv = _dev(eng, ss, se, off_s)
        assert s_iv.has_zero_length
        q = eng.Queries(_dev(eng, qs, qe, off_q), sg)
        ix = eng.ListIndex(s_iv, reach=eng._max_len(q) if maxpos != 5000 else 0)
        a, b = ix.pairs(q)
        wa, wb = brute_pairs(qs, qe, off_q, sg, ss, se, off_s)
        ga, gb = canon_pairs(a.cpu().numpy(), b.cpu().numpy())
        assert np.array_equal(ga, wa) and np.array_equal(gb, wb)
        assert np.array_equal(ix.count(q).cpu().numpy(), np.bincount(wa, minlength=nq))
        with pytest.raises(ValueError):
            eng.SubjectIndex(s_iv)  # the rank index refuses them loudly


def test_zero_length_queries_rank_index(eng):
    """zero-length QUERIES against proper subjects are exact on the rank index too (count, pairs, first)"""
    rng = np.random.default_rng(22)
    ns, nq = 800, 3000
    ss = rng.integers(0, 4000, ns).astype(np.int64)
    se = ss + rng.integers(1, 200, ns)
    qs = rng.integers(0, 4200, nq).astype(np.int64)
    qln = rng.integers(0, 50, nq)
    qln[::2] = 0
    qe = qs + qln
    qs[:200], qe[:200] = ss[:200], ss[:200]
    qs[200:400], qe[200:400] = se[:200], se[:200]
    off_q, off_s = np.asarray([0, nq]), np.asarray([0, ns])
    sg = np.zeros(1, np.int32)
    ix = eng.SubjectIndex(_dev(eng, ss, se, off_s))
    q = eng.Queries(_dev(eng, qs, qe, off_q), sg)
    counts, a, b = ix.pairs(q, eng.IV_MODE_ALL)
    wa, wb = brute_pairs(qs, qe, off_q, sg, ss, se, off_s)
    ga, gb = canon_pairs(a.cpu().numpy(), b.cpu().numpy())
    assert np.array_equal(ga, wa) and np.array_equal(gb, wb)
    assert np.array_equal(counts.cpu().numpy(), np.bincount(wa, minlength=nq))


def test_join_step_protocol(eng):
    """JoinStep: synchronous sizing pass, asynchronous steps, device-side totals validated afterwards; both paths"""
    qs, qe, qoff, ss, se, soff, sg = _setup(CASES[4])
    q_iv, s_iv = _dev(eng, qs, qe, qoff), _dev(eng, ss, se, soff)
    want = canon_pairs(*oracle_pairs(qs, qe, qoff, sg, ss, se, soff, "all"))
    for path in ("fused", "twopass"):
        js = eng.JoinStep(s_iv, None, q_iv, sg, path=path)
        a, b, p = js.run_sync()
        assert p == len(want[0])
        for _ in range(3):
            oq, os_, total = js.step()
        assert js.check() == p
        ga, gb = canon_pairs(oq[:p].cpu().numpy(), os_[:p].cpu().numpy())
        assert np.array_equal(ga, want[0]) and np.array_equal(gb, want[1])
    cs = eng.CountStep(s_iv, None, q_iv, sg)
    assert np.array_equal(cs.step().cpu().numpy(), np.bincount(want[0], minlength=len(qs)))
    ov = eng.OverlapFirstStep(s_iv, None, q_iv, sg)
    rows = ov.run_sync().cpu().numpy()
    assert np.array_equal(rows, np.unique(want[0]))
    r2, t2 = ov.step()
    ov.check()
    assert np.array_equal(r2[:int(t2.item())].cpu().numpy(), rows)
