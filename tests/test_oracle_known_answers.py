"""CPU: the oracle (oracle/iv_oracle.c) against the known answers the REFERENCE itself publishes for this
path -- literal values copied from its doctests / unit tests (cited per test) -- and against brute force.
oracle/make_golden.py additionally replays the same doctests through the unmodified reference Python layer
over the oracle (authoring container only)."""
import numpy as np
import pytest

from oracle import oracle as orc
from tests.golden_util import load_input
from tests.util import canon_pairs, gen_segments


def _ncls(S, E):
    return orc.NCLS(np.asarray(S), np.asarray(E), np.arange(len(S)))


def test_overlap_doctest():
    # pyranges/pyranges_main.py:3536-3605: gr chr1 [1,3) [1,3) [10,11); gr2 chr1 [2,3) [2,9)
    t = _ncls([2, 2], [3, 9])
    s, e, i = np.array([1, 1, 10]), np.array([3, 3, 11]), np.arange(3)
    assert t.has_overlaps(s, e, i).tolist() == [0, 1]                       # how="first": a, a
    assert t.all_overlaps_self(s, e, i).tolist() == [0, 0, 1, 1]            # how=None: a a a a
    assert t.all_containments_both(s, e, i)[0].tolist() == []               # [1,3) not inside [2,3) / [2,9)
    t2 = _ncls([1], [10])                                                    # chr2: [4,9) inside [1,10)
    assert t2.all_containments_both(np.array([4]), np.array([9]), np.array([0]))[0].tolist() == [0]


def test_join_doctest_f1_f2():
    # pyranges/pyranges_main.py:2328-2362: f1 x f2 (strand ignored) -> [5,7) x [6,7)
    f1s, f1e = np.array([3, 8, 5]), np.array([6, 9, 7])
    t = _ncls([1, 6], [2, 7])
    a, b = t.all_overlaps_both(f1s, f1e, np.arange(3))
    assert a.tolist() == [2] and b.tolist() == [1]
    # count_overlaps doctest :1338-1372 and coverage :1438-1474 -> [0, 0, 1], fraction 0.5
    assert t.count(f1s, f1e).tolist() == [0, 0, 1]
    assert (t.coverage(f1s, f1e, np.arange(3)) / (f1e - f1s)).tolist() == [0.0, 0.0, 0.5]


def test_cluster_merge_doctests():
    # pyranges/pyranges_main.py:1111-1184
    S, E = np.array([1, 2, 3, 9]), np.array([3, 3, 10, 12])
    assert orc.annotate_clusters(S, E, 0).tolist() == [1, 1, 1, 1]
    assert orc.annotate_clusters(S, E, -1).tolist() == [1, 1, 2, 2]
    cs, ce, cn = orc.find_clusters(S, E, 0)
    assert (cs.tolist(), ce.tolist(), cn.tolist()) == ([1], [12], [4])
    # cluster(by="Gene") :1139-1150: Gene 1, 2, 3, 3 -> clusters 1, 2, 3, 3 (counts 1, 1, 2, 2)
    G = np.array([1, 2, 3, 3])
    assert orc.cluster_by(S, E, G, 0).tolist() == [1, 2, 3, 3]
    assert orc.cluster_by(S, E, G, -1).tolist() == [1, 2, 3, 3]
    assert orc.cluster_by(S, E, np.zeros(4, np.int64), -1).tolist() == [1, 1, 2, 2]
    mb, ms, me, mn = orc.merge_by(S, E, G, 0)
    assert (mb.tolist(), ms.tolist(), me.tolist(), mn.tolist()) == ([1, 2, 3], [1, 2, 3], [3, 3, 12], [1, 1, 2])


def test_nearest_doctest_distances():
    # pyranges/pyranges_main.py:3258-3309: f1 [3,6) [8,9) [5,7) vs f2 [1,2) [6,7): Distance 1, 2, 0
    Ls = np.array([3, 8])                                   # the non-overlapping queries, sorted by Start
    pi, pd_ = orc.nearest_previous_nonoverlapping(Ls, np.array([2, 7]) - 1, np.array([0, 1]))
    ni, nd = orc.nearest_next_nonoverlapping(np.array([6, 9]) - 1, np.array([1, 6]), np.array([0, 1]))
    assert pd_.tolist() == [2, 2] and nd.tolist() == [1, -1]
    ri, rd = orc.nearest_nonoverlapping(pi, pd_, ni, nd)
    assert rd.tolist() == [1, 2] and ri.tolist() == [1, 1]


def test_to_rle_doctests():
    # pyranges/pyranges_main.py:5800-5872
    r, v = orc.coverage_rle(np.array([3, 8]), np.array([6, 9]))
    assert r.tolist() == [3, 3, 2, 1] and v.tolist() == [0, 1, 0, 1]
    r, v = orc.coverage_rle(np.array([5]), np.array([7]))
    assert r.tolist() == [5, 2] and v.tolist() == [0, 1]
    r, v = orc.coverage_rle(np.array([3, 8, 5]), np.array([6, 9, 7]), np.array([0.1, 5, 3.14]))
    assert r.tolist() == [3, 2, 1, 1, 1, 1] and np.allclose(v, [0.0, 0.1, 3.24, 3.14, 0.0, 5.0])


def test_chipseq_known_answers():
    # bundled chipseq x chipseq_background -> 3 overlapping pairs (brute force), 1 same-strand
    cs, bg = load_input("chipseq"), load_input("chipseq_background")
    pairs = []
    for c in sorted(set(cs.Chromosome)):
        a, b = cs[cs.Chromosome == c], bg[bg.Chromosome == c]
        if len(b) == 0:
            continue
        t = orc.NCLS(b.Start.values, b.End.values, b.index.values)
        x, y = t.all_overlaps_both(a.Start.values, a.End.values, a.index.values)
        pairs += list(zip(x.tolist(), y.tolist()))
    assert len(pairs) == 3
    assert sorted(int(cs.Start[i]) for i, _ in pairs) == sorted([226987592, 38747226, 26105515])
    assert sum(cs.Strand[i] == bg.Strand[j] for i, j in pairs) == 1


@pytest.mark.parametrize("seed", [1, 2, 3])
def test_oracle_vs_brute_force(seed):
    qs, qe, _ = gen_segments(seed, 1, 400, 3000, 200)
    ss, se, _ = gen_segments(seed + 50, 1, 300, 3000, 200)
    t = orc.NCLS(ss, se, np.arange(len(ss)))
    a, b = canon_pairs(*t.all_overlaps_both(qs, qe, np.arange(len(qs))))
    wa, wb = orc.brute_pairs(ss, se, np.arange(len(ss)), qs, qe, np.arange(len(qs)))
    assert np.array_equal(a, wa) and np.array_equal(b, wb)
    assert np.array_equal(t.count(qs, qe), orc.brute_count(ss, se, qs, qe))
    m = (ss[None, :] < qe[:, None]) & (qs[:, None] < se[None, :])
    cov = (np.minimum(qe[:, None], se[None, :]) - np.maximum(qs[:, None], ss[None, :])) * m
    assert np.array_equal(t.coverage(qs, qe, np.arange(len(qs))), cov.sum(1))
    # first_overlap_both = head of the nested-list traversal = min (Start, -End, position)
    fa, fb = t.first_overlap_both(qs, qe, np.arange(len(qs)))
    for i, j in zip(fa, fb):
        hits = np.flatnonzero(m[i])
        best = min(hits, key=lambda h: (ss[h], -se[h], h))
        assert j == best
