"""GPU, BASELINE.json config sizes (10 M reads x 200 k annotations): the oracle is too slow to replay them in a
test, so parity is checked through size-independent properties + an oracle check of a random sample of keys."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def c2():
    import torch

    from pyranges_b200 import engine as E
    from pyranges_b200 import synthetic as S

    reads, annot = S.reads(10_000_000, seed=1), S.annotations(200_000, seed=2)
    sg = (np.arange(reads.n_keys) // 2).astype(np.int32)
    q_iv = E.DeviceIntervals.from_numpy(reads.starts, reads.ends, reads.seg_off)
    s_iv = E.DeviceIntervals.from_numpy(annot.starts, annot.ends, annot.seg_off)
    ix = E.SubjectIndex(s_iv, annot.chrom_ranges(), ends_perm=True)
    return E, torch, reads, annot, sg, q_iv, s_iv, ix


def test_join_properties_full_size(c2):
    E, torch, reads, annot, sg, q_iv, s_iv, ix = c2
    q = E.Queries(q_iv, sg)
    counts, a, b = ix.pairs(q, E.IV_MODE_ALL)
    p = a.numel()
    assert int(counts.sum().item()) == p and p > 1_000_000
    assert bool((a[1:] >= a[:-1]).all()), "pairs come out in query order"
    # every pair is a true overlap inside the right chromosome group, no pair is emitted twice
    qs, qe, ss, se = q_iv.start[a], q_iv.end[a], s_iv.start[b], s_iv.end[b]
    assert bool(((ss < qe) & (qs < se)).all())
    grp_off = torch.from_numpy(annot.seg_off[::2].copy()).cuda()  # chromosome groups = pairs of strand keys
    key_of_q = torch.searchsorted(q_iv.seg_off_dev, a, right=True) - 1
    grp_of_s = torch.searchsorted(grp_off, b, right=True) - 1
    assert bool((grp_of_s == key_of_q // 2).all())
    packed = a * annot.n + b
    assert int(torch.unique(packed).numel()) == p
    # counts = bincount of the emitted query rows; ANY = counts > 0
    assert bool((torch.bincount(a, minlength=reads.n) == counts).all())
    any_c, _, _ = ix.count(q, E.IV_MODE_ANY, for_emit=False)
    assert bool((any_c == (counts > 0).long()).all())


def test_counts_vs_oracle_on_sampled_keys(c2):
    from oracle import oracle as orc

    E, torch, reads, annot, sg, q_iv, s_iv, ix = c2
    counts, _, _ = ix.count(E.Queries(q_iv, sg), E.IV_MODE_ALL, for_emit=False)
    counts = counts.cpu().numpy()
    for k in (0, 17, 47):  # chr1 +, chr9 -, chrY -
        g = k // 2
        a, b = annot.seg_off[2 * g], annot.seg_off[2 * g + 2]
        t = orc.NCLS(annot.starts[a:b], annot.ends[a:b], np.arange(a, b))
        sl = slice(reads.seg_off[k], reads.seg_off[k + 1])
        assert np.array_equal(counts[sl], t.count(reads.starts[sl], reads.ends[sl]))


def test_nearest_properties_full_size(c2):
    E, torch, reads, annot, sg, q_iv, s_iv, ix = c2
    q = E.Queries(q_iv, sg)
    row, dist = ix.nearest(q, overlap=True)
    counts, _, _ = ix.count(q, E.IV_MODE_ALL, for_emit=False)
    assert bool(((dist == 0) == (counts > 0)).all())
    ok = row >= 0
    assert bool(ok.all())  # every chromosome has annotations
    S, En, s, e = s_iv.start[row], s_iv.end[row], q_iv.start, q_iv.end
    gap = torch.where(En <= s, s - En + 1, torch.where(S >= e, S - e + 1, torch.zeros_like(s)))
    assert bool((gap == dist).all())


def test_cluster_merge_rle_properties_full_size(c2):
    E, torch, reads, annot, sg, q_iv, s_iv, ix = c2
    row, cid, nc = E.cluster(q_iv)
    ms, me, mc, mg = E.merge(q_iv)
    m = int(nc.item())
    assert ms.numel() == m and int(mc.sum().item()) == reads.n
    assert bool((cid[1:] >= cid[:-1]).all()) and int(cid[0].item()) == 1 and int(cid[-1].item()) == m
    assert bool((torch.sort(row).values == torch.arange(reads.n, device=row.device)).all())
    # idempotence: merging the merged intervals changes nothing
    off2 = np.concatenate([[0], np.cumsum(np.bincount(mg.cpu().numpy(), minlength=reads.n_keys))])
    again = E.merge(E.DeviceIntervals(ms, me, off2))
    assert bool((again[0] == ms).all()) and bool((again[1] == me).all()) and bool((again[2] == 1).all())
    # coverage: runs of a key sum to its largest End; sum(run * value) = covered bases = sum of lengths
    run_off, runs, vals = E.coverage_rle(q_iv)
    hi = q_iv.hi
    for k in (0, 5, 47):
        sl = slice(int(run_off[k]), int(run_off[k + 1]))
        assert int(runs[sl].sum().item()) == int(hi[k])
        a, b = reads.seg_off[k], reads.seg_off[k + 1]
        assert float((runs[sl].double() * vals[sl]).sum().item()) == float((reads.ends[a:b] - reads.starts[a:b]).sum())
        assert bool((vals[sl][1:] != vals[sl][:-1]).all())  # equal neighbours are coalesced


def test_c3_count_overlap_join_100m_x_1m():
    """BASELINE.json config 3 / the north-star size: 100 M reads x 1 M annotations.  count_overlaps against the oracle
    on sampled keys, overlap (ANY) = counts > 0, and the join's pair list: total = sum of counts, query order, every
    pair a true overlap of the right chromosome, no duplicates, per-query multiplicity = counts."""
    import torch

    from oracle import oracle as orc
    from pyranges_b200 import engine as E
    from pyranges_b200 import synthetic as S

    reads, annot = S.reads(100_000_000, seed=1), S.annotations(1_000_000, seed=2)
    sg = (np.arange(reads.n_keys) // 2).astype(np.int32)
    q_iv = E.DeviceIntervals.from_numpy(reads.starts, reads.ends, reads.seg_off)
    s_iv = E.DeviceIntervals.from_numpy(annot.starts, annot.ends, annot.seg_off)
    ix = E.SubjectIndex(s_iv, annot.chrom_ranges())
    q = E.Queries(q_iv, sg)
    counts, a, b = ix.pairs(q, E.IV_MODE_ALL)
    p = a.numel()
    assert int(counts.sum().item()) == p and p > 100_000_000
    assert bool((a[1:] >= a[:-1]).all())
    assert bool(((s_iv.start[b] < q_iv.end[a]) & (q_iv.start[a] < s_iv.end[b])).all())
    grp_off = torch.from_numpy(annot.seg_off[::2].copy()).cuda()
    key_of_q = torch.searchsorted(q_iv.seg_off_dev, a, right=True) - 1
    assert bool(((torch.searchsorted(grp_off, b, right=True) - 1) == key_of_q // 2).all())
    del key_of_q
    assert bool((torch.bincount(a, minlength=reads.n) == counts).all())
    # no pair twice: inside a query (pairs are grouped by query) the subject rows are distinct
    o = torch.argsort(a * annot.n + b)
    assert bool(((a[o][1:] != a[o][:-1]) | (b[o][1:] != b[o][:-1])).all())
    del o, a, b
    any_c, _, _ = ix.count(q, E.IV_MODE_ANY, for_emit=False)
    assert bool((any_c == (counts > 0).long()).all())
    hc = counts.cpu().numpy()
    for k in (46, 47, 42):  # chrY +, chrY -, chr22 +
        g = k // 2
        sa, sb = annot.seg_off[2 * g], annot.seg_off[2 * g + 2]
        t = orc.NCLS(annot.starts[sa:sb], annot.ends[sa:sb], np.arange(sa, sb))
        sl = slice(reads.seg_off[k], reads.seg_off[k + 1])
        assert np.array_equal(hc[sl], t.count(reads.starts[sl], reads.ends[sl])), k


def test_c4_nearest_cluster_merge_50m():
    """BASELINE.json config 4: nearest (strandedness='same', how='upstream') of 50 M intervals against 1 M annotations
    and cluster / merge of the 50 M intervals, unsorted and sorted input."""
    import torch

    from pyranges_b200 import engine as E
    from pyranges_b200 import synthetic as S

    iv, ann = S.reads(50_000_000, seed=3, min_len=50, max_len=500), S.annotations(1_000_000, seed=2)
    q_iv = E.DeviceIntervals.from_numpy(iv.starts, iv.ends, iv.seg_off)
    s_iv = E.DeviceIntervals.from_numpy(ann.starts, ann.ends, ann.seg_off)
    sg = np.arange(iv.n_keys, dtype=np.int32)  # same strand: every key against its own key
    modes = np.asarray([E.IV_NEAREST_PREVIOUS if k % 2 == 0 else E.IV_NEAREST_NEXT for k in range(iv.n_keys)], np.int32)
    ix = E.SubjectIndex(s_iv, None, ends_perm=True)
    q = E.Queries(q_iv, sg, seg_mode=modes)
    row, dist = ix.nearest(q, overlap=True)
    counts, _, _ = ix.count(E.Queries(q_iv, sg), E.IV_MODE_ALL, for_emit=False)
    assert bool(((dist == 0) == (counts > 0)).all())
    found = row >= 0
    S_, E_ = s_iv.start[row.clamp(min=0)], s_iv.end[row.clamp(min=0)]
    s, e = q_iv.start, q_iv.end
    gap = torch.where(E_ <= s, s - E_ + 1, torch.where(S_ >= e, S_ - e + 1, torch.zeros_like(s)))
    assert bool((gap[found] == dist[found]).all())
    # direction: '+' keys look upstream = towards smaller coordinates, '-' keys towards larger ones
    plus = (torch.searchsorted(q_iv.seg_off_dev, torch.arange(iv.n, device=row.device), right=True) - 1) % 2 == 0
    far = found & (dist > 0)
    assert bool((E_[far & plus] <= s[far & plus]).all()) and bool((S_[far & ~plus] >= e[far & ~plus]).all())
    # the partner sits in the query's own key
    key_q = torch.searchsorted(q_iv.seg_off_dev, torch.arange(iv.n, device=row.device), right=True) - 1
    key_s = torch.searchsorted(s_iv.seg_off_dev, row.clamp(min=0), right=True) - 1
    assert bool((key_q[found] == key_s[found]).all())
    del S_, E_, gap, plus, far, key_q, key_s, row, dist, counts
    # cluster / merge: unsorted input, then the same rows sorted (the sort is skipped: same result)
    r1, c1, n1 = E.cluster(q_iv)
    ms, me, mc, mg = E.merge(q_iv)
    m = int(n1.item())
    assert ms.numel() == m and int(mc.sum().item()) == iv.n
    assert bool((c1[1:] >= c1[:-1]).all()) and int(c1[-1].item()) == m
    o = np.lexsort([iv.ends, iv.starts, np.repeat(np.arange(iv.n_keys), np.diff(iv.seg_off))])
    qs_iv = E.DeviceIntervals.from_numpy(iv.starts[o], iv.ends[o], iv.seg_off)
    ms2, me2, mc2, mg2 = E.merge(qs_iv)
    assert bool((ms2 == ms).all()) and bool((me2 == me).all()) and bool((mc2 == mc).all()) and bool((mg2 == mg).all())
    r2, c2_, n2 = E.cluster(qs_iv)
    assert int(n2.item()) == m and bool((c2_ == c1).all())


@pytest.mark.parametrize("n,bits,vdtype", [(10_000_000, 40, "int32"), (20_000_000, 33, None), (9_000_000, 64, "int64")])
def test_radix_sort_large_vs_torch(n, bits, vdtype):
    """large-input path of the radix sort (4096-key tiles, block-per-row histogram prefix) against torch's stable sort"""
    import torch

    from pyranges_b200 import engine as E

    g = torch.Generator(device="cuda").manual_seed(n % 1000)
    hi = 2 ** min(bits, 62)
    k = torch.randint(0, hi, (n,), dtype=torch.int64, device="cuda", generator=g)
    if bits == 64:
        k = k | (torch.randint(0, 2, (n,), dtype=torch.int64, device="cuda", generator=g) << 63)  # top bit: unsigned order
    m = k[3::7].numel()
    k[::7][:m] = k[3::7]  # duplicates: stability matters
    want_k, want_o = torch.sort(k if bits < 64 else (k ^ (1 << 63)), stable=True)
    if bits == 64:
        want_k = want_k ^ (1 << 63)
    v = None if vdtype is None else torch.arange(n, dtype=getattr(torch, vdtype), device="cuda")
    kk = k.clone()
    E.radix_sort_u64(kk, v, 0, bits)
    assert bool((kk == want_k).all())
    if v is not None:
        assert bool((v.long() == want_o).all())


def test_c3_fused_join_equals_rank_join_100m_x_1m():
    """The north-star size through BOTH index families: the list index + fused single-pass join must produce exactly
    the pair set (and per-row counts) of the rank index + count / emit pair; counts against the oracle on sampled keys."""
    import torch

    from oracle import oracle as orc
    from pyranges_b200 import engine as E
    from pyranges_b200 import synthetic as S

    reads, annot = S.reads_by_chrom(100_000_000, seed=1), S.annotations_by_chrom(1_000_000, seed=2)
    sg = (np.arange(reads.n_keys) // 2).astype(np.int32)
    q_iv = E.DeviceIntervals.from_numpy(reads.starts, reads.ends, reads.seg_off)
    s_iv = E.DeviceIntervals.from_numpy(annot.starts, annot.ends, annot.seg_off)
    groups = annot.chrom_ranges()
    q = E.Queries(q_iv, sg)
    lx = E.ListIndex(s_iv, groups, reach=E._max_len(q))
    a, b = lx.pairs(q)
    p = a.numel()
    assert p > 100_000_000 and bool((a[1:] >= a[:-1]).all())
    counts = lx.count(q)
    assert int(counts.sum().item()) == p
    assert bool((torch.bincount(a, minlength=reads.n) == counts).all())
    key_fused = torch.sort(a * annot.n + b).values
    del a, b
    c2, a2, b2 = E.SubjectIndex(s_iv, groups).pairs(q, E.IV_MODE_ALL)
    assert bool((c2 == counts).all())
    key_rank = torch.sort(a2 * annot.n + b2).values
    del a2, b2, c2
    assert key_fused.numel() == key_rank.numel() and bool((key_fused == key_rank).all())
    del key_fused, key_rank
    rows = torch.empty(reads.n, dtype=torch.int64, device="cuda")
    k = int(lx.join_pairs(q, rows, None, reads.n, mode=E.IV_MODE_ANY).item())
    assert bool((rows[:k] == torch.nonzero(counts > 0).flatten()).all())
    hc = counts.cpu().numpy()
    for key in (46, 47, 40):  # chrY +, chrY -, chr21 +
        g = key // 2
        sa, sb = annot.seg_off[2 * g], annot.seg_off[2 * g + 2]
        t = orc.NCLS(annot.starts[sa:sb], annot.ends[sa:sb], np.arange(sa, sb))
        sl = slice(reads.seg_off[key], reads.seg_off[key + 1])
        assert np.array_equal(hc[sl], t.count(reads.starts[sl], reads.ends[sl])), key


def test_count_beyond_2_to_31_pairs():
    """more than 2^31 overlap pairs in total (count-only: the pairs themselves would take 40 GB): int64 counts"""
    import torch

    from pyranges_b200 import engine as E

    rng = np.random.default_rng(31)
    ns, nq = 50, 60_000_000
    ss = rng.integers(0, 1000, ns).astype(np.int64)
    se = ss + 1_000_000
    qs = rng.integers(2000, 900_000, nq).astype(np.int64)
    qe = qs + 100
    s_iv = E.DeviceIntervals.from_numpy(ss, se, np.asarray([0, ns]))
    q = E.Queries(E.DeviceIntervals.from_numpy(qs, qe, np.asarray([0, nq])), np.zeros(1, np.int32))
    for counts in (E.ListIndex(s_iv).count(q), E.SubjectIndex(s_iv).count(q, E.IV_MODE_ALL, for_emit=False)[0]):
        assert bool((counts == ns).all())
        assert int(counts.sum().item()) == ns * nq > 2 ** 31
    _, ws, total = E.SubjectIndex(s_iv).count(q, E.IV_MODE_ALL, for_emit=True, want_counts=False)
    assert int(total.item()) == ns * nq


def _sample_keys():
    return (46, 47, 40, 41)  # chrY +, chrY -, chr21 +, chr21 -


def test_c4_nearest_cluster_merge_vs_oracle_sampled_keys():
    """config 4 at full size, the oracle on sampled keys: nearest upstream / same strand must return the MINIMAL
    distance (not just a consistent one), cluster ids and merged rows must equal sorted_nearest's restatement."""
    import torch

    from oracle import oracle as orc
    from pyranges_b200 import engine as E
    from pyranges_b200 import synthetic as S

    iv, ann = S.reads_by_chrom(50_000_000, seed=3, min_len=50, max_len=500), S.annotations_by_chrom(1_000_000, seed=2)
    q_iv = E.DeviceIntervals.from_numpy(iv.starts, iv.ends, iv.seg_off)
    s_iv = E.DeviceIntervals.from_numpy(ann.starts, ann.ends, ann.seg_off)
    sg = np.arange(iv.n_keys, dtype=np.int32)
    modes = np.asarray([E.IV_NEAREST_PREVIOUS if k % 2 == 0 else E.IV_NEAREST_NEXT for k in range(iv.n_keys)], np.int32)
    row, dist = E.SubjectIndex(s_iv, None, ends_perm=True).nearest(E.Queries(q_iv, sg, seg_mode=modes), overlap=True)
    row, dist = row.cpu().numpy(), dist.cpu().numpy()
    for k in _sample_keys():
        a, b = iv.seg_off[k], iv.seg_off[k + 1]
        sa, sb = ann.seg_off[k], ann.seg_off[k + 1]
        s, e, S_, E_ = iv.starts[a:b], iv.ends[a:b], ann.starts[sa:sb], ann.ends[sa:sb]
        t = orc.NCLS(S_, E_, np.arange(sb - sa, dtype=np.int64))
        hit = t.count(s, e) > 0
        want = np.full(b - a, -1, np.int64)
        want[hit] = 0
        rest = np.flatnonzero(~hit)
        if k % 2 == 0:  # '+': upstream = previous
            o_e = np.argsort(E_, kind="stable")
            ls = np.argsort(s[rest], kind="stable")
            _, d = orc.nearest_previous_nonoverlapping(s[rest][ls], E_[o_e] - 1, o_e)
            want[rest[ls]] = d
        else:  # '-': upstream = next
            o_s = np.argsort(S_, kind="stable")
            le = np.argsort(e[rest], kind="stable")
            _, d = orc.nearest_next_nonoverlapping(e[rest][le] - 1, S_[o_s], o_s)
            want[rest[le]] = d
        assert np.array_equal(dist[a:b], want), k
        # the returned subject realises the distance
        got = row[a:b]
        ok = got >= 0
        assert np.array_equal(ok, want >= 0)
        Sg, Eg = ann.starts[got[ok]], ann.ends[got[ok]]
        gap = np.where(Eg <= s[ok], s[ok] - Eg + 1, np.where(Sg >= e[ok], Sg - e[ok] + 1, 0))
        assert np.array_equal(gap, want[ok]), k
    del row, dist
    r, cid, nc = E.cluster(q_iv)
    ms, me, mc, mg = (x.cpu().numpy() for x in E.merge(q_iv))
    r, cid = r.cpu().numpy(), cid.cpu().numpy()
    m_off = np.concatenate([[0], np.cumsum(np.bincount(mg, minlength=iv.n_keys))])
    for k in _sample_keys():
        a, b = iv.seg_off[k], iv.seg_off[k + 1]
        s, e = iv.starts[a:b], iv.ends[a:b]
        o = np.lexsort([np.arange(b - a), e, s])
        want_id = orc.annotate_clusters(s[o], e[o], 0)
        assert np.array_equal(r[a:b] - a, o), k  # rows come back sorted by (Start, End, row)
        first = cid[a] - 1  # ids are global over keys: this key's ids start after the earlier keys' clusters
        assert first == m_off[k]
        assert np.array_equal(cid[a:b] - first, want_id), k
        cs, ce, cn = orc.find_clusters(s[o], e[o], 0)
        sl = slice(m_off[k], m_off[k + 1])
        assert np.array_equal(ms[sl], cs) and np.array_equal(me[sl], ce) and np.array_equal(mc[sl], cn), k


def test_c5_to_rle_200m_vs_oracle_sampled_keys():
    """config 5: coverage RLE of 200 M reads; the runs / values of sampled keys against pyrle's restatement, global
    invariants for all keys"""
    from oracle import oracle as orc
    from pyranges_b200 import engine as E
    from pyranges_b200 import synthetic as S

    iv = S.reads_by_chrom(200_000_000, seed=4)
    r_iv = E.DeviceIntervals.from_numpy(iv.starts, iv.ends, iv.seg_off)
    off, runs, vals = E.coverage_rle(r_iv)
    hi = r_iv.hi
    for k in range(iv.n_keys):
        sl = slice(int(off[k]), int(off[k + 1]))
        assert int(runs[sl].sum().item()) == int(hi[k])
    assert float((runs.double() * vals).sum().item()) == float(100 * iv.n)  # covered bases = sum of lengths
    for k in _sample_keys():
        a, b = iv.seg_off[k], iv.seg_off[k + 1]
        wr, wv = orc.coverage_rle(iv.starts[a:b], iv.ends[a:b])
        sl = slice(int(off[k]), int(off[k + 1]))
        assert np.array_equal(runs[sl].cpu().numpy(), wr) and np.array_equal(vals[sl].cpu().numpy(), wv), k
