"""GPU: sharding parity on the real CUDA engine.  The chromosomes are split into N shards by the LPT rule (as
bench.py --gpus N / pyranges_b200.sharding do), every shard runs on THIS GPU one after the other, and the union of the
shard results must be exactly the unsharded result: pairs (as global row ids), per-key pair counts, counts, nearest
distances, merged rows, cluster counts, RLE runs.  (The collective itself -- one all-reduce of per-key counts -- is
covered by the gloo test on CPU and by the checksums bench.py prints at N = 1, 2, 4, 8.)"""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _global_ids(rows, seg_off, global_key):
    """row of a shard's concatenated layout -> (global key id << 32 | row inside the key)"""
    k = np.searchsorted(seg_off, rows, side="right") - 1
    return (global_key[k] << 32) | (rows - seg_off[k])


def _run(E, S, chroms, n_reads, n_annot):
    import torch

    reads = S.reads_by_chrom(n_reads, seed=1, min_len=30, max_len=400, chroms=chroms)
    annot = S.annotations_by_chrom(n_annot, seed=2, chroms=chroms)
    sg = (np.arange(reads.n_keys) // 2).astype(np.int32)
    q_iv = E.DeviceIntervals.from_numpy(reads.starts, reads.ends, reads.seg_off)
    s_iv = E.DeviceIntervals.from_numpy(annot.starts, annot.ends, annot.seg_off)
    groups = annot.chrom_ranges()
    js = E.JoinStep(s_iv, groups, q_iv, sg, seg_counts=True)
    a, b, p = js.run_sync()
    per_key = js.seg_counts.cpu().numpy()
    a, b = a.cpu().numpy(), b.cpu().numpy()
    pairs = np.stack([_global_ids(a, reads.seg_off, reads.global_key), _global_ids(b, annot.seg_off, annot.global_key)], 1)
    counts = E.CountStep(s_iv, groups, q_iv, sg).step().cpu().numpy()
    same = np.arange(reads.n_keys, dtype=np.int32)
    modes = np.asarray([E.IV_NEAREST_PREVIOUS if k % 2 == 0 else E.IV_NEAREST_NEXT for k in range(reads.n_keys)], np.int32)
    _, dist = E.SubjectIndex(s_iv, None, ends_perm=True).nearest(E.Queries(q_iv, same, seg_mode=modes))
    ms, me, mc, mg = (t.cpu().numpy() for t in E.merge(q_iv))
    _, cid, nc = E.cluster(q_iv)
    off, runs, vals = E.coverage_rle(q_iv)
    runs, vals = runs.cpu().numpy(), vals.cpu().numpy()
    out = {"pairs": pairs, "per_key": dict(zip(reads.global_key.tolist(), per_key.tolist())),
           "n_clusters": int(nc.item())}
    rows = _global_ids(np.arange(reads.n), reads.seg_off, reads.global_key)
    out["counts"] = np.stack([rows, counts], 1)
    out["dist"] = np.stack([rows, dist.cpu().numpy()], 1)
    out["merged"] = np.stack([reads.global_key[mg], ms, me, mc], 1)
    out["rle"] = {int(g): (runs[off[k]:off[k + 1]], vals[off[k]:off[k + 1]]) for k, g in enumerate(reads.global_key)}
    torch.cuda.synchronize()
    return out


def _canon(a):
    return a[np.lexsort(a.T[::-1])]


@pytest.mark.parametrize("world", [2, 8])
def test_sharded_equals_unsharded(world):
    from pyranges_b200 import engine as E
    from pyranges_b200 import synthetic as S

    n_reads, n_annot = 3_000_000, 60_000
    full = _run(E, S, None, n_reads, n_annot)
    parts = [_run(E, S, S.shard_chromosomes(world, r), n_reads, n_annot) for r in range(world)]
    for name in ("pairs", "counts", "dist", "merged"):
        got = np.concatenate([p[name] for p in parts])
        assert np.array_equal(_canon(got), _canon(full[name])), name
    per_key = {}
    for p in parts:
        assert not set(per_key) & set(p["per_key"])  # keys are disjoint over the ranks
        per_key.update(p["per_key"])
    assert per_key == full["per_key"]
    assert sum(p["n_clusters"] for p in parts) == full["n_clusters"]
    rle = {}
    for p in parts:
        rle.update(p["rle"])
    assert set(rle) == set(full["rle"])
    for g, (r, v) in full["rle"].items():
        assert np.array_equal(rle[g][0], r) and np.array_equal(rle[g][1], v), g
