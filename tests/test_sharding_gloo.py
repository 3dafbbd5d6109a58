"""CPU, world_size 2 over gloo: the multi-GPU host logic (LPT key assignment, per-rank join of the local
chromosomes, all-gather of per-key counts -> global offsets) must reproduce the single-rank result.
The device engine is replaced by the oracle-backed stand-in (no GPU here)."""
import os
import sys

import numpy as np
import pytest
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _worker(rank, world, port, q):
    sys.path.insert(0, ROOT)
    import torch.distributed as dist

    from pyranges_b200 import multithreaded, sharding
    from pyranges_b200._natsort import natsorted
    from pyranges_b200.pyranges_main import PyRanges
    from tests import fake_engine
    from tests.golden_util import load_input

    multithreaded.E = fake_engine
    multithreaded._device_of = lambda kwargs: "cpu"
    dist.init_process_group("gloo", init_method="tcp://127.0.0.1:%d" % port, rank=rank, world_size=world)
    A, B = PyRanges(load_input("A")), PyRanges(load_input("B"))
    keys = A.keys()
    assign = sharding.shard_keys(keys, [len(A.dfs[k]) for k in keys], world)
    la, lb = sharding.local_part(A, assign, rank), sharding.local_part(B, assign, rank)
    j = la.join(lb)
    c = la.cluster()
    all_keys = natsorted(set(A.keys()))
    counts, off = sharding.gather_key_counts(all_keys, {k: len(v) for k, v in j.dfs.items()})
    ccounts, coff = sharding.gather_key_counts(all_keys, {k: int(v.Cluster.max() - v.Cluster.min() + 1) for k, v in c.dfs.items()})
    # the rank's cluster ids made global with the gathered per-key cluster counts (pyranges_main.py:1210-1223)
    first = {k: int(coff[i]) for i, k in enumerate(all_keys)}
    gl = {k: (v.Cluster.values - int(v.Cluster.min()) + 1 + first[k]).tolist() for k, v in c.dfs.items()}
    q.put((rank, {k: len(v) for k, v in j.dfs.items()}, counts.tolist(), off.tolist(), ccounts.tolist(),
           sorted(set(assign.values())), dict(j.dfs), gl))
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_join_matches_single_rank(monkeypatch):
    from pyranges_b200 import multithreaded
    from pyranges_b200._natsort import natsorted
    from pyranges_b200.pyranges_main import PyRanges
    from tests import fake_engine
    from tests.golden_util import load_input

    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + (os.getpid() % 2000)
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=180) for _ in range(2)]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    monkeypatch.setattr(multithreaded, "E", fake_engine)
    monkeypatch.setattr(multithreaded, "_device_of", lambda kwargs: "cpu")
    A, B = PyRanges(load_input("A")), PyRanges(load_input("B"))
    full = A.join(B)
    keys = natsorted(set(A.keys()))
    want = [len(full.dfs[k]) if k in full.dfs else 0 for k in keys]
    cl = A.cluster()
    want_c = [int(cl.dfs[k].Cluster.max() - cl.dfs[k].Cluster.min() + 1) if k in cl.dfs else 0 for k in keys]
    from tests.golden_util import assert_frames_equal

    local, frames, gids = {}, {}, {}
    for rank, lens, counts, off, ccounts, ranks_used, jdfs, gl in res:
        assert counts == want  # every rank sees the global per-key counts
        assert off == np.concatenate([[0], np.cumsum(want)]).tolist()
        assert ccounts == want_c
        assert ranks_used == [0, 1]
        for k, n in lens.items():
            assert k not in local  # keys are disjoint across ranks
            local[k] = n
        frames.update(jdfs)
        gids.update(gl)
    assert local == {k: len(v) for k, v in full.dfs.items()}
    # the shards' results ARE the single-rank result, frame by frame, and the offset cluster ids are the global ones
    assert set(frames) == set(full.dfs)
    for k, v in full.dfs.items():
        assert_frames_equal(frames[k], v, str(k))
    for k, v in cl.dfs.items():
        assert gids[k] == v.Cluster.values.tolist(), k


OPS = {
    "count_overlaps": lambda a, b: a.count_overlaps(b),
    "overlap": lambda a, b: a.overlap(b),
    "nearest": lambda a, b: a.nearest(b),
    "k_nearest": lambda a, b: a.k_nearest(b, k=2, ties="different"),
    "merge": lambda a, b: a.merge(),
    "subtract": lambda a, b: a.subtract(b),
    "coverage": lambda a, b: a.coverage(b),
}


def _ops_worker(rank, world, port, q):
    sys.path.insert(0, ROOT)
    import torch.distributed as dist

    from pyranges_b200 import multithreaded, sharding
    from pyranges_b200.pyranges_main import PyRanges
    from tests import fake_engine
    from tests.golden_util import load_input

    multithreaded.E = fake_engine
    multithreaded._device_of = lambda kwargs: "cpu"
    dist.init_process_group("gloo", init_method="tcp://127.0.0.1:%d" % port, rank=rank, world_size=world)
    A, B = PyRanges(load_input("A")), PyRanges(load_input("B"))
    keys = A.keys()
    assign = sharding.shard_keys(keys, [len(A.dfs[k]) for k in keys], world)
    la, lb = sharding.local_part(A, assign, rank), sharding.local_part(B, assign, rank)
    out = {name: dict(f(la, lb).dfs) for name, f in OPS.items()}
    rle = la.to_rle()
    out["to_rle"] = {k: (np.asarray(v.runs).tolist(), np.asarray(v.values).tolist()) for k, v in rle.rles.items()}
    # per-key row counts of every operation, summed over the ranks: the offsets of the concatenated results
    all_keys = sorted(set(map(str, A.keys())))
    n_rows = {name: sharding.gather_key_counts(all_keys, {str(k): len(v) for k, v in d.items()} if name != "to_rle"
                                               else {str(k): len(v[0]) for k, v in d.items()})[0].tolist()
              for name, d in out.items()}
    q.put((rank, out, n_rows))
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_other_operations_match_single_rank(monkeypatch):
    """count_overlaps / overlap / nearest / k_nearest / merge / subtract / coverage / to_rle on two ranks: every key's
    frame (or run list) is the single-rank one, and the gathered per-key row counts are the global ones on both ranks"""
    from pyranges_b200 import multithreaded
    from pyranges_b200.pyranges_main import PyRanges
    from tests import fake_engine
    from tests.golden_util import assert_frames_equal, load_input

    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 31500 + (os.getpid() % 2000)
    procs = [ctx.Process(target=_ops_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=240) for _ in range(2)]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    monkeypatch.setattr(multithreaded, "E", fake_engine)
    monkeypatch.setattr(multithreaded, "_device_of", lambda kwargs: "cpu")
    A, B = PyRanges(load_input("A")), PyRanges(load_input("B"))
    all_keys = sorted(set(map(str, A.keys())))
    for name, f in OPS.items():
        full = f(A, B).dfs
        got = {}
        for _, out, _ in res:
            assert not (set(out[name]) & set(got)), name  # keys are disjoint across ranks
            got.update(out[name])
        assert set(got) == set(full), name
        for k, v in full.items():
            assert_frames_equal(got[k], v, "%s %s" % (name, k))
        want_rows = [sum(len(v) for k, v in full.items() if str(k) == ks) for ks in all_keys]
        for _, _, n_rows in res:
            assert n_rows[name] == want_rows, name
    full = A.to_rle().rles
    got = {}
    for _, out, _ in res:
        got.update(out["to_rle"])
    assert set(got) == set(full)
    for k, v in full.items():
        assert got[k] == (np.asarray(v.runs).tolist(), np.asarray(v.values).tolist()), k


def test_lpt_balance():
    from pyranges_b200 import sharding
    from pyranges_b200.synthetic import HG38

    w = [s for _, s in HG38]
    for n in (2, 4, 8):
        a = sharding.lpt_assign(w, n)
        load = np.bincount(a, weights=w, minlength=n)
        assert load.max() / (sum(w) / n) < 1.08, (n, load)
