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


def test_lpt_balance():
    from pyranges_b200 import sharding
    from pyranges_b200.synthetic import HG38

    w = [s for _, s in HG38]
    for n in (2, 4, 8):
        a = sharding.lpt_assign(w, n)
        load = np.bincount(a, weights=w, minlength=n)
        assert load.max() / (sum(w) / n) < 1.08, (n, load)
