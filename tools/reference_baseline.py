"""The reference's OWN pyranges layer timed on the host cores (SURVEY.md 8(d) "CPU baseline
beside it"): the UNMODIFIED package from /root/reference over the oracle shims (oracle/shims -> oracle/iv_oracle.c,
the restated ncls / sorted_nearest / pyrle), PyRanges built through the dict path.

  python tools/reference_baseline.py [--reads N --annot M --procs P]

Prints one JSON line: gr.join(other) and gr.count_overlaps(other) end to end (pandas glue included) at nb_cpu = 1,
the same per-key tasks over a fork pool of P processes ("ray-equivalent": the reference's nb_cpu > 1 needs Ray, which
is not installable here -- pyranges/multithreaded.py:160-179 ships one task per key to Ray workers; the pool does the
same with fork), and the index-only time (NCLS build + all_overlaps_both per key, no pandas).  Only runs where
/root/reference exists (the authoring container); bench.py --impl reference calls it when it can and falls back to the
index-only port elsewhere.  Caveat printed with the numbers: the native layer under the reference is the restatement,
not the upstream ncls wheel."""
import argparse
import importlib.metadata as md
import json
import os
import sys
import time

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
REF = "/root/reference"


def available():
    return os.path.isdir(os.path.join(REF, "pyranges"))


def _import_reference():
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "oracle", "shims"))
    sys.path.insert(0, REF)
    _v = md.version
    md.version = lambda n: "0.1.4" if n == "pyranges" else _v(n)
    import pandas as pd

    pd.set_option("future.infer_string", False)
    import pyranges as pr

    return pr, pd


def _frames(pd, np, ka, names):
    """dict key -> DataFrame with original global row labels (pyranges/methods/init.py:35-45)"""
    dfs = {}
    for k in range(ka.n_keys):
        a, b = ka.seg_off[k], ka.seg_off[k + 1]
        if a == b:
            continue
        c, s = names[ka.chrom_ids[k // 2]], "+-"[k % 2]
        dfs[(c, s)] = pd.DataFrame({"Chromosome": np.full(b - a, c, dtype=object), "Start": ka.starts[a:b],
                                    "End": ka.ends[a:b], "Strand": np.full(b - a, s, dtype=object)},
                                   index=np.arange(a, b))
    return dfs


def _join_key(args):
    """one task of the fork pool: the reference's per-key function on one (self, partner) frame pair"""
    from pyranges.methods.join import _write_both

    scdf, ocdf, kwargs = args
    df = _write_both(scdf, ocdf, **kwargs)
    return 0 if df is None else len(df)


def run(reads_n, annot_n, procs, frac=1.0):
    import numpy as np

    pr, pd = _import_reference()
    from oracle import oracle as orc
    from pyranges_b200 import synthetic as S

    orc.build()
    reads = S.reads_by_chrom(reads_n, seed=1)
    annot = S.annotations_by_chrom(annot_n, seed=2)
    names = [c for c, _ in S.HG38]
    gr = pr.PyRanges(_frames(pd, np, reads, names))
    other = pr.PyRanges(_frames(pd, np, annot, names))
    out = {"reads": reads.n, "annotations": annot.n, "host_cores": os.cpu_count(),
           "native_layer": "oracle/iv_oracle.c restatement of ncls (the wheel is not installable here)"}
    t = time.perf_counter()
    j = gr.join(other, strandedness=None)
    dt = time.perf_counter() - t
    pairs = len(j)
    out["join_nb_cpu_1"] = {"seconds": dt, "pairs": pairs, "pairs_per_s": pairs / dt}
    t = time.perf_counter()
    c = gr.count_overlaps(other, strandedness=None)
    dt = time.perf_counter() - t
    out["count_overlaps_nb_cpu_1"] = {"seconds": dt, "queries_per_s": len(c) / dt}
    # index only: what the GPU kernels replace (NCLS build + query per key, both strands of the chromosome as partner)
    t = time.perf_counter()
    tot = 0
    for g in range(annot.n_keys // 2):
        a, b = annot.seg_off[2 * g], annot.seg_off[2 * g + 2]
        ncls = orc.NCLS(annot.starts[a:b], annot.ends[a:b], np.arange(a, b, dtype=np.int64))
        for k in (2 * g, 2 * g + 1):
            q0, q1 = reads.seg_off[k], reads.seg_off[k + 1]
            x, _ = ncls.all_overlaps_both(reads.starts[q0:q1], reads.ends[q0:q1], np.arange(q0, q1, dtype=np.int64))
            tot += len(x)
    dt = time.perf_counter() - t
    assert tot == pairs
    out["index_only_1_core"] = {"seconds": dt, "pairs_per_s": tot / dt}
    # ray-equivalent: one task per key over a fork pool
    if procs > 1:
        import multiprocessing as mp

        from pyranges.multithreaded import merge_dfs

        kwargs = {"strandedness": None, "how": None, "suffix": "_b", "report_overlap": False, "slack": 0,
                  "apply_strand_suffix": None, "preserve_order": None, "sparse": {"self": False, "other": False}}
        tasks = []
        for (c, s), df in gr.dfs.items():
            parts = [other.dfs[k] for k in ((c, "+"), (c, "-")) if k in other.dfs]
            odf = parts[0] if len(parts) == 1 else merge_dfs(parts[0], parts[1])
            tasks.append((df, odf, kwargs))
        t = time.perf_counter()
        with mp.get_context("fork").Pool(procs) as pool:
            got = sum(pool.map(_join_key, tasks, chunksize=1))
        dt = time.perf_counter() - t
        assert got == pairs
        out["join_fork_pool"] = {"processes": procs, "seconds": dt, "pairs_per_s": pairs / dt,
                                 "note": "per-key _write_both tasks; results counted, not shipped back (Ray would ship them)"}
    return out


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--reads", type=int, default=2_000_000)
    ap.add_argument("--annot", type=int, default=40_000)
    ap.add_argument("--procs", type=int, default=os.cpu_count() or 1)
    a = ap.parse_args()
    if not available():
        print(json.dumps({"unavailable": "/root/reference is not present on this machine"}))
        sys.exit(0)
    print(json.dumps(run(a.reads, a.annot, a.procs)))
