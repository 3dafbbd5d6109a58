"""Multi-GPU layout: chromosomes are independent units of every hot-path operation
(pyranges/multithreaded.py:222-294, :324-363 never look across keys), so they are assigned to ranks by a
size-balanced (LPT) rule and every rank runs the single-GPU engine on its own keys.  The only
communication is one all_gather of the per-key result counts (int64[n_keys]) from which every rank derives
the global output offsets of a final concat -- and the running cluster-id offsets of
PyRanges.cluster (pyranges/pyranges_main.py:1210-1223).  One process per GPU, torch.distributed (NCCL on
the GPUs, gloo in the CPU tests)."""
import numpy as np


def lpt_assign(weights, n_ranks):
    """Longest-processing-time greedy: -> rank of every unit; deterministic (ties by unit index)."""
    w = np.asarray(weights, dtype=np.float64)
    order = np.argsort(-w, kind="stable")
    load = np.zeros(n_ranks)
    out = np.zeros(len(w), dtype=np.int64)
    for u in order:
        r = int(np.argmin(load))
        out[u] = r
        load[r] += w[u]
    return out


def chromosome_of(key):
    return key[0] if isinstance(key, tuple) else key


def shard_keys(keys, weights, n_ranks):
    """keys: natsorted PyRanges keys; weights: per key (e.g. rows).  Both strands of a chromosome stay on one
    rank (strand-ignoring operations need both).  -> {chromosome: rank}"""
    chroms, cw = [], {}
    for k, w in zip(keys, weights):
        c = chromosome_of(k)
        if c not in cw:
            chroms.append(c)
            cw[c] = 0.0
        cw[c] += float(w)
    ranks = lpt_assign([cw[c] for c in chroms], n_ranks)
    return {c: int(r) for c, r in zip(chroms, ranks)}


def local_part(gr, assignment, rank):
    """the PyRanges-like shard of `gr` owned by `rank` (same class, only its chromosomes)."""
    dfs = {k: v for k, v in gr.dfs.items() if assignment.get(chromosome_of(k), 0) == rank}
    return type(gr)(dfs)


def gather_key_counts(all_keys, local_counts, group=None):
    """all_keys: the global natsorted key list (identical on every rank); local_counts: {key: rows} of the
    keys this rank produced.  -> (counts int64[n_keys], offsets int64[n_keys+1]) of the global concat."""
    import torch
    import torch.distributed as dist

    v = torch.zeros(len(all_keys), dtype=torch.int64)
    for i, k in enumerate(all_keys):
        v[i] = int(local_counts.get(k, 0))
    if dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1:
        dev = torch.device("cuda", torch.cuda.current_device()) if dist.get_backend(group) == "nccl" else torch.device("cpu")
        v = v.to(dev)
        dist.all_reduce(v, op=dist.ReduceOp.SUM, group=group)  # keys are disjoint across ranks: sum == gather
        v = v.cpu()
    counts = v.numpy()
    off = np.zeros(len(counts) + 1, np.int64)
    np.cumsum(counts, out=off[1:])
    return counts, off
