"""pr.count_overlaps(grs, features): the multi-sample count matrix.

Mirrors pyranges/multioverlap.py:6-165.  The reference loops over the samples and, per sample, over the keys
(``features.apply_pair(gr, _count_overlaps)``, one NCLS per key and sample).  Here the feature columns go to the
device once; every sample is one batched index build + one count pass over all keys."""
import numpy as np

from . import multithreaded as M
from .pyranges_main import PyRanges, concat, fill_kwargs


def count_overlaps(grs, features=None, strandedness=None, how=None, nb_cpu=1):
    """One int64 column per entry of `grs` (dict name -> PyRanges) on a copy of `features`: the number of intervals
    of that sample overlapping each feature (how=None), contained ... ("containment").  features=None: the samples'
    own intervals cut at every Start / End (``concat(grs).split(between=True)``)."""
    kwargs = fill_kwargs({"as_pyranges": False, "nb_cpu": nb_cpu, "strandedness": strandedness, "how": how})
    names = list(grs.keys())
    if features is None:
        features = concat(list(grs.values())).split(between=True)
    mode = M.E.IV_MODE_ALL if not how else (M.E.IV_MODE_CONTAIN if how == "containment" else M.E.IV_MODE_ANY)
    device = M._device_of(kwargs)
    q = None
    columns = {}
    for name, gr in grs.items():
        plan = M._Plan(features, gr.drop(), kwargs["strandedness"], device, q_frames=q)
        q = plan.q  # the feature columns stay on the device for the next sample
        counts, _, _ = plan.index().count(plan.queries(), mode, for_emit=False)
        columns[name] = counts.cpu().numpy()  # keys without partner frame count 0
    dfs = {}
    if q is None:
        return features.copy()
    for k, key in enumerate(q.keys):
        df = q.dfs[k].copy(deep=False)
        a, b = q.seg_off[k], q.seg_off[k + 1]
        for name in names:
            df.insert(df.shape[1], name, columns[name][a:b].astype(np.int64, copy=False))
        dfs[key] = df.reset_index(drop=True)
    return PyRanges._wrap(dfs)
