"""CPU: the seam-B1 patch PRINTED in INTEGRATION.md, executed against the UNMODIFIED reference package.

The code block is cut out of INTEGRATION.md and exec'd inside pyranges.multithreaded (what a maintainer would paste at
the bottom of that file); the reference's own PyRanges methods then reach pyranges_b200.multithreaded.pyrange_apply /
pyrange_apply_single with the reference's function OBJECTS (pyranges/multithreaded.py:182,306) and kwargs, and the
golden cases must come out as the unpatched reference produced them.  The device engine is the oracle-backed stand-in
(tests/fake_engine.py) -- this test is about the wiring.  Skipped where /root/reference does not exist (the GPU box)."""
import importlib.metadata as md
import os
import re
import sys

import numpy as np
import pytest

from tests import fake_engine
from tests.golden_util import assert_frames_equal, case_names, load_case, load_input

REF = "/root/reference"
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
pytestmark = pytest.mark.skipif(not os.path.isdir(os.path.join(REF, "pyranges")), reason="reference tree not present")


@pytest.fixture(scope="module")
def patched_reference():
    for p in (os.path.join(ROOT, "oracle", "shims"), REF):
        if p not in sys.path:
            sys.path.insert(0, p)
    _v = md.version
    md.version = lambda n: "0.1.4" if n == "pyranges" else _v(n)
    import pandas as pd

    pd.set_option("future.infer_string", False)
    import pyranges as pr
    import pyranges.multithreaded as ref_mt
    import pyranges.pyranges_main as ref_main

    from pyranges_b200 import multithreaded as b200_mt

    text = open(os.path.join(ROOT, "INTEGRATION.md")).read()
    block = re.search(r"```python\n# pyranges/multithreaded.py \(reference\).*?\n(.*?)```", text, re.S).group(1)
    saved = (ref_mt.pyrange_apply, ref_mt.pyrange_apply_single, ref_main.pyrange_apply, ref_main.pyrange_apply_single,
             b200_mt.E, b200_mt._device_of)
    b200_mt.E, b200_mt._device_of = fake_engine, (lambda kwargs: "cpu")
    calls = []
    exec(compile(block, "INTEGRATION.md:B1", "exec"), ref_mt.__dict__)
    assert ref_mt.pyrange_apply is not saved[0], "the patch did not take"
    # pyranges_main bound the names at import time (pyranges/pyranges_main.py:9); a patch inside the file would have
    # been seen by that import
    inner, inner1 = ref_mt.pyrange_apply, ref_mt.pyrange_apply_single

    def spy(function, self, other, **kw):
        calls.append(getattr(function, "__name__", function))
        return inner(function, self, other, **kw)

    def spy1(function, self, **kw):
        calls.append(getattr(function, "__name__", function))
        return inner1(function, self, **kw)

    ref_main.pyrange_apply, ref_main.pyrange_apply_single = spy, spy1
    yield pr, calls
    (ref_mt.pyrange_apply, ref_mt.pyrange_apply_single, ref_main.pyrange_apply, ref_main.pyrange_apply_single,
     b200_mt.E, b200_mt._device_of) = saved
    md.version = _v


def _mk(pr, tag):
    df = load_input(tag).reset_index(drop=True)
    for c in ("Chromosome", "Strand"):
        if c in df.columns:
            df[c] = df[c].astype(str).astype(object)
    gr = pr.PyRanges(df)
    return pr.PyRanges({k: v for k, v in gr.dfs.items() if not v.empty})


_OPS = ("join", "overlap", "count_overlaps", "coverage", "intersect", "subtract", "cluster", "merge", "k_nearest")
_CASES = [n for n in case_names() if n.split("_")[0] in ("join", "overlap", "count", "coverage", "intersect", "subtract",
                                                          "cluster", "merge")][::4]
# k_nearest reaches the dispatcher with pyranges.methods.k_nearest._nearest (same __name__ as nearest's callable); the
# cases whose answer is a set (ties="different")
_CASES += [n for n in case_names() if n.startswith("knear_") and "_different_" in n][::3]


@pytest.mark.parametrize("name", _CASES)
def test_reference_methods_through_b200_dispatcher(patched_reference, name):
    pr, calls = patched_reference
    meta, want = load_case(name)
    if meta["op"] not in _OPS or meta["kwargs"].get("by"):
        pytest.skip("not a B1 case")
    del calls[:]
    s = _mk(pr, meta["self"])
    kw = dict(meta["kwargs"])
    if meta["other"] is not None:
        got = getattr(s, meta["op"])(_mk(pr, meta["other"]), **kw)
    else:
        got = getattr(s, meta["op"])(**kw)
    assert calls, "the reference method never reached the dispatcher"
    gdf = got.df
    if len(want) == 0:
        assert len(gdf) == 0
        return
    assert_frames_equal(gdf, want, name)
