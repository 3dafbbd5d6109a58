"""CPU: host logic of pyranges_b200 (key order, partner selection, fillers, frame glue) against the golden
fixtures of the reference, with the device engine replaced by the oracle-backed stand-in
(tests/fake_engine.py).  The same cases run through the real CUDA engine in tests/test_gpu_golden.py."""
import pytest

from tests import fake_engine
from tests.golden_util import case_names, check_case


@pytest.fixture(autouse=True)
def oracle_engine(monkeypatch):
    from pyranges_b200 import multithreaded

    monkeypatch.setattr(multithreaded, "E", fake_engine)
    monkeypatch.setattr(multithreaded, "_device_of", lambda kwargs: "cpu")


@pytest.mark.parametrize("name", case_names())
def test_golden_host(name):
    from pyranges_b200.pyranges_main import PyRanges

    check_case(name, PyRanges)


def test_count_overlaps_helper_matches_number_overlapping():
    """_count_overlaps (pyranges/methods/intersection.py:86-99, the pr.count_overlaps helper) = the NumberOverlaps
    column of count_overlaps as int32, per `how`."""
    import numpy as np

    from pyranges_b200.multithreaded import pyrange_apply
    from pyranges_b200.pyranges_main import PyRanges, fill_kwargs
    from tests.golden_util import load_input

    a, b = PyRanges(load_input("A")), PyRanges(load_input("B"))
    for st in (None, "same"):
        want = a.count_overlaps(b, strandedness=st).df
        got = pyrange_apply("_count_overlaps", a, b, **fill_kwargs({"strandedness": st, "how": None, "name": "s1"}))
        gdf = PyRanges(got).df
        assert str(gdf.s1.dtype) == "int32"
        assert np.array_equal(gdf.s1.values, want.NumberOverlaps.values)
        first = PyRanges(pyrange_apply("_count_overlaps", a, b, **fill_kwargs({"strandedness": st, "how": "first", "name": "s1"}))).df
        assert np.array_equal(first.s1.values, (want.NumberOverlaps.values > 0).astype(np.int32))
