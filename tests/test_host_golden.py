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
