"""GPU: every golden fixture of the reference (tests/golden/, oracle/make_golden.py) through the public
PyRanges API of pyranges_b200 -> pyrange_apply -> C ABI -> sm_100a kernels."""
import pytest

from tests.golden_util import case_names, check_case

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("name", case_names())
def test_golden_gpu(name):
    from pyranges_b200.pyranges_main import PyRanges

    check_case(name, PyRanges)
