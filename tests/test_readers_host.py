"""Host side of read_bed (pyranges_b200/readers.py): the typing of a text column -- int64 / float64 / text, missing
values -- against pandas.read_csv, the call the reference's read_bed makes (pyranges/readers.py:133-139).  The column
assembly is torch + numpy, so it runs here on CPU tensors; the parse kernels are covered by tests/test_gpu_ingest.py."""
import io

import numpy as np
import pandas as pd
import pytest
import torch

from pyranges_b200 import readers as R


def _column(values):
    """one text column of a two-column TSV -> (bytes tensor, nbytes, offsets, lengths)"""
    text = "".join("x\t%s\n" % v for v in values).encode()
    off, length, p = [], [], 0
    for v in values:
        b = v.encode()
        off.append(p + 2)
        length.append(len(b))
        p += 2 + len(b) + 1
    buf = torch.from_numpy(np.frombuffer(text + b"\0" * 16, dtype=np.uint8).copy())
    return text, buf, len(text), torch.tensor(off, dtype=torch.int64), torch.tensor(length, dtype=torch.int32)


CASES = {
    "few_short": ["g%d" % (i % 7) for i in range(500)],
    "distinct_short": ["u%05d" % i for i in range(500)],
    "wide": ["a_rather_long_identifier_%d" % i for i in range(300)],
    "ints_short": [str(i % 50) for i in range(400)],
    "ints_wide": [str(10 ** 10 + i) for i in range(400)],
    "negative": [str(-i) for i in range(300)],
    "floats": ["%.3f" % (i / 7) for i in range(300)],
    "floats_few": ["0.5", "1.5", "2"] * 100,
    "exp": ["1e3", "2.5e-2", "7"] * 50,
    "mixed_text": ["12", "x12", "3.5"] * 60,
    "rgb": ["255,0,%d" % (i % 255) for i in range(300)],
    "missing_text": ["a", "", "NA", "b", "nan", "null", "None", "c"] * 30,
    "missing_number": ["1", "", "3", "NA", "5"] * 40,
    "missing_float_wide": ["1.25", "N/A", "1234567890.5", ""] * 40,
    "all_missing": [""] * 64,
    "dot": ["."] * 100,
    "utf8": ["gène%d" % (i % 5) for i in range(100)],
    "utf8_wide": ["gène_avec_un_long_nom_%d" % i for i in range(100)],
}


@pytest.mark.parametrize("name", sorted(CASES))
def test_text_column_matches_read_csv(name):
    values = CASES[name]
    text, buf, nbytes, off, length = _column(values)
    want = pd.read_csv(io.BytesIO(text), header=None, sep="\t")[1]
    got = pd.Series(R._text_column(buf, nbytes, off, length, torch.device("cpu")))
    assert str(got.dtype) == str(want.dtype), (got.dtype, want.dtype)
    assert len(got) == len(want)
    assert (got.isna().values == want.isna().values).all()
    keep = ~want.isna().values
    if want.dtype.kind == "f":
        assert np.array_equal(got.values[keep], want.values[keep])
    else:
        assert (np.asarray(got[keep], dtype=object) == np.asarray(want[keep], dtype=object)).all()


def test_text_column_chunked(monkeypatch):
    """more rows than one chunk of fixed-width records"""
    monkeypatch.setattr(R, "_TEXT_CHUNK", 64)
    values = ["name_number_%d" % i for i in range(1000)]
    text, buf, nbytes, off, length = _column(values)
    got = R._text_column(buf, nbytes, off, length, torch.device("cpu"))
    assert list(got) == values
