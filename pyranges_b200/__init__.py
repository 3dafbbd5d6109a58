"""pyranges_b200 -- B200-native engine for the pyranges interval hot path."""
__version__ = "0.1.0"


def read_bed(f, as_df=False, nrows=None, device=None):
    """pyranges.read_bed with the parse on the device (pyranges/readers.py:63-153); see pyranges_b200.readers."""
    from .readers import read_bed as _read_bed

    return _read_bed(f, as_df=as_df, nrows=nrows, device=device)
