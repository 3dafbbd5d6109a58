"""read_bed with the parse on the device (pyranges/readers.py:63-153).

The reference hands the file to ``pandas.read_csv(sep="\\t")`` and the frame to ``PyRanges``; every later operation
then concatenates and uploads Start / End again.  Here the file's bytes go to the device once: lines and fields are
found there, Start / End / Score are parsed there (``csrc/ingest.cu``), the rows are brought into the
(Chromosome[, Strand]) key order of the PyRanges on the device, and the coordinate columns stay there as the cached
device ingest of the new object -- the first join / count / cluster on it uploads nothing.

The DataFrame the PyRanges holds is assembled on the host from the parsed columns: integers come back as arrays,
Chromosome / Strand as categoricals built from a dictionary of the few distinct values, free-text columns (Name,
ItemRGB ...) from their packed bytes.  Column naming, header detection (a non-integer second field in the first line)
and ``nrows`` follow the reference.  There is no host fallback for the parse: rows whose Start / End are not integers
raise.
"""
import ctypes as C
import gzip

import numpy as np
import pandas as pd
import torch

from . import _cabi as cabi
from . import engine as E
from ._natsort import natsorted

BED_COLUMNS = ("Chromosome Start End Name Score Strand ThickStart ThickEnd ItemRGB BlockCount BlockSizes "
               "BlockStarts").split()


def _ptr(t):
    return C.c_void_p(t.data_ptr()) if t is not None and t.numel() else C.c_void_p(0)


def _stream():
    return C.c_void_p(torch.cuda.current_stream().cuda_stream)


def _device_text(path, device):
    """file bytes -> uint8 device tensor padded to a multiple of 16 (pinned staging, one copy)"""
    if str(path).endswith(".gz"):
        with gzip.open(path, "rb") as f:
            raw = np.frombuffer(bytearray(f.read()), dtype=np.uint8)
    else:
        raw = np.fromfile(path, dtype=np.uint8)
    n = int(raw.size)
    host = torch.empty((n + 15) // 16 * 16 + 16, dtype=torch.uint8).pin_memory()
    host[:n] = torch.from_numpy(raw)
    host[n:] = 0
    return host.to(device, non_blocking=True), n


def _strings(buf, off, length, device):
    """the spans (off, length) of one text column -> list of str"""
    L = cabi.lib()
    n = off.numel()
    out_off = torch.empty(n + 1, dtype=torch.int64, device=device)
    ws = torch.empty((n // 4096 + 64) * 8 * 4 + 4096, dtype=torch.uint8, device=device)
    cabi.check(L.iv_text_gather(_ptr(buf), _ptr(off), _ptr(length), n, _ptr(out_off), None, _ptr(ws), ws.numel(),
                                _stream()), "iv_text_gather")
    total = int(out_off[n].item())
    data = torch.empty(max(total, 1), dtype=torch.uint8, device=device)
    cabi.check(L.iv_text_gather(_ptr(buf), _ptr(off), _ptr(length), n, _ptr(out_off), _ptr(data), None, 0, _stream()),
               "iv_text_gather")
    o = out_off.cpu().numpy()
    b = data[:total].cpu().numpy().tobytes()
    return [b[o[i]:o[i + 1]].decode() for i in range(n)]


def _infer(values):
    """dtype inference of one text column as pandas.read_csv would do it: int64, else float64, else object"""
    s = pd.Series(values, dtype=object)
    for conv in (lambda x: x.astype(np.int64), lambda x: x.astype(np.float64)):
        try:
            return conv(s).values
        except (ValueError, TypeError):
            pass
    return s.values


def read_bed(f, as_df=False, nrows=None, device=None):
    """Return bed file as PyRanges (pyranges/readers.py:63-153); the coordinate columns are parsed on `device`
    (default: the current CUDA device) and stay there as the object's device ingest."""
    from . import multithreaded as M
    from .pyranges_main import PyRanges

    L = cabi.lib()
    dev = E._use(device if device is not None else M._device_of({}))
    buf, nbytes = _device_text(f, dev)
    if nbytes == 0:
        df = pd.DataFrame(columns=BED_COLUMNS[:3])
        return df if as_df else PyRanges(df)
    ws = torch.empty(L.iv_text_workspace_bytes(nbytes), dtype=torch.uint8, device=dev)
    newlines = torch.zeros(1, dtype=torch.int64, device=dev)
    cabi.check(L.iv_text_count_lines(_ptr(buf), nbytes, _ptr(ws), ws.numel(), _ptr(newlines), _stream()),
               "iv_text_count_lines")
    n_lines = int(newlines.item()) + (0 if int(buf[nbytes - 1].item()) == 10 else 1)
    line_start = torch.empty(n_lines + 1, dtype=torch.int64, device=dev)
    cabi.check(L.iv_text_line_starts(_ptr(buf), nbytes, _ptr(ws), n_lines, _ptr(line_start), _stream()),
               "iv_text_line_starts")
    # header: the reference looks at the second field of the first line (readers.py:118-131)
    first = buf[:int(line_start[1].item())].cpu().numpy().tobytes().decode(errors="replace").split()
    header = 1
    try:
        int(first[1])
        header = 0
    except (ValueError, IndexError):
        pass
    n_rows = n_lines - header
    # a trailing empty line (file ending in "\n\n") holds no row
    while n_rows > 0 and int((line_start[header + n_rows] - line_start[header + n_rows - 1]).item()) <= 1:
        n_rows -= 1
    if nrows is not None:
        n_rows = min(n_rows, int(nrows))
    i64 = lambda: torch.empty(max(n_rows, 1), dtype=torch.int64, device=dev)  # noqa: E731
    start, end, score, chash = i64(), i64(), i64(), i64()
    strand = torch.empty(max(n_rows, 1), dtype=torch.uint8, device=dev)
    nfields = torch.empty(max(n_rows, 1), dtype=torch.uint8, device=dev)
    foff = torch.empty(12 * max(n_rows, 1), dtype=torch.int64, device=dev)
    flen = torch.empty(12 * max(n_rows, 1), dtype=torch.int32, device=dev)
    bad = torch.zeros(3, dtype=torch.int32, device=dev)
    cabi.check(L.iv_bed_parse(_ptr(buf), _ptr(line_start), header, n_rows, _ptr(start), _ptr(end), _ptr(score),
                              _ptr(chash), _ptr(strand), _ptr(nfields), _ptr(foff), _ptr(flen), _ptr(bad), _stream()),
               "iv_bed_parse")
    start, end, score, chash, strand, nfields = (t[:n_rows] for t in (start, end, score, chash, strand, nfields))
    if n_rows == 0:
        df = pd.DataFrame(columns=BED_COLUMNS[:3])
        return df if as_df else PyRanges(df)
    bad_h = bad.cpu().numpy()
    ncols = int(nfields.min().item())
    if int(nfields.max().item()) != ncols or ncols < 3:
        raise ValueError("read_bed: %s does not have the same number (3-12) of tab-separated fields on every line" % f)
    if bad_h[0]:
        raise ValueError("read_bed: %d rows of %s have a Start / End that is not an integer" % (bad_h[0], f))
    ncols = min(ncols, 12)
    col = lambda c: (foff[c * max(n_rows, 1):c * max(n_rows, 1) + n_rows],  # noqa: E731
                     flen[c * max(n_rows, 1):c * max(n_rows, 1) + n_rows])
    # dictionary of the chromosome names: one representative span per distinct hash
    uniq, inv = torch.unique(chash, return_inverse=True)
    rep = torch.full((uniq.numel(),), n_rows, dtype=torch.int64, device=dev)
    rep.scatter_reduce_(0, inv, torch.arange(n_rows, device=dev), "amin")
    o0, l0 = col(0)
    names = _strings(buf, o0[rep], l0[rep], dev)
    # two different names behind one 64-bit FNV-1a hash: practically impossible for a handful of names; a length
    # mismatch inside a hash class would give it away
    if len(set(names)) != len(names) or not bool((l0 == l0[rep][inv]).all()):
        raise ValueError("read_bed: chromosome-name hash collision")
    data = {"Chromosome": pd.Categorical.from_codes(inv.cpu().numpy(), categories=names),
            "Start": start.cpu().numpy(), "End": end.cpu().numpy()}
    for c in range(3, ncols):
        name = BED_COLUMNS[c]
        if c == 4 and not bad_h[1]:
            data[name] = score.cpu().numpy()
        elif c == 5 and not bad_h[2]:
            sv = strand.cpu().numpy()
            cats = sorted(set(sv.tolist()))
            lut = np.zeros(256, np.int64)
            lut[cats] = np.arange(len(cats))
            data[name] = pd.Categorical.from_codes(lut[sv], categories=[chr(x) for x in cats])
        else:
            o, l = col(c)
            vals = _infer(_strings(buf, o, l, dev))
            data[name] = pd.Categorical(vals) if c == 5 else vals
    df = pd.DataFrame(data)
    if as_df:
        return df
    gr = PyRanges(df)
    _seed_device_columns(gr, start, end, inv, strand if ("Strand" in df.columns and gr.stranded) else None, names, dev)
    return gr


def _seed_device_columns(gr, start, end, chrom_code, strand, names, dev):
    """The parsed Start / End, brought into the key order of `gr` on the device (stable: rows keep their file order
    inside a key, like the groupby of the constructor), become the cached device ingest of the object."""
    from . import multithreaded as M

    items = natsorted(gr.dfs.items())
    if not items:
        return
    n = start.numel()
    # key id of every row -> position of the key in natsorted order
    nkeys = len(names) * (256 if strand is not None else 1)
    pos_of = np.full(nkeys, len(items), dtype=np.int64)
    code_of = {nm: i for i, nm in enumerate(names)}
    for p, (k, _) in enumerate(items):
        if isinstance(k, tuple):
            pos_of[code_of[str(k[0])] * 256 + ord(k[1])] = p
        else:
            pos_of[code_of[str(k)]] = p
    key = chrom_code * 256 + strand.to(torch.int64) if strand is not None else chrom_code
    pos = torch.from_numpy(pos_of).to(dev)[key]
    if int(pos.max().item()) >= len(items):
        return  # rows the constructor regrouped differently: leave the upload to the first operation
    rows = torch.arange(n, dtype=torch.int32, device=dev)
    E.radix_sort_u64(pos, rows, 0, max(1, int(len(items)).bit_length()))
    order = rows.to(torch.int64)
    s_dev, e_dev = start[order].contiguous(), end[order].contiguous()
    f = M._Frames(items, dev, upload=False)
    f.starts = np.concatenate([np.asarray(d["Start"].values, dtype=np.int64) for d in f.dfs])
    f.ends = np.concatenate([np.asarray(d["End"].values, dtype=np.int64) for d in f.dfs])
    f.iv = E.DeviceIntervals(s_dev, e_dev, f.seg_off)
    M.seed_frame_cache(gr, dev, f)
