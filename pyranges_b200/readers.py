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


_TEXT_CHUNK = 1 << 21

# the spellings pandas.read_csv reads as a missing value (pandas/_libs/parsers.pyx STR_NA_VALUES)
_NA_BYTES = np.asarray([b"", b"#N/A", b"#N/A N/A", b"#NA", b"-1.#IND", b"-1.#QNAN", b"-NaN", b"-nan", b"1.#IND",
                        b"1.#QNAN", b"<NA>", b"N/A", b"NA", b"NULL", b"NaN", b"None", b"n/a", b"nan", b"null"])


def _convert_bytes(arr):
    """A fixed-width bytes column the way pandas.read_csv would type it: int64, else float64 (missing values as NaN),
    else text -> (values, None) for numbers, (object array of str with NaN for missing values, True) for text."""
    missing = np.isin(arr, _NA_BYTES)
    if not missing.any():
        for dt in (np.int64, np.float64):
            try:
                return arr.astype(dt), None
            except ValueError:
                pass
        return _to_str_objects(arr), True
    out = np.full(len(arr), np.nan)
    try:
        out[~missing] = arr[~missing].astype(np.float64)
        return out, None
    except ValueError:
        pass
    txt = _to_str_objects(arr)
    txt[missing] = np.nan
    return txt, True


def _to_str_objects(arr):
    """fixed-width bytes array -> object array of str"""
    out = np.empty(len(arr), dtype=object)
    for a in range(0, len(arr), _TEXT_CHUNK):  # in pieces: the "U" intermediate is four times as wide
        part = arr[a:a + _TEXT_CHUNK]
        try:
            out[a:a + _TEXT_CHUNK] = part.astype("U").astype(object)
        except UnicodeDecodeError:  # non-ASCII names: decode element-wise
            out[a:a + _TEXT_CHUNK] = [x.decode("utf-8", errors="replace") for x in part.tolist()]
    return out


def _as_text(vals, inv=None):
    """object array of str (optionally a dictionary `vals` with row codes `inv`) -> the column pandas.read_csv would
    hold: the "str" extension dtype where pandas infers it (pandas >= 3), else the object array"""
    if getattr(pd.options.future, "infer_string", False):
        arr = pd.array(vals, dtype="str")
        return arr.take(inv) if inv is not None else arr
    return vals[inv] if inv is not None else vals


def _text_column(buf, nbytes, off, length, dev):
    """One text column (spans off / length into the file bytes) -> int64 / float64 / text column, without a Python
    loop over the rows: the fields are laid out as fixed-width, zero-padded records on the device; fields of up to 8
    bytes with few distinct values are dictionary-encoded there (one 64-bit word per field, torch.unique) so that only
    the dictionary is decoded; everything else crosses as an "S<width>" array that numpy converts."""
    n = int(off.numel())
    if n == 0:
        return np.empty(0, dtype=object)
    maxlen = int(length.max().item())
    if maxlen > 64:  # long list fields (BlockSizes ...): spans packed on the device, sliced on the host
        return _as_text(np.asarray(_strings(buf, off, length, dev), dtype=object))
    W = max(8, (maxlen + 7) // 8 * 8)
    lane = torch.arange(W, device=dev)
    parts = []
    for a in range(0, n, _TEXT_CHUNK):
        o, l = off[a:a + _TEXT_CHUNK], length[a:a + _TEXT_CHUNK]
        idx = (o[:, None] + lane[None, :]).clamp_(max=max(nbytes - 1, 0))
        rec = torch.where(lane[None, :] < l[:, None], buf[idx], torch.zeros((), dtype=torch.uint8, device=dev))
        parts.append(rec)
    rec = parts[0] if len(parts) == 1 else torch.cat(parts)
    if W == 8:
        uniq, inv = torch.unique(rec.view(torch.int64).reshape(-1), return_inverse=True)
        if uniq.numel() <= max(n // 4, 1):
            vals, text = _convert_bytes(uniq.cpu().numpy().view("S8"))
            return _as_text(vals, inv.cpu().numpy()) if text else vals[inv.cpu().numpy()]
    host = rec.cpu().numpy().view("S%d" % W).reshape(-1)
    vals, text = _convert_bytes(host)
    return _as_text(vals) if text else vals


_STRAND_DTYPE = pd.CategoricalDtype([".", "-", "+"], ordered=True)  # pyranges/methods/init.py:28-30


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
    # categories in sorted order, as astype("category") leaves them (pyranges/methods/init.py:24)
    lex = sorted(range(len(names)), key=lambda i: names[i])
    rank = np.empty(len(names), np.int64)
    rank[lex] = np.arange(len(names))
    names = [names[i] for i in lex]
    chrom_code = torch.from_numpy(rank).to(dev)[inv]
    strand_chars = None
    if ncols >= 6 and not bad_h[2]:
        strand_chars = sorted(torch.unique(strand).cpu().numpy().tolist())

    # the PyRanges is assembled directly in its key order when the Strand column (if any) holds nothing but the
    # characters the reference's dtype knows; otherwise through the constructor
    direct = not as_df and (ncols < 6 or (strand_chars is not None and set(strand_chars) <= {43, 45, 46}))
    order = seg = keys = None
    if direct:
        stranded = ncols >= 6 and set(strand_chars) <= {43, 45}
        key = chrom_code * 256 + strand.to(torch.int64) if stranded else chrom_code
        ukeys, kinv = torch.unique(key, return_inverse=True)
        labels = [((names[k >> 8], chr(k & 255)) if stranded else names[k]) for k in ukeys.cpu().numpy().tolist()]
        keys = natsorted(labels)
        where = {lab: p for p, lab in enumerate(keys)}
        pos = torch.from_numpy(np.asarray([where[lab] for lab in labels], np.int64)).to(dev)[kinv]
        seg = np.zeros(len(keys) + 1, np.int64)
        np.cumsum(torch.bincount(pos, minlength=len(keys)).cpu().numpy(), out=seg[1:])
        rows = torch.arange(n_rows, dtype=torch.int32, device=dev)
        E.radix_sort_u64(pos, rows, 0, max(1, int(len(keys)).bit_length()))  # stable: file order inside a key
        order = rows.to(torch.int64)
    take = (lambda t: t[order].contiguous()) if order is not None else (lambda t: t)  # noqa: E731

    start_d, end_d = take(start), take(end)
    data = {"Chromosome": pd.Categorical.from_codes(take(chrom_code).cpu().numpy(), categories=names),
            "Start": start_d.cpu().numpy(), "End": end_d.cpu().numpy()}
    for c in range(3, ncols):
        name = BED_COLUMNS[c]
        if c == 4 and not bad_h[1]:
            data[name] = take(score).cpu().numpy()
        elif c == 5 and strand_chars is not None:
            sv = take(strand).cpu().numpy()
            lut = np.zeros(256, np.int8)
            if direct:
                lut[[46, 45, 43]] = [0, 1, 2]
                data[name] = pd.Categorical.from_codes(lut[sv], dtype=_STRAND_DTYPE)
            else:
                lut[strand_chars] = np.arange(len(strand_chars))
                data[name] = pd.Categorical.from_codes(lut[sv], categories=[chr(x) for x in strand_chars])
        else:
            o, l = col(c)
            vals = _text_column(buf, nbytes, take(o), take(l), dev)
            data[name] = pd.Categorical(vals) if c == 5 else vals
    if not direct:
        df = pd.DataFrame(data)
        return df if as_df else PyRanges(df)
    # frames of the keys: row ranges of one frame in key order whose index holds the rows' positions in the file,
    # which is what the groupby of the constructor leaves (pyranges/methods/init.py:140-150)
    big = M._new_frame(data, n_rows)
    big.index = pd.Index(order.cpu().numpy())
    gr = PyRanges._wrap({k: big.iloc[int(seg[i]):int(seg[i + 1])] for i, k in enumerate(keys)})
    fr = M._Frames(natsorted(gr.dfs.items()), dev, upload=False)
    if fr.keys == keys:
        fr.starts, fr.ends = data["Start"], data["End"]
        fr.iv = E.DeviceIntervals(start_d, end_d, fr.seg_off)
        M.seed_frame_cache(gr, dev, fr)
    return gr
