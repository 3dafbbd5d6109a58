"""numpy-in / numpy-out facade over the device engine: the batched twin of ``ncls.NCLS`` as the
reference calls it (pyranges/methods/join.py:8-23): host arrays go in, host index arrays come out,
host<->device copies happen inside.  This is the call bench.py times for the end-to-end number."""
import numpy as np
import torch

from . import engine as E


def _host_tensor(a):
    if isinstance(a, torch.Tensor):
        assert a.dtype == torch.int64 and a.device.type == "cpu"
        return a
    return torch.from_numpy(np.ascontiguousarray(a, dtype=np.int64))


_streams = {}


def _side_streams(dev):
    """copy-in / copy-out streams, created once per device"""
    key = (dev.type, dev.index)
    if key not in _streams:
        _streams[key] = (torch.cuda.Stream(dev), torch.cuda.Stream(dev))
    return _streams[key]


def _to_device(a, device):
    t = _host_tensor(a)
    return t.to(device, non_blocking=t.is_pinned())


class NCLSBatch:
    """Subject index over all keys.  starts/ends: host int64 arrays (numpy or pinned torch) of the
    concatenated partner frames; seg_off: their key segments; group_ranges: which consecutive segments
    form one partner frame (None: every segment)."""

    def __init__(self, starts, ends, seg_off, group_ranges=None, device="cuda", ends_perm=False, bounds=None):
        dev = E._use(device)
        self.device = dev
        self.iv = E.DeviceIntervals(_to_device(starts, dev), _to_device(ends, dev), seg_off, bounds=bounds)
        self.group_ranges, self.ends_perm = group_ranges, ends_perm
        self._rank = None   # rank index (first / last / containment / nearest ...): built on first use
        self._lists = None  # list index of the join / count path, keyed by its reach hint

    @property
    def index(self):
        if self._rank is None:
            self._rank = E.SubjectIndex(self.iv, self.group_ranges, ends_perm=self.ends_perm)
        return self._rank

    def _pairs_index(self, hs, he):
        """the index all_overlaps_both / count run on: the list index (fused join) when these subjects and queries
        suit it -- its reach hint is the longest of a sample of the query rows (a hint only: longer queries stay exact)
        -- else the rank index"""
        n = hs.numel()
        step = max(1, n // 65536)
        ln = (he[::step] - hs[::step])
        mean_len = float(ln.double().mean()) if n else 0.0
        if not E.ListIndex.supported(self.iv, self.group_ranges, mean_len):
            return self.index
        reach = int(ln.max()) if n else 0
        if self._lists is None or self._lists.reach < reach:
            self._lists = E.ListIndex(self.iv, self.group_ranges, reach=reach)
        return self._lists

    @staticmethod
    def _pairs(ix, q, capacity=None):
        if isinstance(ix, E.ListIndex):
            return ix.pairs(q, capacity=capacity)
        _, a, b = ix.pairs(q, E.IV_MODE_ALL, capacity=capacity, want_counts=False)
        return a, b

    def _queries(self, starts, ends, seg_off, seg_group, slack=0):
        q = E.DeviceIntervals(_to_device(starts, self.device), _to_device(ends, self.device), seg_off)
        return E.Queries(q, seg_group, slack=slack)

    def all_overlaps_both(self, starts, ends, seg_off, seg_group, slack=0, out=None, chunks=None):
        """-> (query rows, subject rows) host int64 arrays, query order (NCLS.all_overlaps_both).

        Large pinned inputs are processed in row chunks on three streams: the host->device copies of all
        chunks are queued up front on a copy stream, every chunk is counted / emitted as soon as it has
        landed, and its pairs travel back on a third stream while the next chunk is computed -- PCIe runs in
        both directions and the kernels hide behind it."""
        hs, he = _host_tensor(starts), _host_tensor(ends)
        n = hs.numel()
        seg_off = np.asarray(seg_off, dtype=np.int64)
        if chunks is None:
            chunks = 8 if (n >= (1 << 21) and hs.is_pinned() and he.is_pinned() and out is not None) else 1
        ix = self._pairs_index(hs, he)
        if chunks <= 1:
            q = self._queries(hs, he, seg_off, seg_group, slack)
            a, b = self._pairs(ix, q)
            if out is not None:  # pinned staging buffers supplied by the caller
                p = a.numel()
                out[0][:p].copy_(a, non_blocking=True)
                out[1][:p].copy_(b, non_blocking=True)
                torch.cuda.current_stream().synchronize()
                return out[0][:p], out[1][:p]
            return a.cpu().numpy(), b.cpu().numpy()
        dev = self.device
        main = torch.cuda.current_stream(dev)
        s_copy, s_back = _side_streams(dev)
        step = -(-n // chunks)
        step = (step + 4095) // 4096 * 4096  # chunk edges on the count kernel's 4096-row blocks
        bounds = [(a, min(a + step, n)) for a in range(0, n, step)]
        # the chunks' small descriptors (segment offsets, partner map) go up first, in one copy: issued later they
        # would queue behind the bulk uploads on the copy engine and stall the host once per chunk
        offs = [np.clip(seg_off, a, b) - a for a, b in bounds]
        sg32 = np.ascontiguousarray(seg_group, dtype=np.int32)
        k1 = len(seg_off)
        packed = np.concatenate(offs + [sg32.astype(np.int64)])
        packed_dev = torch.from_numpy(packed).to(dev)
        sg_dev = packed_dev[len(bounds) * k1:].to(torch.int32)
        # device columns from the COPY stream's pool (reused call after call): the uploads start at once, they do not
        # wait for whatever the main stream is still doing (e.g. the index build queued just before this call)
        landed = []
        with torch.cuda.stream(s_copy):
            d_s = torch.empty(n, dtype=torch.int64, device=dev)
            d_e = torch.empty(n, dtype=torch.int64, device=dev)
            for a, b in bounds:
                d_s[a:b].copy_(hs[a:b], non_blocking=True)
                d_e[a:b].copy_(he[a:b], non_blocking=True)
                ev = torch.cuda.Event()
                ev.record(s_copy)
                landed.append((d_s[a:b], d_e[a:b], ev))
        d_s.record_stream(main)
        d_e.record_stream(main)
        pos = 0
        keep = []
        hint = None  # pairs of the previous chunk: the next chunk's emit is launched before its total is known
        for ci, ((a, b), (ds, de, ev)) in enumerate(zip(bounds, landed)):
            main.wait_event(ev)
            # the chunk's own segment offsets (empty segments allowed)
            civ = E.DeviceIntervals(ds, de, offs[ci], seg_off_dev=packed_dev[ci * k1:(ci + 1) * k1])
            q = E.Queries(civ, sg32, slack=slack, desc_dev=sg_dev)
            qa, qb = self._pairs(ix, q, capacity=hint)
            hint = int(qa.numel() * 1.25) + 4096
            if a:
                qa.add_(a)
            p = qa.numel()
            if pos + p > out[0].numel():
                raise ValueError("all_overlaps_both: output buffers too small (%d pairs so far)" % (pos + p))
            done = torch.cuda.Event()
            done.record(main)
            s_back.wait_event(done)
            with torch.cuda.stream(s_back):
                out[0][pos:pos + p].copy_(qa, non_blocking=True)
                out[1][pos:pos + p].copy_(qb, non_blocking=True)
            qa.record_stream(s_back)
            qb.record_stream(s_back)
            keep.append((q, qa, qb))
            pos += p
        s_back.synchronize()
        return out[0][:pos], out[1][:pos]

    def count(self, starts, ends, seg_off, seg_group, slack=0, out=None):
        """-> overlaps per query, host int64 array (the NumberOverlaps column, coverage.py:26-40)."""
        hs, he = _host_tensor(starts), _host_tensor(ends)
        q = self._queries(hs, he, seg_off, seg_group, slack)
        ix = self._pairs_index(hs, he)
        c = ix.count(q, E.IV_MODE_ALL) if isinstance(ix, E.ListIndex) else ix.count(q, E.IV_MODE_ALL, for_emit=False)[0]
        if out is not None:
            out.copy_(c, non_blocking=True)
            torch.cuda.current_stream().synchronize()
            return out
        return c.cpu().numpy()
