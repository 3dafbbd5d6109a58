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
        self.index = E.SubjectIndex(self.iv, group_ranges, ends_perm=ends_perm)

    def _queries(self, starts, ends, seg_off, seg_group, slack=0):
        q = E.DeviceIntervals(_to_device(starts, self.device), _to_device(ends, self.device), seg_off)
        return E.Queries(q, seg_group, slack=slack)

    def all_overlaps_both(self, starts, ends, seg_off, seg_group, slack=0, out=None):
        """-> (query rows, subject rows) host int64 arrays, query order (NCLS.all_overlaps_both)."""
        q = self._queries(starts, ends, seg_off, seg_group, slack)
        _, a, b = self.index.pairs(q, E.IV_MODE_ALL)
        if out is not None:  # pinned staging buffers supplied by the caller
            p = a.numel()
            out[0][:p].copy_(a, non_blocking=True)
            out[1][:p].copy_(b, non_blocking=True)
            torch.cuda.current_stream().synchronize()
            return out[0][:p], out[1][:p]
        return a.cpu().numpy(), b.cpu().numpy()

    def count(self, starts, ends, seg_off, seg_group, slack=0, out=None):
        """-> overlaps per query, host int64 array (the NumberOverlaps column, coverage.py:26-40)."""
        q = self._queries(starts, ends, seg_off, seg_group, slack)
        c, _, _ = self.index.count(q, E.IV_MODE_ALL, for_emit=False)
        if out is not None:
            out.copy_(c, non_blocking=True)
            torch.cuda.current_stream().synchronize()
            return out
        return c.cpu().numpy()
