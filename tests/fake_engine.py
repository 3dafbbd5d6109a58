"""ORACLE-BACKED stand-in for pyranges_b200.engine -- TEST INFRASTRUCTURE ONLY.

The CPU test-suite (-m "not gpu") monkeypatches it into pyranges_b200.multithreaded so that the host
logic (key order, partner selection, -1 fillers, frame glue) can be checked against the golden fixtures
without a GPU.  The product never imports this module; on a GPU box the same golden cases run through the
real CUDA engine (tests/test_gpu_golden.py)."""
import numpy as np
import torch

from oracle import oracle as orc

IV_MODE_ALL, IV_MODE_ANY, IV_MODE_CONTAIN, IV_MODE_LAST = 0, 1, 2, 3
IV_NEAREST_BOTH, IV_NEAREST_PREVIOUS, IV_NEAREST_NEXT = 0, 1, 2


def _t(a):
    return torch.from_numpy(np.ascontiguousarray(a))


class DeviceIntervals:
    def __init__(self, s, e, seg_off):
        self.s, self.e, self.seg_off = np.asarray(s, np.int64), np.asarray(e, np.int64), np.asarray(seg_off, np.int64)
        self.n, self.n_seg = len(self.s), len(self.seg_off) - 1

    @classmethod
    def from_numpy(cls, s, e, seg_off, device=None):
        return cls(s, e, seg_off)


class Queries:
    def __init__(self, iv, seg_group, seg_mode=None, slack=0, label=None):
        self.iv, self.seg_group, self.seg_mode, self.slack = iv, np.asarray(seg_group), seg_mode, slack


class SubjectIndex:
    def __init__(self, iv, seg_ranges=None, ends_perm=False):
        self.iv = iv
        if seg_ranges is None:
            seg_ranges = [(i, i + 1) for i in range(iv.n_seg)]
        self.off = np.asarray([iv.seg_off[a] for a, _ in seg_ranges] + [iv.n], np.int64)

    def _per_key(self, q):
        for k in range(q.iv.n_seg):
            g = q.seg_group[k]
            a, b = q.iv.seg_off[k], q.iv.seg_off[k + 1]
            if g < 0 or a == b or self.off[g] == self.off[g + 1]:
                continue
            sa, sb = self.off[g], self.off[g + 1]
            s, e = q.iv.s[a:b].copy(), q.iv.e[a:b].copy()
            if q.slack:
                s, e = np.maximum(s - q.slack, 0), e + q.slack
            yield k, a, b, sa, sb, s, e, orc.NCLS(self.iv.s[sa:sb], self.iv.e[sa:sb], np.arange(sa, sb, dtype=np.int64))

    def pairs(self, q, mode=IV_MODE_ALL, want_q=True, want_s=True, s_label=None, capacity=None, want_counts=True):
        A, B = [np.zeros(0, np.int64)], [np.zeros(0, np.int64)]
        for k, a, b, sa, sb, s, e, t in self._per_key(q):
            ids = np.arange(a, b, dtype=np.int64)
            fn = {IV_MODE_ALL: t.all_overlaps_both, IV_MODE_CONTAIN: t.all_containments_both,
                  IV_MODE_ANY: t.first_overlap_both, IV_MODE_LAST: t.last_overlap_both}[mode]
            x, y = fn(s, e, ids)
            o = np.argsort(x, kind="stable")
            A.append(x[o])
            B.append(y[o])
        A, B = np.concatenate(A), np.concatenate(B)
        return (_t(np.bincount(A, minlength=q.iv.n).astype(np.int64)) if want_counts else None), _t(A), _t(B)

    def count(self, q, mode=IV_MODE_ALL, for_emit=True, want_counts=True):
        c, _, _ = self.pairs(q, mode)
        return c, None, _t(np.asarray([int(c.sum())]))

    def coverage_bp(self, q, want_fraction=True):
        bp = np.zeros(q.iv.n, np.int64)
        for k, a, b, sa, sb, s, e, t in self._per_key(q):
            bp[a:b] = t.coverage(s, e, np.arange(a, b))
        ln = (q.iv.e - q.iv.s).astype(np.float64)
        with np.errstate(all="ignore"):
            fr = np.nan_to_num(bp / ln)
        return _t(bp), _t(fr)

    def subtract(self, q):
        Q, S, En = [np.zeros(0, np.int64)], [np.zeros(0, np.int64)], [np.zeros(0, np.int64)]
        hit_keys = {}
        for k, a, b, sa, sb, s, e, t in self._per_key(q):
            hit_keys[k] = (a, b, s, e, t)
        for k in range(q.iv.n_seg):
            a, b = q.iv.seg_off[k], q.iv.seg_off[k + 1]
            s, e = q.iv.s[a:b], q.iv.e[a:b]
            rows = np.arange(a, b, dtype=np.int64)
            if k not in hit_keys:
                Q.append(rows), S.append(s), En.append(e)
                continue
            t = hit_keys[k][4]
            idx, ns, ne = t.set_difference_helper(s, e, rows)
            untouched = np.setdiff1d(rows, idx)
            keep = ns != -1
            qq = np.concatenate([untouched, idx[keep]])
            ss = np.concatenate([q.iv.s[untouched], ns[keep]])
            ee = np.concatenate([q.iv.e[untouched], ne[keep]])
            o = np.lexsort([ss, qq])
            Q.append(qq[o]), S.append(ss[o]), En.append(ee[o])
        return _t(np.concatenate(Q)), _t(np.concatenate(S)), _t(np.concatenate(En))

    def nearest(self, q, overlap=True):
        row = np.full(q.iv.n, -1, np.int64)
        dist = np.full(q.iv.n, -1, np.int64)
        for k, a, b, sa, sb, s, e, t in self._per_key(q):
            S, En = self.iv.s[sa:sb], self.iv.e[sa:sb]
            mode = IV_NEAREST_BOTH if q.seg_mode is None else q.seg_mode[k]
            hit = np.zeros(b - a, bool)
            if overlap:
                x, y = t.first_overlap_both(s, e, np.arange(b - a, dtype=np.int64))
                hit[x] = True
                row[a + x], dist[a + x] = y, 0
            rest = np.flatnonzero(~hit)
            o_e, o_s = np.argsort(En, kind="stable"), np.argsort(S, kind="stable")
            ls, le = np.argsort(s[rest], kind="stable"), np.argsort(e[rest], kind="stable")
            pi, pd_ = orc.nearest_previous_nonoverlapping(s[rest][ls], En[o_e] - 1, o_e)
            ni, nd = orc.nearest_next_nonoverlapping(e[rest][le] - 1, S[o_s], o_s)
            pi2, pd2, ni2, nd2 = (np.empty(len(rest), np.int64) for _ in range(4))
            pi2[ls], pd2[ls], ni2[le], nd2[le] = pi, pd_, ni, nd
            if mode == IV_NEAREST_PREVIOUS:
                ri, rd = pi2, pd2
            elif mode == IV_NEAREST_NEXT:
                ri, rd = ni2, nd2
            else:
                ri, rd = orc.nearest_nonoverlapping(pi2, pd2, ni2, nd2)
            row[a + rest] = np.where(ri >= 0, ri + sa, -1)
            dist[a + rest] = rd
        return _t(row), _t(dist)

    def k_nearest(self, q, k, ties=None):
        k = np.asarray(k.cpu().numpy() if hasattr(k, "cpu") else k, dtype=np.int64)
        Q, R, D = [np.zeros(0, np.int64)], [np.zeros(0, np.int64)], [np.zeros(0, np.int64)]
        for kk, a, b, sa, sb, s, e, t in self._per_key(q):
            S, En = self.iv.s[sa:sb], self.iv.e[sa:sb]
            mode = IV_NEAREST_BOTH if q.seg_mode is None else q.seg_mode[kk]
            o_e, o_s = np.argsort(En, kind="stable"), np.argsort(S, kind="stable")
            lab = np.arange(b - a, dtype=np.int64)
            kq = k[a:b]
            parts = []
            if mode != IV_NEAREST_NEXT:
                ix = np.searchsorted(En[o_e], s, side="right") - 1
                v = ix >= 0
                li, rp, dd = orc.k_nearest_previous_nonoverlapping(s[v], En[o_e], lab[v], ix[v], kq, ties)
                parts.append((li, o_e[rp], dd + 1, np.zeros(len(li), np.int64)))
            if mode != IV_NEAREST_PREVIOUS:
                ix = np.searchsorted(S[o_s], e, side="left")
                v = ix < len(S)
                li, rp, dd = orc.k_nearest_next_nonoverlapping(e[v], S[o_s], lab[v], ix[v], kq, ties)
                parts.append((li, o_s[rp], dd + 1, np.ones(len(li), np.int64)))
            li = np.concatenate([p[0] for p in parts])
            rr = np.concatenate([p[1] for p in parts])
            dd = np.concatenate([p[2] for p in parts])
            pn = np.concatenate([p[3] for p in parts])
            o = np.lexsort((np.arange(len(li)), pn, li))  # query order, previous before next, walking order inside
            Q.append(li[o] + a), R.append(rr[o] + sa), D.append(dd[o])
        return _t(np.concatenate(Q)), _t(np.concatenate(R)), _t(np.concatenate(D))


def to_device_i64(a, device):
    return _t(np.ascontiguousarray(a, dtype=np.int64))


def _ranges(iv, seg_ranges):
    if seg_ranges is None:
        seg_ranges = [(i, i + 1) for i in range(iv.n_seg)]
    return [(iv.seg_off[a], iv.seg_off[b]) for a, b in seg_ranges]


def cluster(iv, seg_ranges=None, slack=0):
    rows, ids, base = [np.zeros(0, np.int64)], [np.zeros(0, np.int64)], 0
    for a, b in _ranges(iv, seg_ranges):
        if a == b:
            continue
        o = np.lexsort([np.arange(a, b), iv.e[a:b], iv.s[a:b]])
        c = orc.annotate_clusters(iv.s[a:b][o], iv.e[a:b][o], slack)
        rows.append(o + a)
        ids.append(c + base)
        base += c.max()
    return _t(np.concatenate(rows)), _t(np.concatenate(ids)), _t(np.asarray([base]))


def merge(iv, seg_ranges=None, slack=0, want_count=True):
    S, En, N, G = ([np.zeros(0, np.int64)] for _ in range(4))
    for g, (a, b) in enumerate(_ranges(iv, seg_ranges)):
        if a == b:
            continue
        o = np.lexsort([iv.e[a:b], iv.s[a:b]])
        cs, ce, cn = orc.find_clusters(iv.s[a:b][o], iv.e[a:b][o], slack)
        S.append(cs), En.append(ce), N.append(cn), G.append(np.full(len(cs), g, np.int64))
    return _t(np.concatenate(S)), _t(np.concatenate(En)), _t(np.concatenate(N)), _t(np.concatenate(G).astype(np.int32))


def coverage_rle(iv, seg_ranges=None, weights=None):
    off, R, V = [0], [np.zeros(0, np.int64)], [np.zeros(0, np.float64)]
    w = None if weights is None else weights.numpy()
    for a, b in _ranges(iv, seg_ranges):
        r, v = (orc.coverage_rle(iv.s[a:b], iv.e[a:b], None if w is None else w[a:b]) if b > a
                else (np.zeros(0, np.int64), np.zeros(0)))
        R.append(r), V.append(v)
        off.append(off[-1] + len(r))
    return np.asarray(off, np.int64), _t(np.concatenate(R)), _t(np.concatenate(V))


def materialize_pairs(q_iv, s_iv, pair_q, pair_s, clip=False, want=("q_start", "q_end", "s_start", "s_end"),
                      q_row_offset=0, s_row_offset=0):
    qr, sr = pair_q.numpy() - q_row_offset, pair_s.numpy() - s_row_offset
    qs, qe, ss, se = q_iv.s[qr], q_iv.e[qr], s_iv.s[sr], s_iv.e[sr]
    out = {"overlap": np.minimum(qe, se) - np.maximum(qs, ss)}
    if clip:
        qs, qe = np.maximum(qs, ss), np.minimum(qe, se)
    out.update({"q_start": qs, "q_end": qe, "s_start": ss, "s_end": se})
    return {k: _t(out[k]) for k in want}


def split(iv, seg_ranges=None):
    off, S, E = [0], [], []
    for a, b in _ranges(iv, seg_ranges):
        fs, fe = orc.split_points(iv.s[a:b], iv.e[a:b]) if b > a else (np.zeros(0, np.int64),) * 2
        S.append(fs)
        E.append(fe)
        off.append(off[-1] + len(fs))
    cat = lambda x: np.concatenate(x) if x else np.zeros(0, np.int64)  # noqa: E731
    return np.asarray(off, np.int64), _t(cat(S)), _t(cat(E))


def overlap_pairs(s_iv, group_ranges, q, want_counts=False, capacity=None):
    return SubjectIndex(s_iv, group_ranges).pairs(q, IV_MODE_ALL, want_counts=want_counts)


def overlap_counts(s_iv, group_ranges, q, mode=IV_MODE_ALL):
    return SubjectIndex(s_iv, group_ranges).count(q, mode, for_emit=False)[0]
