"""Synthetic "hg38-shaped" interval sets of the benchmark configs, modelled on
pr.random (pyranges/__init__.py:305-331): chromosome ~ Categorical(p proportional to size), strand uniform,
rows left in generation order inside a key.  Host-side numpy only."""
import numpy as np

HG38 = [("chr1", 248956422), ("chr2", 242193529), ("chr3", 198295559), ("chr4", 190214555), ("chr5", 181538259),
        ("chr6", 170805979), ("chr7", 159345973), ("chr8", 145138636), ("chr9", 138394717), ("chr10", 133797422),
        ("chr11", 135086622), ("chr12", 133275309), ("chr13", 114364328), ("chr14", 107043718),
        ("chr15", 101991189), ("chr16", 90338345), ("chr17", 83257441), ("chr18", 80373285), ("chr19", 58617616),
        ("chr20", 64444167), ("chr21", 46709983), ("chr22", 50818468), ("chrX", 156040895), ("chrY", 57227415)]
SIZES = np.asarray([s for _, s in HG38], dtype=np.int64)
STRANDS = ("+", "-")


def keys(stranded=True):
    """natsorted key list: (chrom, '+'), (chrom, '-') ... or chrom."""
    if stranded:
        return [(c, s) for c, _ in HG38 for s in STRANDS]
    return [c for c, _ in HG38]


class KeyedArrays:
    """starts/ends grouped by (chromosome, strand) key: segment k = rows [seg_off[k], seg_off[k+1])."""

    def __init__(self, starts, ends, key_id, n_keys):
        order = np.argsort(key_id, kind="stable")
        self.starts = np.ascontiguousarray(starts[order])
        self.ends = np.ascontiguousarray(ends[order])
        self.seg_off = np.zeros(n_keys + 1, np.int64)
        np.cumsum(np.bincount(key_id, minlength=n_keys), out=self.seg_off[1:])
        self.n = len(starts)
        self.n_keys = n_keys

    def chrom_ranges(self):
        """segment ranges of the per-chromosome groups (both strands of a chromosome are adjacent)."""
        return [(2 * c, 2 * c + 2) for c in range(self.n_keys // 2)]


def shard_chromosomes(world, rank):
    """chromosome ids owned by `rank`: size-balanced (LPT) assignment, both strands of a chromosome together"""
    from .sharding import lpt_assign

    a = lpt_assign(SIZES, world)
    return [c for c in range(len(SIZES)) if a[c] == rank]


def _split_counts(n_total):
    """rows per chromosome, proportional to size, summing to n_total exactly"""
    w = SIZES / SIZES.sum()
    c = np.floor(w * n_total).astype(np.int64)
    c[: int(n_total - c.sum())] += 1
    return c


def _chrom_for(rng, n_total, chroms):
    per = _split_counts(n_total)
    c = np.repeat(np.asarray(chroms, dtype=np.int64), per[chroms])
    return c[rng.permutation(len(c))]


def _chrom(rng, n):
    return rng.choice(len(SIZES), size=n, p=SIZES / SIZES.sum())


def reads(n, seed=1, length=100, min_len=None, max_len=None, chroms=None):
    """n reads: Start ~ U[0, size-len), fixed length (or U{min_len..max_len}).
    chroms: only the rows of these chromosome ids of an n-row genome-wide set (a rank's shard)."""
    rng = np.random.default_rng(seed)
    c = _chrom(rng, n) if chroms is None else _chrom_for(rng, n, chroms)
    n = len(c)
    strand = rng.integers(0, 2, n)
    ln = np.full(n, length, np.int64) if min_len is None else rng.integers(min_len, max_len + 1, n)
    s = np.floor(rng.random(n) * (SIZES[c] - ln)).astype(np.int64)
    return KeyedArrays(s, s + ln, c * 2 + strand, 2 * len(SIZES))


def annotations(n, seed=2, chroms=None):
    """10 % genes (log-normal length, 0.5 kb .. 2 Mb), 90 % exons nested inside a random gene.
    chroms: only the genes (and their exons) on these chromosome ids of an n-row genome-wide set."""
    rng = np.random.default_rng(seed)
    ng = max(1, n // 10)
    gc = _chrom(rng, ng) if chroms is None else _chrom_for(rng, ng, chroms)
    if chroms is not None:
        n = int(round(n * len(gc) / ng))
    ng = len(gc)
    ne = n - ng
    gstrand = rng.integers(0, 2, ng)
    glen = np.clip(np.floor(rng.lognormal(np.log(25000.0), 1.2, ng)), 500, 2000000).astype(np.int64)
    glen = np.minimum(glen, SIZES[gc] - 1)
    gs = np.floor(rng.random(ng) * (SIZES[gc] - glen)).astype(np.int64)
    parent = rng.integers(0, ng, ne)
    elen = np.clip(np.floor(rng.lognormal(np.log(150.0), 0.8, ne)), 20, 10000).astype(np.int64)
    elen = np.minimum(elen, glen[parent])
    es = gs[parent] + np.floor(rng.random(ne) * (glen[parent] - elen + 1)).astype(np.int64)
    s = np.concatenate([gs, es])
    e = np.concatenate([gs + glen, es + elen])
    key = np.concatenate([gc * 2 + gstrand, gc[parent] * 2 + gstrand[parent]])
    perm = rng.permutation(n)  # interleave genes and exons: rows are unsorted
    return KeyedArrays(s[perm], e[perm], key[perm], 2 * len(SIZES))


# ---- per-chromosome generators (bench.py) ---------------------------------------------------------------------------
# The same distributions as reads() / annotations() above, but every chromosome draws from its own seeded stream and
# holds an exact size-proportional share of the rows.  A rank that owns a subset of the chromosomes therefore generates
# exactly the rows a single-GPU run holds for them: the genome-wide set is FIXED whatever the number of GPUs (strong
# scaling), and order-invariant checksums of the results must agree between N = 1, 2, 4 and 8.
def _keyed_from_chroms(parts, chroms):
    """parts: per chromosome (starts, ends, strand); -> KeyedArrays whose key k = (chroms[k // 2], strand k % 2)"""
    S, E, off = [], [], [0]
    for s, e, strand in parts:
        plus = strand == 0
        for m in (plus, ~plus):
            S.append(s[m])
            E.append(e[m])
            off.append(off[-1] + int(m.sum()))
    ka = KeyedArrays.__new__(KeyedArrays)
    ka.starts = np.ascontiguousarray(np.concatenate(S)) if S else np.zeros(0, np.int64)
    ka.ends = np.ascontiguousarray(np.concatenate(E)) if E else np.zeros(0, np.int64)
    ka.seg_off = np.asarray(off, np.int64)
    ka.n = int(off[-1])
    ka.n_keys = 2 * len(chroms)
    ka.chrom_ids = np.asarray(chroms, np.int64)
    ka.global_key = (np.repeat(ka.chrom_ids, 2) * 2 + np.tile(np.arange(2), len(chroms))).astype(np.int64)
    return ka


def reads_by_chrom(n_total, seed=1, length=100, min_len=None, max_len=None, chroms=None):
    """the rows of chromosomes `chroms` (default: all) of an n_total-row genome-wide read set"""
    chroms = list(range(len(SIZES))) if chroms is None else list(chroms)
    per = _split_counts(n_total)
    parts = []
    for c in chroms:
        rng = np.random.default_rng([seed, c])
        n = int(per[c])
        strand = rng.integers(0, 2, n)
        ln = np.full(n, length, np.int64) if min_len is None else rng.integers(min_len, max_len + 1, n)
        s = np.floor(rng.random(n) * (SIZES[c] - ln)).astype(np.int64)
        parts.append((s, s + ln, strand))
    return _keyed_from_chroms(parts, chroms)


def annotations_by_chrom(n_total, seed=2, chroms=None):
    """the rows of chromosomes `chroms` of an n_total-row genome-wide gene/exon annotation (10 % genes)"""
    chroms = list(range(len(SIZES))) if chroms is None else list(chroms)
    ng_all, ne_all = _split_counts(max(1, n_total // 10)), _split_counts(n_total - max(1, n_total // 10))
    parts = []
    for c in chroms:
        rng = np.random.default_rng([seed, c])
        ng, ne = max(1, int(ng_all[c])), int(ne_all[c])
        gstrand = rng.integers(0, 2, ng)
        glen = np.clip(np.floor(rng.lognormal(np.log(25000.0), 1.2, ng)), 500, 2000000).astype(np.int64)
        glen = np.minimum(glen, SIZES[c] - 1)
        gs = np.floor(rng.random(ng) * (SIZES[c] - glen)).astype(np.int64)
        parent = rng.integers(0, ng, ne)
        elen = np.clip(np.floor(rng.lognormal(np.log(150.0), 0.8, ne)), 20, 10000).astype(np.int64)
        elen = np.minimum(elen, glen[parent])
        es = gs[parent] + np.floor(rng.random(ne) * (glen[parent] - elen + 1)).astype(np.int64)
        s = np.concatenate([gs, es])
        e = np.concatenate([gs + glen, es + elen])
        strand = np.concatenate([gstrand, gstrand[parent]])
        perm = rng.permutation(ng + ne)  # interleave genes and exons: rows are unsorted
        parts.append((s[perm], e[perm], strand[perm]))
    return _keyed_from_chroms(parts, chroms)


def sorted_within_keys(ka):
    """the same rows, every key sorted by (Start, End): position-sorted input (BAM order)"""
    key = np.repeat(np.arange(ka.n_keys), np.diff(ka.seg_off))
    o = np.lexsort([ka.ends, ka.starts, key])
    out = KeyedArrays.__new__(KeyedArrays)
    out.__dict__.update(ka.__dict__)
    out.starts, out.ends = np.ascontiguousarray(ka.starts[o]), np.ascontiguousarray(ka.ends[o])
    return out
