"""Soak run of the randomised GPU parity sweep over many more seeds than the test-suite uses.
Usage: python tools/soak_fuzz.py [first_seed] [last_seed]"""
import sys
import time

sys.path.insert(0, ".")
from tests import test_gpu_fuzz as F

a = int(sys.argv[1]) if len(sys.argv) > 1 else 40
b = int(sys.argv[2]) if len(sys.argv) > 2 else 400
t = time.time()
for seed in range(a, b):
    F.test_fuzz_binary_ops.__wrapped__(seed) if hasattr(F.test_fuzz_binary_ops, "__wrapped__") else F.test_fuzz_binary_ops(seed)
    F.test_fuzz_unary_ops.__wrapped__(seed) if hasattr(F.test_fuzz_unary_ops, "__wrapped__") else F.test_fuzz_unary_ops(seed)
print("fuzz seeds %d..%d: binary + unary ops parity-green in %.1f s" % (a, b - 1, time.time() - t))


def large_case(seed):
    """lean count kernel + hit lists + emit on randomised inputs above the lean threshold: random key sizes, subject
    density, maximal lengths (long subjects -> coarse bins, dense bins), slack, coordinate offsets beyond 32 bits."""
    import numpy as np

    from pyranges_b200 import engine as E
    from tests.util import canon_pairs, gen_segments, oracle_counts, oracle_pairs

    rng = np.random.default_rng(seed)
    kq = int(rng.integers(1, 7))
    ks = int(rng.integers(1, 6))
    maxpos = int(rng.choice([3000, 200_000, 5_000_000, 400_000_000]))
    qlen = int(rng.choice([2, 100, 5000]))
    n_per = [int(rng.integers(0, 60_000)) for _ in range(kq)]
    n_per[int(rng.integers(0, kq))] += 70_000  # at least 16 lean blocks
    qs, qe, qoff = [], [], [0]
    for n in n_per:
        s = rng.integers(0, maxpos, n)
        qs.append(s)
        qe.append(s + rng.integers(1, qlen + 1, n))
        qoff.append(qoff[-1] + n)
    qs, qe, qoff = np.concatenate(qs).astype(np.int64), np.concatenate(qe).astype(np.int64), np.asarray(qoff, np.int64)
    ns_per, slen = int(rng.choice([50, 3000, 40_000])), int(rng.choice([10, 3000, 200_000]))
    # keep the oracle fast: about 30 hits per query at most (gen_segments makes 10 % of the subjects 5-20 x longer)
    while ns_per * 2.5 * slen > 30 * maxpos and ns_per > 50:
        ns_per //= 4
    while ns_per * 2.5 * slen > 30 * maxpos and slen > 10:
        slen //= 4
    ss, se, soff = gen_segments(seed + 1, ks, ns_per, maxpos, slen)
    shift = int(rng.choice([0, 0, 7_000_000_000]))
    qs, qe, ss, se = qs + shift, qe + shift, ss + shift, se + shift
    sg = rng.integers(-1, ks, kq).astype(np.int32)
    slack = int(rng.choice([0, 0, 40])) if shift == 0 else 0
    q = E.Queries(E.DeviceIntervals.from_numpy(qs, qe, qoff), sg, slack=slack)
    ix = E.SubjectIndex(E.DeviceIntervals.from_numpy(ss, se, soff))
    for mode, name in [(E.IV_MODE_ALL, "all"), (E.IV_MODE_ANY, "any")]:
        c, _, total = ix.count(q, mode)
        want = oracle_counts(qs, qe, qoff, sg, ss, se, soff, name, slack)
        assert np.array_equal(c.cpu().numpy(), want), (seed, name)
        assert int(total.item()) == int(want.sum())
    _, a, b = ix.pairs(q, E.IV_MODE_ALL, want_counts=False)
    wa, wb = canon_pairs(*oracle_pairs(qs, qe, qoff, sg, ss, se, soff, "all", slack))
    ga, gb = canon_pairs(a.cpu().numpy(), b.cpu().numpy())
    assert np.array_equal(ga, wa) and np.array_equal(gb, wb), seed
    for mode, name in [(E.IV_MODE_ANY, "first"), (E.IV_MODE_LAST, "last")]:
        _, a, b = ix.pairs(q, mode, want_counts=False)
        wa, wb = oracle_pairs(qs, qe, qoff, sg, ss, se, soff, name, slack)
        assert np.array_equal(a.cpu().numpy(), wa) and np.array_equal(b.cpu().numpy(), wb), (seed, name)
    return len(qs), len(ga)


t = time.time()
rows = pairs = 0
n_large = min(24, max(1, (b - a) // 10))
for seed in range(a, a + n_large):
    r, p = large_case(seed)
    rows, pairs = rows + r, pairs + p
    print("  large seed %d: %d rows, %d pairs, %.1f s so far" % (seed, r, p, time.time() - t), flush=True)
print("large cases: %d seeds, %d query rows, %d pairs parity-green in %.1f s" % (n_large, rows, pairs, time.time() - t))
