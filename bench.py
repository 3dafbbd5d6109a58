#!/usr/bin/env python
"""bench.py -- join overlap pairs/s and count_overlaps queries/s on synthetic hg38-shaped intervals, 1/2/4/8 B200.

Headline workload (BASELINE.json configs[2], the one the metric's "1/2/4/8 B200" is quoted on): 100 M reads (100 bp) x
1 M gene/exon annotations, strandedness=None (reads keyed by (chromosome, strand), every key joined against both strands
of its chromosome).  `--workload c2` selects configs[1] (10 M x 200 k).  One "step" = one full join pass: subject index
build (the reference rebuilds its NCLS on every call, pyranges/methods/join.py:12) -> all (query row, subject row)
int64 pairs in query order.  Extra keys of the same JSON line: count_overlaps and overlap(how="first") on the same
inputs, the C4 operations (nearest upstream/same, cluster, merge on 50 M intervals, unsorted and sorted) and C5 (to_rle
of 200 M reads), each with its own roofline entry (SURVEY.md section 8(d) bytes).

  python bench.py [--gpus N --steps K --warmup W]             this repo's CUDA path
  python bench.py --impl reference [...]                       CPU restatement of ncls on the host cores

N > 1 (torchrun): STRONG scaling of the fixed genome-wide set: chromosomes are assigned to ranks by a size-balanced
(LPT) rule, every rank generates and processes exactly the rows a single-GPU run holds for its chromosomes, no
data-path collective; the ranks exchange only per-key result counts (global offsets of the concatenated result /
cluster-id offsets / run counts).  Order-invariant 64-bit checksums of the results are printed: they must agree
between N = 1, 2, 4 and 8.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

WORKLOADS = {"c2": (10_000_000, 200_000), "c3": (100_000_000, 1_000_000)}
C4_ROWS, C5_ROWS = 50_000_000, 200_000_000


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--workload", default="c3", choices=sorted(WORKLOADS))
    ap.add_argument("--reads", type=int, default=0, help="override the read count of the workload")
    ap.add_argument("--annot", type=int, default=0, help="override the annotation count of the workload")
    ap.add_argument("--ops", default="all", help="all | none | comma list of count,overlap,sorted,nearest,cluster,merge,rle,api")
    ap.add_argument("--ops-scale", type=float, default=1.0, help="scale of the C4 / C5 row counts (tests)")
    ap.add_argument("--ops-steps", type=int, default=5)
    ap.add_argument("--path", default="auto", choices=["auto", "fused", "twopass"], help="join kernels (auto: fused)")
    ap.add_argument("--cpu-seconds", type=float, default=15.0, help="budget of the cpu_baseline sample")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-graph", action="store_true", help="multi-GPU: launch the step eagerly instead of replaying a CUDA graph")
    a = ap.parse_args()
    r, s = WORKLOADS[a.workload]
    a.reads = a.reads or r
    a.annot = a.annot or s
    want = {"count", "overlap", "sorted", "nearest", "cluster", "merge", "rle", "api"}
    a.ops = want if a.ops == "all" else (set() if a.ops == "none" else set(a.ops.split(",")))
    return a


# ---------------------------------------------------------------------------------------------------
class ClockSampler:
    """nvidia-smi sampling of SM clocks / throttle reasons during the timed region."""

    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.rows = []
        self.proc = None
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(gpu_index), "--query-gpu=" + self.Q,
                                          "--format=csv,noheader,nounits", "-lms", "100"], stdout=subprocess.PIPE,
                                         stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append((time.perf_counter(), [x.strip() for x in line.split(",")]))

    def summary(self, t0, t1):
        if self.proc is not None:
            self.proc.terminate()
        rows = [r for t, r in self.rows if t0 <= t <= t1 and len(r) >= 7] or [r for _, r in self.rows if len(r) >= 7]
        if not rows:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["unavailable"]}
        sm = sorted(float(r[0]) for r in rows)
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = [n for i, n in enumerate(names) if any(r[3 + i].lower().startswith("active") for r in rows)]
        return {"sm_mhz": sm[len(sm) // 2], "sm_max_mhz": float(rows[0][1]), "reasons": reasons,
                "samples": len(rows)}


def measured_peak_gbs():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md)"


def ncu_traffic(kernel, workload):
    """DRAM bytes per launch of `kernel` on `workload` from a committed `ncu --set full` capture of this command
    (profiles/traffic.json: {"<workload>/<kernel>": {"bytes": ..., "source": "profiles/<file>"}}), else None."""
    try:
        with open(os.path.join(ROOT, "profiles", "traffic.json")) as f:
            e = json.load(f).get("%s/%s" % (workload, kernel))
        return (e["bytes"], e["source"]) if e else (None, None)
    except Exception:
        return None, None


def shard(world, rank):
    from pyranges_b200 import synthetic as S

    return None if world == 1 else S.shard_chromosomes(world, rank)


def make_workload(args, rank, world=1):
    from pyranges_b200 import synthetic as S

    chroms = shard(world, rank)
    reads = S.reads_by_chrom(args.reads, seed=1, chroms=chroms)
    annot = S.annotations_by_chrom(args.annot, seed=2, chroms=chroms)
    seg_group = (np.arange(reads.n_keys) // 2).astype(np.int32)  # strandedness=None: partner = whole chromosome
    return reads, annot, seg_group


# ---- order-invariant checksums (identical for every sharding of the same genome-wide set) ------------------------
_M1, _M2, _M3 = -7046029254386353131, -4417276706812531889, -4658895280553007687  # odd 64-bit constants, as int64


def _mix(x):
    x = x ^ (x >> 29)
    x = x * _M3
    return x ^ (x >> 32)


def key_local(torch, rows, seg_off_dev, global_key_dev):
    """row of the concatenated local layout -> (global key id << 32) | row inside its key"""
    k = torch.searchsorted(seg_off_dev, rows, right=True) - 1
    return (global_key_dev[k] << 32) | (rows - seg_off_dev[k])


def checksum_pairs(torch, a, b, q_meta, s_meta, chunk=1 << 26):
    tot = torch.zeros((), dtype=torch.int64, device=a.device)
    for i in range(0, a.numel(), chunk):
        qa = key_local(torch, a[i:i + chunk], *q_meta)
        sb = key_local(torch, b[i:i + chunk], *s_meta)
        tot += _mix(qa * _M1 + sb * _M2).sum()
    return tot


def checksum_rows(torch, values, meta, chunk=1 << 26):
    """checksum of one int64 value per row (counts, flags, cluster ids ...)"""
    tot = torch.zeros((), dtype=torch.int64, device=values.device)
    for i in range(0, values.numel(), chunk):
        rows = torch.arange(i, min(i + chunk, values.numel()), device=values.device)
        tot += _mix(key_local(torch, rows, *meta) * _M1 + values[i:i + chunk] * _M2).sum()
    return tot


# ---------------------------------------------------------------------------------------------------
def oracle_join(reads, annot, seg_group, frac=1.0, threads=1, budget_s=None):
    """CPU restatement of the reference path: per chromosome NCLS build + all_overlaps_both of the
    chromosome's two read keys (first `frac` of each key's rows).  -> (pairs, queries, seconds)"""
    from concurrent.futures import ThreadPoolExecutor

    from oracle import oracle as orc

    groups = annot.chrom_ranges()

    def one(g):
        a, b = annot.seg_off[groups[g][0]], annot.seg_off[groups[g][1]]
        t = orc.NCLS(annot.starts[a:b], annot.ends[a:b], np.arange(a, b, dtype=np.int64))
        pairs = nq = 0
        for k in np.flatnonzero(seg_group == g):
            q0, q1 = reads.seg_off[k], reads.seg_off[k + 1]
            q1 = q0 + int((q1 - q0) * frac)
            x, _ = t.all_overlaps_both(reads.starts[q0:q1], reads.ends[q0:q1], np.arange(q0, q1, dtype=np.int64))
            pairs += len(x)
            nq += q1 - q0
        return pairs, nq

    t0 = time.perf_counter()
    tot_p = tot_q = 0
    done = 0
    if threads > 1:
        with ThreadPoolExecutor(threads) as ex:
            for p, q in ex.map(one, range(len(groups))):
                tot_p += p
                tot_q += q
        done = len(groups)
    else:
        for g in range(len(groups)):
            p, q = one(g)
            tot_p += p
            tot_q += q
            done += 1
            if budget_s is not None and time.perf_counter() - t0 > budget_s:
                break
    return tot_p, tot_q, time.perf_counter() - t0, done


def run_reference(args, rank, world):
    """--impl reference: the CPU path on the host cores (oracle restatement of ncls; the ncls wheel itself
    is not installable here)."""
    if rank != 0:
        return
    from oracle import oracle as orc

    orc.build()
    reads, annot, seg_group = make_workload(args, 0, 1)
    cores = os.cpu_count() or 1
    threads = min(cores, len(annot.chrom_ranges()))
    # calibrate the per-step sample so that the whole run stays within a few minutes
    f0 = min(1.0, 500_000.0 / max(reads.n, 1))
    p, q, dt, _ = oracle_join(reads, annot, seg_group, frac=f0, threads=threads)
    budget = 120.0 / max(args.steps + args.warmup, 1)  # seconds per step
    frac = float(min(1.0, f0 * min(budget, 6.0) / max(dt, 1e-3)))
    for _ in range(args.warmup):
        oracle_join(reads, annot, seg_group, frac=frac, threads=threads)
    t0 = time.perf_counter()
    pairs = queries = 0
    for _ in range(args.steps):
        p, q, _, _ = oracle_join(reads, annot, seg_group, frac=frac, threads=threads)
        pairs += p
        queries += q
    dt = time.perf_counter() - t0
    val = pairs / dt
    sample = ("first %.2f%% of the rows of every read key (%d queries/step) x all %d annotations (NCLS rebuilt per step), "
              "%d threads over chromosomes" % (100 * frac, queries // max(args.steps, 1), annot.n, threads))
    # the reference's own Python layer (pandas glue included) over the restated native layer, where /root/reference exists
    layer = None
    try:
        sys.path.insert(0, os.path.join(ROOT, "tools"))
        import reference_baseline as rb

        if rb.available():
            layer = rb.run(2_000_000, 40_000, min(cores, 16))
            layer["note"] = ("UNMODIFIED /root/reference pyranges layer, builder-sized sample (2 M x 40 k rows): end-to-end "
                             "gr.join / gr.count_overlaps at nb_cpu=1, per-key fork pool, index only")
    except Exception as ex:  # never take the baseline line down
        layer = {"error": repr(ex)}
    line = {
        "impl": "reference", "metric": "join overlap pairs/s", "value": val, "unit": "pairs/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * dt / max(args.steps, 1),
        "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "int64", "data": "synthetic",
        "config": workload_config(args, args.gpus),
        "count_overlaps_queries_per_s": queries / dt,
        "cpu_baseline": {"value": val, "unit": "pairs/s", "cores": threads, "kind": "port", "sample": sample},
        "e2e": {"value": val, "unit": "pairs/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "pyranges_layer": layer,
    }
    print(json.dumps(line), flush=True)


def workload_config(args, n_gpus):
    return {"workload": "%s: synthetic hg38-shaped %d reads (100 bp) x %d gene/exon annotations, join (+ count_overlaps, "
                        "overlap), strandedness=None, genome-wide set fixed for every N" % (args.workload, args.reads, args.annot),
            "reads": args.reads, "annotations": args.annot, "keys": 48,
            "l2": "inputs (%.0f MB of int64 query columns over %d GPU(s)) exceed the 126 MB L2; no flush" % (
                16 * args.reads / 1e6, n_gpus),
            "index_build": "inside the timed region",
            "parallelism": "chromosomes sharded over %d GPU(s) by size (LPT), strong scaling" % n_gpus}


# ---------------------------------------------------------------------------------------------------
def run_b200(args, rank, world, local_rank):
    import torch
    import torch.distributed as dist

    from pyranges_b200 import _cabi as cabi
    from pyranges_b200 import engine as E
    from pyranges_b200 import hostapi
    from pyranges_b200 import synthetic as S

    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    L = cabi.lib()
    peak, peak_src = measured_peak_gbs()
    ev = lambda: torch.cuda.Event(enable_timing=True)  # noqa: E731

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def reduce_max(x):
        if world == 1:
            return float(x)
        t = torch.tensor([x], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    def reduce_sum_i64(x):
        t = x.clone().reshape(-1) if isinstance(x, torch.Tensor) else torch.tensor([int(x)], dtype=torch.int64, device=dev)
        if world > 1:
            dist.all_reduce(t)
        return int(t[0].item()) if t.numel() == 1 else t

    def timed(fn, steps, warmup):
        """W untimed + K timed steps between barriers; device time (CUDA events), max over ranks; -> (ms total, launches)"""
        for _ in range(warmup):
            fn()
        barrier()
        a, b = ev(), ev()
        n0 = L.iv_launch_count()
        a.record()
        for _ in range(steps):
            fn()
        b.record()
        barrier()
        return reduce_max(a.elapsed_time(b)), L.iv_launch_count() - n0

    def meta(ka, iv):
        return iv.seg_off_dev, torch.from_numpy(ka.global_key).to(dev)

    def roof(alg_bytes, ms):
        ach = alg_bytes / (ms / 1e3) / 1e9
        return {"algorithmic_bytes": int(alg_bytes), "achieved": ach, "frac": ach / peak}

    # ---- the headline workload ------------------------------------------------------------------
    reads, annot, seg_group = make_workload(args, rank, world)
    groups = annot.chrom_ranges()
    q_iv = E.DeviceIntervals.from_numpy(reads.starts, reads.ends, reads.seg_off, dev)
    s_iv = E.DeviceIntervals.from_numpy(annot.starts, annot.ends, annot.seg_off, dev)
    s_iv.bounds  # ingest-time metadata (per-key min Start / max End), like the chromosome table
    nq, ns = reads.n, annot.n
    nq_all, ns_all = args.reads, args.annot
    q_meta, s_meta = meta(reads, q_iv), meta(annot, s_iv)

    join = E.JoinStep(s_iv, groups, q_iv, seg_group, path=args.path, seg_counts=world > 1)
    # one untimed pass sizes the output buffers (the step launches its emit into buffers of the previous step's size and
    # checks the device-side total afterwards: no host synchronisation inside a step)
    a, b, p_local = join.run_sync()
    pairs_all = reduce_sum_i64(p_local)
    ck_pairs = reduce_sum_i64(checksum_pairs(torch, a, b, q_meta, s_meta))
    del a, b

    sampler = ClockSampler(local_rank) if rank == 0 else None
    time.sleep(0.3)
    t0 = time.perf_counter()
    join.phase_events = []
    graph = world > 1 and not args.no_graph
    if graph:
        # per-rank steps of a multi-GPU run are short (a few hundred microseconds): the whole step is replayed as one
        # CUDA graph so that eight Python launch loops do not pace the GPUs; the per-phase times of the roofline entry
        # are then taken from an eager loop of the same step right after the timed region
        join.capture()
    if world > 1:
        # the only communication of the path: per-key pair counts (global offsets of the concatenated result, keys are
        # disjoint over ranks so the sum is the gather).  Every step leaves its counts in a row of a device table;
        # ONE all-reduce of the table closes the timed region -- no collective and no host wait inside a step.
        gkq = torch.from_numpy(reads.global_key).to(dev)
        rows = args.steps + args.warmup
        key_pairs = torch.zeros(rows, 48, dtype=torch.int64, device=dev)
        it = [0]

        def join_step():
            join.step()
            key_pairs[it[0] % rows].index_copy_(0, gkq, join.seg_counts)
            it[0] += 1
            if it[0] == rows:
                dist.all_reduce(key_pairs)
    else:
        join_step = join.step
    join_ms, launches = timed(join_step, args.steps, args.warmup)
    join.check()  # every step's device-side total fitted its buffers (else the run is invalid: raises)
    if graph:
        launches += args.steps * join.graph_launches  # kernels inside the replayed graphs of the timed steps
        eager = E.JoinStep(s_iv, groups, q_iv, seg_group, path=args.path)
        eager.run_sync()
        eager.phase_events = []
        for _ in range(args.steps):
            eager.step()
        eager.check()
        join.phase_events = eager.phase_events
    if world > 1:  # the exchanged per-key counts of the last step add up to the global pair count
        assert int(key_pairs[-1].sum().item()) == pairs_all, "per-key pair counts do not add up"
    t1 = time.perf_counter()
    clocks = sampler.summary(t0, t1) if sampler else None
    value = pairs_all * args.steps / (join_ms / 1e3)
    torch.cuda.synchronize()
    phases = join.phase_summary(args.steps)  # per-phase device times of the timed steps (events around the C calls)
    join.phase_events = None

    # roofline of the dominant phase.  Algorithmic bytes (no credit for temporaries): every input
    # column the phase must read once + every output it must write once.
    alg = {"index_build": 24 * ns, "join": 16 * nq + 16 * p_local, "count": 16 * nq, "emit": 16 * p_local}
    dom = max(phases, key=lambda k: phases[k])
    dom_ms = phases[dom]
    # (the committed capture is of the single-GPU launch: no figure for the smaller per-rank launches of N > 1)
    traffic, traffic_src = ncu_traffic(join.kernel_of(dom), args.workload) if world == 1 else (None, None)
    pipe = 24 * nq + 24 * ns + 16 * p_local  # the join's algorithmic bytes (labels included)
    pipe_req = 16 * nq + 16 * ns + 16 * p_local  # what this implementation has to move (labels = row numbers)
    step_ms_local = join_ms / args.steps
    roofline = {"bound": "hbm", "kernel": join.kernel_of(dom), "phase": dom,
                "achieved": alg[dom] / (dom_ms / 1e3) / 1e9, "peak": peak, "unit": "GB/s",
                "frac": alg[dom] / (dom_ms / 1e3) / 1e9 / peak, "traffic": traffic, "traffic_source": traffic_src,
                "peak_source": peak_src, "algorithmic_bytes_per_launch": int(alg[dom]),
                "rank0_phase_ms": phases,
                "phase_source": ("eager loop of the same step after the timed region (the timed steps replay one CUDA graph)"
                                 if graph else "CUDA events inside the timed region"),
                "rank0_phase_achieved_gbs": {k: alg[k] / (m / 1e3) / 1e9 for k, m in phases.items() if m > 0},
                "pipeline": {"bytes_8d": int(pipe), "bytes_no_labels": int(pipe_req), "ms": step_ms_local,
                             "achieved_8d": pipe / (step_ms_local / 1e3) / 1e9, "frac_8d": pipe / (step_ms_local / 1e3) / 1e9 / peak,
                             "frac_no_labels": pipe_req / (step_ms_local / 1e3) / 1e9 / peak,
                             "note": "rank 0's bytes over the max-over-ranks step time, index build included"}}

    ops = {}
    # ---- count_overlaps / overlap(how="first") on the same inputs -------------------------------------
    if "count" in args.ops:
        cnt = E.CountStep(s_iv, groups, q_iv, seg_group, E.IV_MODE_ALL)
        c = cnt.step()
        ck = reduce_sum_i64(checksum_rows(torch, c, q_meta))
        ms, _ = timed(cnt.step, args.steps, args.warmup)
        ms /= args.steps
        ops["count_overlaps"] = dict(roof(16 * nq + 16 * ns + 8 * nq, ms), ms=ms, queries_per_s=nq_all / (ms / 1e3),
                                     checksum=ck, note="index build + counts, int64 NumberOverlaps column")
        del c
    if "overlap" in args.ops:
        ov = E.OverlapFirstStep(s_iv, groups, q_iv, seg_group)
        rows = ov.run_sync()
        k_all = reduce_sum_i64(rows.numel())
        ck = reduce_sum_i64(_mix(key_local(torch, rows, *q_meta) * _M1).sum())
        ms, _ = timed(ov.step, args.steps, args.warmup)
        ov.check()
        ms /= args.steps
        ops["overlap_first"] = dict(roof(16 * nq + 16 * ns + 8 * rows.numel(), ms), ms=ms, queries_per_s=nq_all / (ms / 1e3),
                                    rows_out=k_all, checksum=ck, note="index build + surviving self rows (how='first')")
        del rows
    if "sorted" in args.ops:
        # the same reads in position order inside every key (BAM order): the record gathers coalesce
        o1 = torch.argsort(q_iv.end, stable=True)
        key = torch.repeat_interleave(torch.arange(reads.n_keys, device=dev), torch.from_numpy(np.diff(reads.seg_off)).to(dev))
        o2 = torch.argsort(((key << 32) | q_iv.start)[o1], stable=True)
        o = o1[o2]
        qs_iv = E.DeviceIntervals(q_iv.start[o], q_iv.end[o], reads.seg_off, seg_off_dev=q_iv.seg_off_dev)
        del o, o1, o2, key
        js = E.JoinStep(s_iv, groups, qs_iv, seg_group, path=args.path)
        _, _, ps = js.run_sync()
        assert reduce_sum_i64(ps) == pairs_all
        ms, _ = timed(js.step, args.steps, args.warmup)
        js.check()
        ms /= args.steps
        ops["join_position_sorted"] = dict(roof(24 * nq + 24 * ns + 16 * p_local, ms), ms=ms,
                                           pairs_per_s=pairs_all / (ms / 1e3), note="8(d) join bytes over the whole step")
        del js, qs_iv

    # ---- end to end through the host-facing call: pinned host arrays in, host index arrays out ----
    e2e = None
    if not args.no_e2e:
        hq_s, hq_e = torch.from_numpy(reads.starts).pin_memory(), torch.from_numpy(reads.ends).pin_memory()
        hs_s, hs_e = torch.from_numpy(annot.starts).pin_memory(), torch.from_numpy(annot.ends).pin_memory()
        cap = int(p_local * 1.05) + 4096
        out = (torch.empty(cap, dtype=torch.int64).pin_memory(), torch.empty(cap, dtype=torch.int64).pin_memory())

        def e2e_step():
            nb = hostapi.NCLSBatch(hs_s, hs_e, annot.seg_off, groups, device=dev)
            x, _ = nb.all_overlaps_both(hq_s, hq_e, reads.seg_off, seg_group, out=out)
            return x.numel()

        e_steps = max(3, min(args.steps, 10))
        for _ in range(max(1, min(args.warmup, 3))):
            e2e_step()
        barrier()
        te = time.perf_counter()
        pe = 0
        for _ in range(e_steps):
            pe += e2e_step()
        barrier()
        dt = reduce_max(time.perf_counter() - te)
        pe = reduce_sum_i64(pe)
        e2e = {"value": pe / dt, "unit": "pairs/s", "h2d_bytes_per_step": 16 * nq + 16 * ns,
               "d2h_bytes_per_step": 16 * p_local, "ms_per_step": 1e3 * dt / e_steps, "steps": e_steps,
               "call": "pyranges_b200.hostapi.NCLSBatch(starts, ends, ...).all_overlaps_both(...) on pinned host int64 "
                       "arrays (bytes are rank 0's)"}
        del hq_s, hq_e, hs_s, hs_e, out

    # ---- CPU baseline beside it (rank 0, N = 1 only) ------------------------------------------------
    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        from oracle import oracle as orc

        orc.build()
        cp = cq = 0
        cdt = 0.0
        passes = 0
        while cdt < args.cpu_seconds and passes < 50:
            x, y, d, done = oracle_join(reads, annot, seg_group, frac=1.0, threads=1, budget_s=args.cpu_seconds - cdt)
            cp, cq, cdt, passes = cp + x, cq + y, cdt + d, passes + 1
        cpu = {"value": cp / cdt, "unit": "pairs/s", "cores": 1, "kind": "port",
               "queries_per_s": cq / cdt,
               "sample": "%d pass(es), the last over the first %d chromosomes (%d reads x their annotations in all): NCLS "
                         "build + all_overlaps_both per chromosome, %.1f s on 1 core" % (passes, done, cq, cdt)}

    join_path = join.path
    del join, q_iv, s_iv
    torch.cuda.empty_cache()

    # ---- C4: nearest (strandedness="same", how="upstream"), cluster, merge on 50 M intervals ----------
    chroms = shard(world, rank)
    if args.ops & {"nearest", "cluster", "merge"}:
        n4 = int(C4_ROWS * args.ops_scale)
        iv4 = S.reads_by_chrom(n4, seed=3, min_len=50, max_len=500, chroms=chroms)
        ann = S.annotations_by_chrom(int(1_000_000 * args.ops_scale), seed=2, chroms=chroms)
        q4 = E.DeviceIntervals.from_numpy(iv4.starts, iv4.ends, iv4.seg_off, dev)
        s4 = E.DeviceIntervals.from_numpy(ann.starts, ann.ends, ann.seg_off, dev)
        s4.bounds
        q4.bounds
        m4 = meta(iv4, q4)
        if "nearest" in args.ops:
            # partner = same key; '+' keys look upstream = previous, '-' keys = next (pyranges/methods/nearest.py:85-90)
            sg = np.arange(iv4.n_keys, dtype=np.int32)
            modes = np.asarray([E.IV_NEAREST_PREVIOUS if k % 2 == 0 else E.IV_NEAREST_NEXT for k in range(iv4.n_keys)], np.int32)
            qq = E.Queries(q4, sg, seg_mode=modes)

            def nearest_step():
                return E.SubjectIndex(s4, None, ends_perm=True).nearest(qq, overlap=True)

            row, dist_ = nearest_step()
            ck = reduce_sum_i64(checksum_rows(torch, dist_, m4))
            ms, _ = timed(nearest_step, args.ops_steps, 2)
            ms /= args.ops_steps
            ops["nearest_upstream_same"] = dict(roof(24 * iv4.n + 24 * ann.n + 16 * iv4.n, ms), ms=ms, rows=n4,
                                                rows_per_s=n4 / (ms / 1e3), checksum=ck,
                                                note="index build (with end permutation) + iv_nearest; checksum over Distance")
            del row, dist_, qq
        variants = [("unsorted", q4)]
        if args.ops & {"cluster", "merge"}:
            o1 = torch.argsort(q4.end, stable=True)
            key = torch.repeat_interleave(torch.arange(iv4.n_keys, device=dev), torch.from_numpy(np.diff(iv4.seg_off)).to(dev))
            o = o1[torch.argsort(((key << 32) | q4.start)[o1], stable=True)]
            q4s = E.DeviceIntervals(q4.start[o], q4.end[o], iv4.seg_off, bounds=q4.bounds, seg_off_dev=q4.seg_off_dev)
            del o, o1, key
            variants.append(("sorted", q4s))
        key_counts = torch.zeros(48, dtype=torch.int64, device=dev)
        gk = torch.from_numpy(iv4.global_key).to(dev)
        for name, ivx in variants:
            if "cluster" in args.ops:
                def cluster_step():
                    row, cid, nc = E.cluster(ivx)
                    if world > 1:
                        # cluster ids are made globally unique in natsorted key order (pyranges_main.py:1210-1223): the
                        # ranks exchange their per-key cluster counts (keys are disjoint: the sum is the gather)
                        key_counts.zero_()
                        last = cid[ivx.seg_off_dev[1:] - 1] if ivx.n else cid
                        key_counts[gk] = last - torch.cat([last.new_zeros(1), last[:-1]])
                        dist.all_reduce(key_counts)
                    return row, cid, nc

                row, cid, nc = cluster_step()
                ncl = reduce_sum_i64(nc)
                ms, _ = timed(cluster_step, args.ops_steps, 2)
                ms /= args.ops_steps
                ops["cluster_" + name] = dict(roof(24 * ivx.n, ms), ms=ms, rows=n4, rows_per_s=n4 / (ms / 1e3), clusters=ncl)
                del row, cid
            if "merge" in args.ops:
                def merge_step():
                    return E.merge(ivx)

                st, en, cnt_, grp = merge_step()
                m_loc = st.numel()
                ck = reduce_sum_i64(_mix(st * _M1 + en * _M2 + cnt_ * _M3).sum())
                ms, _ = timed(merge_step, args.ops_steps, 2)
                ms /= args.ops_steps
                ops["merge_" + name] = dict(roof(16 * ivx.n + 24 * m_loc, ms), ms=ms, rows=n4, rows_per_s=n4 / (ms / 1e3),
                                            merged=reduce_sum_i64(m_loc), checksum=ck)
                del st, en, cnt_, grp
        del variants, q4, s4
        torch.cuda.empty_cache()

    # ---- C5: per-chromosome coverage RLE (to_rle) of 200 M reads -------------------------------------------
    if "rle" in args.ops:
        n5 = int(C5_ROWS * args.ops_scale)
        iv5 = S.reads_by_chrom(n5, seed=4, chroms=chroms)
        r5 = E.DeviceIntervals.from_numpy(iv5.starts, iv5.ends, iv5.seg_off, dev)
        r5.bounds
        gk5 = torch.from_numpy(iv5.global_key).to(dev)
        run_counts = torch.zeros(48, dtype=torch.int64, device=dev)

        def rle_step():
            off, runs, vals = E.coverage_rle(r5)
            if world > 1:  # per-key run counts -> offsets of the concatenated PyRles on every rank
                run_counts.zero_()
                run_counts[gk5] = torch.from_numpy(np.diff(off)).to(dev)
                dist.all_reduce(run_counts)
            return off, runs, vals

        off, runs, vals = rle_step()
        r_loc = runs.numel()
        ck = reduce_sum_i64(_mix(runs * _M1 + vals.to(torch.int64) * _M2).sum())
        ms, _ = timed(rle_step, max(2, args.ops_steps // 2), 1)
        ms /= max(2, args.ops_steps // 2)
        ops["to_rle"] = dict(roof(16 * iv5.n + 16 * r_loc, ms), ms=ms, reads=n5, reads_per_s=n5 / (ms / 1e3),
                             runs=reduce_sum_i64(r_loc), checksum=ck)
        del off, runs, vals, r5
        torch.cuda.empty_cache()

    # ---- the public API (PyRanges.join on DataFrames), builder-sized: 2 M x 40 k rows ---------------------------
    if "api" in args.ops and rank == 0:
        try:
            ops["api_join"] = api_join_bench(dev)
        except Exception as ex:  # the API leg must not take the headline down
            ops["api_join"] = {"error": repr(ex)}

    if rank == 0:
        line = {
            "metric": "join overlap pairs/s", "value": value, "unit": "pairs/s", "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": join_ms / args.steps, "higher_is_better": True, "scaling": "strong",
            "vs_baseline": None, "dtype": "int64", "data": "synthetic", "config": workload_config(args, world),
            "pairs_per_step": pairs_all, "pairs_checksum": ck_pairs, "join_path": join_path,
            "count_overlaps_queries_per_s": ops.get("count_overlaps", {}).get("queries_per_s"),
            "roofline": roofline, "ops": ops, "cpu_baseline": cpu, "e2e": e2e, "gpu_launches": int(launches),
            "clocks": clocks,
        }
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


def api_join_bench(dev, n_reads=2_000_000, n_annot=40_000, reps=5):
    """gr.join(other) through the drop-in PyRanges class (DataFrames in, DataFrames out)."""
    import pandas as pd
    import torch

    from pyranges_b200 import pyranges_main as pr
    from pyranges_b200 import synthetic as S

    def frame(ka):
        key = np.repeat(np.arange(ka.n_keys), np.diff(ka.seg_off))
        names = np.asarray([c for c, _ in S.HG38])
        return pd.DataFrame({"Chromosome": names[ka.chrom_ids[key // 2]], "Start": ka.starts, "End": ka.ends,
                             "Strand": np.where(key % 2 == 0, "+", "-")})

    gr = pr.PyRanges(frame(S.reads_by_chrom(n_reads, seed=1)))
    other = pr.PyRanges(frame(S.annotations_by_chrom(n_annot, seed=2)))
    res = gr.join(other)
    rows = len(res)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(reps):
        gr.join(other)
    torch.cuda.synchronize()
    dt = (time.perf_counter() - t0) / reps
    return {"ms": 1e3 * dt, "rows_out": rows, "pairs_per_s": rows / dt,
            "workload": "PyRanges.join, %d x %d rows, strandedness=None, DataFrames in / out" % (n_reads, n_annot)}


def main():
    args = parse_args()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if args.impl == "reference":
        run_reference(args, rank, world)
        return
    run_b200(args, rank, world, local_rank)


if __name__ == "__main__":
    main()
