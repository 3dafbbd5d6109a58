#!/usr/bin/env python
"""bench.py -- join overlap pairs/s and count_overlaps queries/s on synthetic hg38-shaped intervals.

Workload (BASELINE.json configs[1]): 10 M reads (100 bp) x 200 k gene/exon annotations, strandedness=None
(reads are keyed by (chromosome, strand), every key is joined against both strands of its chromosome).
One "step" = one full join pass: subject index build (the reference rebuilds its NCLS on every call,
pyranges/methods/join.py:12) -> count -> scan -> emit of all (query row, subject row) int64 pairs.

  python bench.py [--gpus N --steps K --warmup W]             this repo's CUDA path
  python bench.py --impl reference [...]                       CPU restatement of ncls on the host cores

N > 1 (torchrun): the read set grows with N (N x 10 M reads against the same 200 k annotations: reads and pairs
per GPU stay fixed = weak scaling); chromosomes are assigned to ranks by a size-balanced (LPT) rule, every rank joins its own
chromosomes with no data-path collective, and the ranks exchange only their per-key pair counts (one small NCCL
all-reduce per step) from which every rank knows the global output offsets.
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


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--reads", type=int, default=10_000_000)
    ap.add_argument("--annot", type=int, default=200_000)
    ap.add_argument("--cpu-seconds", type=float, default=15.0, help="budget of the cpu_baseline sample")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    return ap.parse_args()


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


def ncu_traffic(kernel):
    """dram bytes per launch of `kernel` from the committed ncu --set full summary, if any."""
    try:
        with open(os.path.join(ROOT, "profiles", "traffic.json")) as f:
            return json.load(f).get(kernel)
    except Exception:
        return None


def make_workload(args, rank, world=1):
    from pyranges_b200 import synthetic as S

    chroms = None if world == 1 else S.shard_chromosomes(world, rank)
    reads = S.reads(args.reads * world, seed=1, chroms=chroms)
    annot = S.annotations(args.annot, seed=2, chroms=chroms)  # the genome annotation does not grow with N
    seg_group = (np.arange(reads.n_keys) // 2).astype(np.int32)  # strandedness=None: partner = whole chromosome
    return reads, annot, seg_group


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
    p, q, dt, _ = oracle_join(reads, annot, seg_group, frac=0.05, threads=threads)
    frac = float(min(1.0, 0.05 * 6.0 / max(dt, 1e-3)))  # ~6 s per step at most
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
    sample = "first %.1f%% of the rows of every read key (%d queries/step) x all %d annotations, %d threads over chromosomes" % (
        100 * frac, queries // max(args.steps, 1), annot.n, threads)
    line = {
        "impl": "reference", "metric": "join overlap pairs/s", "value": val, "unit": "pairs/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * dt / max(args.steps, 1),
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "int64", "data": "synthetic",
        "config": workload_config(args, args.gpus),
        "count_overlaps_queries_per_s": queries / dt,
        "cpu_baseline": {"value": val, "unit": "pairs/s", "cores": threads, "kind": "port", "sample": sample},
        "e2e": {"value": val, "unit": "pairs/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line), flush=True)


def workload_config(args, n_gpus):
    return {"workload": "synthetic hg38-shaped %d reads (100 bp) x %d gene/exon annotations: join (+ count_overlaps), "
                        "strandedness=None, per GPU" % (args.reads, args.annot),
            "reads_per_gpu": args.reads, "annotations_genome_wide": args.annot, "keys": 48,
            "l2": "inputs (%.0f MB of int64 query columns) exceed the 126 MB L2; no flush" % (16 * args.reads / 1e6),
            "index_build": "inside the timed region",
            "parallelism": "chromosomes sharded over %d GPU(s) by size (LPT); %d x %d reads genome-wide against the "
                           "same %d annotations" % (n_gpus, n_gpus, args.reads, args.annot)}


# ---------------------------------------------------------------------------------------------------
def run_b200(args, rank, world, local_rank):
    import ctypes as C

    import torch
    import torch.distributed as dist

    from pyranges_b200 import _cabi as cabi
    from pyranges_b200 import engine as E
    from pyranges_b200 import hostapi

    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    L = cabi.lib()

    reads, annot, seg_group = make_workload(args, rank, world)
    groups = annot.chrom_ranges()
    q_iv = E.DeviceIntervals.from_numpy(reads.starts, reads.ends, reads.seg_off, dev)
    s_iv = E.DeviceIntervals.from_numpy(annot.starts, annot.ends, annot.seg_off, dev)
    s_iv.bounds  # ingest-time metadata (per-key min Start / max End), like the chromosome table
    nq, ns = reads.n, annot.n
    ev = lambda: torch.cuda.Event(enable_timing=True)  # noqa: E731

    state = {}
    side = torch.cuda.Stream(dev) if world > 1 else None  # carries the per-key count exchange of the multi-GPU step

    def join_step(timers=None):
        e = [ev() for _ in range(5)] if timers is not None else None
        if e:
            e[0].record()
        ix = E.SubjectIndex(s_iv, groups)
        q = E.Queries(q_iv, seg_group)
        if e:
            e[1].record()
        # join needs the pairs, not the per-row counts: the count pass only leaves its hit lists in ws
        _, ws, total = ix.count(q, E.IV_MODE_ALL, want_counts=False)
        if world > 1:
            # per-key pair counts -> every rank can derive the global output offsets (the only collective; keys are
            # disjoint across ranks, so the sum over ranks is the gathered vector).  They are known right after the
            # count pass: the all-reduce runs on a side stream while the emit pass writes the pairs.
            per_key = ix.segment_pair_counts(q, ws, total)
            side.wait_stream(torch.cuda.current_stream())
            with torch.cuda.stream(side):
                dist.all_reduce(per_key)
            per_key.record_stream(side)
        if e:
            e[2].record()
        # speculative emission (engine.SubjectIndex.pairs(capacity=...)): buffers sized from the previous step's
        # pair count, emit launched right behind the count, total read afterwards; exact redo if it was too small
        cap = state.get("cap", 0)
        p = cap if cap else int(total.item())
        out_q = torch.empty(p, dtype=torch.int64, device=dev)
        out_s = torch.empty(p, dtype=torch.int64, device=dev)
        if e:
            e[3].record()
        cabi.check(L.iv_emit_pairs(C.byref(ix.c), C.byref(q.c), E.IV_MODE_ALL, None, E._ptr(ws), E._ptr(out_q),
                                   E._ptr(out_s), None, out_q.numel(), E._stream()), "iv_emit_pairs")
        if e:
            e[4].record()
            timers.append(e)
        if cap:
            p = int(total.item())
            if p > cap:
                out_q = torch.empty(p, dtype=torch.int64, device=dev)
                out_s = torch.empty(p, dtype=torch.int64, device=dev)
                cabi.check(L.iv_emit_pairs(C.byref(ix.c), C.byref(q.c), E.IV_MODE_ALL, None, E._ptr(ws),
                                           E._ptr(out_q), E._ptr(out_s), None, p, E._stream()), "iv_emit_pairs")
            else:
                out_q, out_s = out_q[:p], out_s[:p]
        state["cap"] = int(p * 1.02) + 4096
        if world > 1:
            torch.cuda.current_stream().wait_stream(side)
        return p, out_q, out_s

    def count_step():
        ix = E.SubjectIndex(s_iv, groups)
        q = E.Queries(q_iv, seg_group)
        c, _, _ = ix.count(q, E.IV_MODE_ALL, for_emit=False)
        return c

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(fn, steps, warmup, collect=None):
        for _ in range(warmup):
            fn()
        barrier()
        a, b = ev(), ev()
        n0 = L.iv_launch_count()
        a.record()
        for _ in range(steps):
            fn() if collect is None else fn(collect)
        b.record()
        barrier()
        ms = a.elapsed_time(b)
        launches = L.iv_launch_count() - n0
        if world > 1:
            t = torch.tensor([ms], dtype=torch.float64, device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms = float(t.item())
        return ms, launches

    # ---- device-resident join (the headline `value`) ---------------------------------------------
    p, _, _ = join_step()
    sampler = ClockSampler(local_rank) if rank == 0 else None
    time.sleep(0.3)
    t0 = time.perf_counter()
    timers = []
    join_ms, launches = timed(join_step, args.steps, args.warmup, collect=timers)
    count_ms, _ = timed(count_step, args.steps, args.warmup)
    t1 = time.perf_counter()
    clocks = sampler.summary(t0, t1) if sampler else None
    pairs_all = p
    if world > 1:
        t = torch.tensor([p], dtype=torch.int64, device=dev)
        dist.all_reduce(t)
        pairs_all = int(t.item())
    value = pairs_all * args.steps / (join_ms / 1e3)
    count_qps = nq * world * args.steps / (count_ms / 1e3)

    # per-phase device times (same timed region): index build | count+scan | emit
    torch.cuda.synchronize()
    ph = np.asarray([[e[i].elapsed_time(e[i + 1]) for i in range(4)] for e in timers])
    ph_ms = ph.mean(axis=0)
    peak, peak_src = measured_peak_gbs()
    # Phases of the step (events sit directly around the C calls): iv_count_overlaps = k_count_lean (+ the ragged
    # tail, k_tile_sums and the single-block scan of the 10^4 tile sums: a few us), iv_emit_pairs = the one k_emit
    # launch.  Algorithmic bytes: the count pass reads Start/End of every query and writes one 16-byte hit-list entry
    # per query with hits (+ 12 bytes of granule metadata per 64 rows); the emit pass reads those entries and writes
    # the pairs.  The roofline is quoted on whichever of the two takes longer.
    hits = int((count_step() > 0).sum().item())
    alg = {"iv_index_build": 24 * ns, "k_count_lean": 16 * nq + 16 * hits + 12 * (nq // 64), "alloc": 0,
           "k_emit": 16 * hits + 12 * (nq // 64) + 16 * p}
    names = ["iv_index_build", "k_count_lean", "alloc", "k_emit"]
    dom = 1 if ph_ms[1] >= ph_ms[3] else 3
    ach = alg[names[dom]] / (ph_ms[dom] / 1e3) / 1e9
    pipe = 24 * nq + 24 * ns + 16 * p  # the join's algorithmic bytes (both tables incl. labels in, pairs out)
    roofline = {"bound": "hbm", "kernel": names[dom], "achieved": ach, "peak": peak, "unit": "GB/s",
                "frac": ach / peak, "traffic": ncu_traffic(names[dom]), "peak_source": peak_src,
                "algorithmic_bytes_per_launch": alg[names[dom]], "hit_queries": hits,
                "phase_ms": {n: float(m) for n, m in zip(names, ph_ms)},
                "phase_achieved_gbs": {n: alg[n] / (m / 1e3) / 1e9 for n, m in zip(names, ph_ms) if alg[n]},
                "pipeline_algorithmic_bytes": pipe,
                "pipeline_achieved": pipe / (join_ms / args.steps / 1e3) / 1e9,
                "pipeline_frac": pipe / (join_ms / args.steps / 1e3) / 1e9 / peak}

    # ---- end to end through the host-facing call: pinned host arrays in, host index arrays out ----
    e2e = None
    if not args.no_e2e:
        hq_s, hq_e = torch.from_numpy(reads.starts).pin_memory(), torch.from_numpy(reads.ends).pin_memory()
        hs_s, hs_e = torch.from_numpy(annot.starts).pin_memory(), torch.from_numpy(annot.ends).pin_memory()
        cap = int(p * 1.1) + 1024
        out = (torch.empty(cap, dtype=torch.int64).pin_memory(), torch.empty(cap, dtype=torch.int64).pin_memory())
        bounds = s_iv.bounds

        def e2e_step():
            nb = hostapi.NCLSBatch(hs_s, hs_e, annot.seg_off, groups, device=dev, bounds=bounds)
            a, _ = nb.all_overlaps_both(hq_s, hq_e, reads.seg_off, seg_group, out=out)
            return a.numel()

        for _ in range(max(1, args.warmup)):
            e2e_step()
        barrier()
        te = time.perf_counter()
        pe = 0
        for _ in range(args.steps):
            pe += e2e_step()
        barrier()
        dt = time.perf_counter() - te
        if world > 1:
            t = torch.tensor([dt, float(pe)], dtype=torch.float64, device=dev)
            tm = t.clone()
            dist.all_reduce(tm, op=dist.ReduceOp.MAX)
            dist.all_reduce(t)
            dt, pe = float(tm[0].item()), float(t[1].item())
        e2e = {"value": pe / dt, "unit": "pairs/s", "h2d_bytes_per_step": 16 * nq + 16 * ns,
               "d2h_bytes_per_step": 16 * p, "ms_per_step": 1e3 * dt / args.steps,
               "call": "pyranges_b200.hostapi.NCLSBatch(...).all_overlaps_both(...) on pinned host int64 arrays"}

    # ---- CPU baseline beside it (rank 0, N = 1 only) ------------------------------------------------
    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        from oracle import oracle as orc

        orc.build()
        cp = cq = 0
        cdt = 0.0
        passes = 0
        while cdt < args.cpu_seconds and passes < 50:
            a, b, d, done = oracle_join(reads, annot, seg_group, frac=1.0, threads=1, budget_s=args.cpu_seconds)
            cp, cq, cdt, passes = cp + a, cq + b, cdt + d, passes + 1
        cpu = {"value": cp / cdt, "unit": "pairs/s", "cores": 1, "kind": "port",
               "queries_per_s": cq / cdt,
               "sample": "%d pass(es) over %d chromosomes (%d reads x their annotations each): NCLS build + "
                         "all_overlaps_both per chromosome, %.1f s on 1 core" % (passes, done, cq // max(passes, 1), cdt)}

    if rank == 0:
        line = {
            "metric": "join overlap pairs/s", "value": value, "unit": "pairs/s", "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": join_ms / args.steps, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "int64", "data": "synthetic", "config": workload_config(args, world),
            "pairs_per_step": pairs_all, "count_overlaps_queries_per_s": count_qps,
            "count_overlaps_ms_per_step": count_ms / args.steps,
            "roofline": roofline, "cpu_baseline": cpu, "e2e": e2e, "gpu_launches": int(launches), "clocks": clocks,
        }
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


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
