"""Compact view of a bench.py JSON line: python tools/brief.py <file>"""
import json
import sys

for line in open(sys.argv[1]):
    line = line.strip()
    if not line.startswith("{"):
        continue
    d = json.loads(line)
    print("value %.3e %s  ms/step %.4f  n_gpus %s  path %s  pairs %s  checksum %s" % (
        d["value"], d["unit"], d["ms_per_step"], d["n_gpus"], d.get("join_path"), d.get("pairs_per_step"), d.get("pairs_checksum")))
    r = d.get("roofline") or {}
    print("roofline: %s frac %.3f achieved %.0f GB/s | phases %s" % (r.get("kernel"), r.get("frac", 0), r.get("achieved", 0),
                                                                     {k: round(v, 4) for k, v in (r.get("rank0_phase_ms") or {}).items()}))
    p = r.get("pipeline") or {}
    print("pipeline: frac_8d %.3f frac_no_labels %.3f" % (p.get("frac_8d", 0), p.get("frac_no_labels", 0)))
    for k, v in (d.get("ops") or {}).items():
        if "error" in v:
            print("  %-24s ERROR %s" % (k, v["error"]))
            continue
        print("  %-24s %9.4f ms  frac %.3f  %s" % (k, v.get("ms", 0), v.get("frac", 0),
                                                   {a: b for a, b in v.items() if a in ("checksum", "clusters", "merged", "runs", "rows_out", "pairs_per_s", "queries_per_s", "rows_per_s", "reads_per_s")}))
    if d.get("e2e"):
        print("e2e: %.3e %s  %.2f ms/step" % (d["e2e"]["value"], d["e2e"]["unit"], d["e2e"].get("ms_per_step", 0)))
    if d.get("cpu_baseline"):
        print("cpu: %.3e %s (%s cores, %s)" % (d["cpu_baseline"]["value"], d["cpu_baseline"]["unit"], d["cpu_baseline"]["cores"], d["cpu_baseline"]["kind"]))
    print("clocks:", d.get("clocks"), "launches:", d.get("gpu_launches"))
