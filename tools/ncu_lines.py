"""Per-source-line totals of one profiled kernel: ncu's per-SASS counters (report) joined, by instruction order, with
nvdisasm -g line info of the same function in a cubin.
   python tools/ncu_lines.py <report.ncu-rep> <kernel regex> <launch idx> <cubin> <mangled name substring> [top n]"""
import collections
import csv
import re
import subprocess
import sys

rep, kre, skip, cubin, mangled = sys.argv[1:6]
top = int(sys.argv[6]) if len(sys.argv) > 6 else 30
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--kernel-name", "regex:" + kre, "--launch-skip", skip,
                      "--launch-count", "1"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
h = rows[1]
col = {n: i for i, n in enumerate(h)}
data, seen = [], set()
for r in rows[2:]:
    if len(r) < len(h) or r[0] in seen or r[0] == "Address":
        continue
    seen.add(r[0])
    data.append(r)
dis = subprocess.run(["nvdisasm", "-g", cubin], capture_output=True, text=True).stdout.splitlines()
lines, cur, on = [], None, False
for ln in dis:
    if ln.startswith("\t.section\t.text.") or ln.startswith(".section\t.text.") or ".section" in ln and ".text." in ln:
        on = mangled in ln
        continue
    if not on:
        continue
    m = re.search(r'//## File "([^"]+)", line (\d+)', ln)
    if m:
        cur = (m.group(1).split("/")[-1], int(m.group(2)))
        continue
    if re.match(r"\s*/\*[0-9a-f]{4,}\*/\s+\S", ln):
        lines.append(cur)
print("ncu instructions %d, nvdisasm instructions %d" % (len(data), len(lines)))
n = min(len(data), len(lines))
agg = collections.defaultdict(lambda: [0, 0])
for i in range(n):
    a = agg[lines[i]]
    a[0] += int(data[i][col["Instructions Executed"]] or 0)
    a[1] += int(data[i][col["# Samples"]] or 0)
ti = sum(a[0] for a in agg.values())
ts = sum(a[1] for a in agg.values())
src = {}
for (f, l), (i, s) in sorted(agg.items(), key=lambda x: -x[1][0])[:top]:
    if f not in src:
        try:
            src[f] = open("pyranges_b200/csrc/" + f).read().splitlines()
        except OSError:
            src[f] = []
    text = src[f][l - 1].strip()[:90] if 0 < l <= len(src[f]) else ""
    print("%5.1f%% inst %5.1f%% samples  %s:%d  %s" % (100.0 * i / ti, 100.0 * s / max(ts, 1), f, l, text))
