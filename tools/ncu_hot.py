"""Hot spots of one profiled kernel from an ncu report (needs -lineinfo / --import-source on):
   python tools/ncu_hot.py <report.ncu-rep> <kernel regex> [launch index] [top n]
Prints the stall-reason totals and the SASS instructions with the most stall samples (with the two instructions before
them: the sampled instruction usually WAITS for what precedes it)."""
import csv
import subprocess
import sys

rep, kre = sys.argv[1], sys.argv[2]
skip = sys.argv[3] if len(sys.argv) > 3 else "0"
top = int(sys.argv[4]) if len(sys.argv) > 4 else 25
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--kernel-name", "regex:" + kre, "--launch-skip", skip,
                      "--launch-count", "1"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
print(rows[0][1][:150])
h = rows[1]
col = {n: i for i, n in enumerate(h)}
data = []
seen = set()
for r in rows[2:]:
    if len(r) < len(h) or r[0] in seen or r[0] == "Address":  # the listing can come twice
        continue
    seen.add(r[0])
    data.append(r)
ns, ie = col["# Samples"], col["Instructions Executed"]
tot_s = sum(int(r[ns] or 0) for r in data)
tot_i = sum(int(r[ie] or 0) for r in data)
print("SASS instructions %d, warp instructions executed %d, stall samples %d" % (len(data), tot_i, tot_s))
reasons = [n for n in h if n.startswith("stall_") and "Not Issued" not in n]
tot = {n: sum(int(r[col[n]] or 0) for r in data) for n in reasons}
print("stall reasons:", ", ".join("%s %.1f%%" % (n[6:], 100.0 * v / max(tot_s, 1)) for n, v in sorted(tot.items(), key=lambda x: -x[1]) if v))
order = sorted(range(len(data)), key=lambda i: -int(data[i][ns] or 0))[:top]
for i in order:
    r = data[i]
    why = sorted(((int(r[col[n]] or 0), n[6:]) for n in reasons), reverse=True)[:2]
    print("%5.2f%% samples %5.2f%% inst  [%s]" % (100.0 * int(r[ns] or 0) / max(tot_s, 1), 100.0 * int(r[ie] or 0) / max(tot_i, 1),
                                               ", ".join("%s %d" % (n, v) for v, n in why if v)))
    for j in range(max(0, i - 2), i + 1):
        print("      %s %s" % ("->" if j == i else "  ", data[j][col["Source"]].strip()[:110]))
