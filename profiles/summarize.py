"""Summarise ncu outputs into small text files that can be committed:
   python profiles/summarize.py launches <launches.csv>           per-kernel totals / shares
   python profiles/summarize.py full <report.ncu-rep>             key metrics of every profiled launch"""
import collections
import csv
import subprocess
import sys

KEYS = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_sectors.sum",
        "lts__throughput.avg.pct_of_peak_sustained_elapsed", "l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum",
        "l1tex__throughput.avg.pct_of_peak_sustained_elapsed", "l1tex__t_sector_hit_rate.pct",
        "lts__t_sector_hit_rate.pct", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__thread_inst_executed_per_inst_executed.ratio",
        "launch__registers_per_thread", "launch__grid_size", "launch__block_size"]


def launches(path):
    hdr, agg = None, collections.OrderedDict()
    for r in csv.reader(open(path)):
        if len(r) < 6:
            continue
        if r[0] == "ID":
            hdr = r
            continue
        if hdr is None:
            continue
        d = dict(zip(hdr, r))
        a = agg.setdefault(d["Kernel Name"][:70], [0, 0.0])
        a[0] += 1
        a[1] += float(d["Metric Value"].replace(",", ""))
    tot = sum(a[1] for a in agg.values())
    print("# per-kernel device time (ncu gpu__time_duration.sum, cold-cache, serialised: compare SHARES)")
    for k, (n, t) in sorted(agg.items(), key=lambda x: -x[1][1]):
        print("%-72s n=%4d total=%10.1f us avg=%8.1f us share=%5.1f%%" % (k, n, t / 1e3, t / 1e3 / n, 100 * t / tot))


def full(path):
    out = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr, units = rows[0], rows[1]
    idx = {h: i for i, h in enumerate(hdr)}
    for r in rows[2:]:
        print("--- %s" % r[idx["Kernel Name"]][:100])
        for k in KEYS:
            if k in idx:
                print("  %-66s %s %s" % (k, r[idx[k]], units[idx[k]]))


if __name__ == "__main__":
    {"launches": launches, "full": full}[sys.argv[1]](sys.argv[2])
