#!/usr/bin/env python
"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list into per-kernel totals and
per-step shares (cold-cache, serialised launches: compare SHARES with bench.py's live numbers)."""
import collections
import csv
import sys


def main(path, out):
    with open(path) as f:
        lines = [l for l in f if not l.startswith("==")]
    agg = collections.OrderedDict()
    for row in csv.DictReader(lines):
        k = row["Kernel Name"]
        v = float(row["Metric Value"].replace(",", ""))
        u = row["Metric Unit"]
        v = v / 1e3 if u in ("ns", "nsecond") else (v * 1e3 if u in ("ms", "msecond") else v)
        a = agg.setdefault(k, [0, 0.0])
        a[0] += 1
        a[1] += v
    msm = {k: a for k, a in agg.items() if "msm" in k or "scan" in k}
    ntt = {k: a for k, a in agg.items() if "ntt_pass" in k}
    with open(out, "w") as o:
        o.write("# ncu launch list summary (%s)\n\n" % path)
        o.write("| kernel | launches | total us | avg us | share of all | share of MSM step |\n|---|---:|---:|---:|---:|---:|\n")
        tot = sum(a[1] for a in agg.values())
        msm_tot = sum(a[1] for k, a in msm.items() if "precompute" not in k)
        for k, a in sorted(agg.items(), key=lambda kv: -kv[1][1]):
            share_msm = "%.1f%%" % (100 * a[1] / msm_tot) if (k in msm and "precompute" not in k) else ""
            o.write("| `%s` | %d | %.1f | %.1f | %.1f%% | %s |\n" % (k.split("(")[0][:70], a[0], a[1], a[1] / a[0], 100 * a[1] / tot, share_msm))
        if ntt:
            o.write("\nNTT passes: %d launches, avg %.1f us\n" % (sum(a[0] for a in ntt.values()), sum(a[1] for a in ntt.values()) / sum(a[0] for a in ntt.values())))


if __name__ == "__main__":
    main(sys.argv[1], sys.argv[2])
