"""Summarise an ncu launch list (gpu__time_duration.sum per launch) of `bench.py --ncu`: the LAST step's launches grouped by
kernel. usage: launch_summary.py <csv> <launches_per_step> [--list]"""
import collections
import csv
import re
import sys

rows = list(csv.reader(open(sys.argv[1])))
hdr = [i for i, r in enumerate(rows) if r and r[0] == "ID"][0]
H = rows[hdr]
ni, vi, gi = H.index("Kernel Name"), H.index("Metric Value"), H.index("Grid Size")
L = [(r[ni], float(r[vi].replace(",", "")) / 1e3, r[gi]) for r in rows[hdr + 1:] if len(r) > vi and r[vi].replace(",", "").replace(".", "").isdigit()]
n = int(sys.argv[2])
print("%d launches in the list; last %d = one step" % (len(L), n))
L = L[-n:]
if "--list" in sys.argv:
    for i, (k, v, g) in enumerate(L):
        print("%3d %8.1f us  %-14s %s" % (i, v, g, k[:150]))
agg = collections.OrderedDict()
for k, v, g in L:
    k = re.sub(r"\(.*", "", k)
    k = re.sub(r"^void ", "", k)[:70]
    agg.setdefault(k, [0.0, 0])
    agg[k][0] += v
    agg[k][1] += 1
tot = sum(v[0] for v in agg.values())
print("| kernel | launches | us / step | share |\n|---|---:|---:|---:|")
for k, v in sorted(agg.items(), key=lambda kv: -kv[1][0]):
    print("| `%s` | %d | %.1f | %.1f%% |" % (k, v[1], v[0], 100 * v[0] / tot))
print("| total | %d | %.1f | |" % (len(L), tot))
