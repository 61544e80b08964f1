#!/usr/bin/env python
"""Per-kernel totals of an ncu launch list (`--metrics gpu__time_duration.sum --csv`): count, total us, share.
usage: python tools/launch_shares.py gpurun_out/x.csv"""
import collections
import csv
import re
import sys

rows = list(csv.reader(l for l in open(sys.argv[1]) if l.startswith('"')))
h = rows[0]
ik, im, iv = h.index("Kernel Name"), h.index("Metric Name"), h.index("Metric Value")
tot = collections.OrderedDict()
for r in rows[1:]:
    if r[im] != "gpu__time_duration.sum":
        continue
    name = re.sub(r"\(.*", "", r[ik])[:110]
    c = tot.setdefault(name, [0, 0.0])
    c[0] += 1
    c[1] += float(r[iv].replace(",", "")) / 1e3
s = sum(v[1] for v in tot.values())
print("# %d launches, %.1f us of kernel time (cold-cache, serialised by ncu)" % (sum(v[0] for v in tot.values()), s))
for k, v in sorted(tot.items(), key=lambda kv: -kv[1][1]):
    print("%6.1f us %5.1f %%  x%-3d %s" % (v[1], 100 * v[1] / s, v[0], k))
