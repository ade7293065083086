"""Per-kernel shares from an `ncu --metrics gpu__time_duration.sum --csv` launch list.

    python tools/launch_summary.py launches.csv "header line" [top=40]
"""
import collections
import csv
import sys

path, header = sys.argv[1], sys.argv[2]
top = int(sys.argv[3]) if len(sys.argv) > 3 else 40
rows = list(csv.reader(l for l in open(path) if l.startswith('"')))
hdr = rows[0]
ik, iv, iu = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
agg = collections.defaultdict(lambda: [0.0, 0])
for r in rows[1:]:
    v, u = float(r[iv].replace(",", "")), r[iu]
    ms = v / 1e6 if u in ("nsecond", "ns") else v / 1e3 if u in ("usecond", "us") else v
    agg[r[ik][:150]][0] += ms
    agg[r[ik][:150]][1] += 1
tot = sum(v[0] for v in agg.values())
print(header)
print("(per-launch times are cold-cache and serialised: SHARES matter)")
print("total %.1f ms over %d launches" % (tot, sum(v[1] for v in agg.values())))
for k, v in sorted(agg.items(), key=lambda kv: -kv[1][0])[:top]:
    print("%9.2f ms %5.1f%% n=%5d %s" % (v[0], 100 * v[0] / tot, v[1], k))
