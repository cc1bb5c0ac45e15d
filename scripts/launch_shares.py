"""Share of every kernel in the launches of an ncu `--metrics gpu__time_duration.sum --csv` log (cold-cache, serialised: compare shares).
usage: python scripts/launch_shares.py <launches.csv>"""
import collections
import csv
import sys

rows = list(csv.reader(open(sys.argv[1])))
hi = next(i for i, r in enumerate(rows) if "Kernel Name" in r)
hdr = rows[hi]; data = rows[hi + 1:]
ik = hdr.index("Kernel Name"); iv = hdr.index("Metric Value"); iu = hdr.index("Metric Unit")
agg = collections.OrderedDict()
for r in data:
    if len(r) <= iv:
        continue
    name = r[ik].split("(")[0].replace("void ", "").replace("thb::", "")
    name = name.split("<")[0] if name.startswith("cub::") else name
    v = float(r[iv].replace(",", "")) * {"ns": 1e-3, "us": 1.0, "ms": 1e3}.get(r[iu], 1.0)
    a = agg.setdefault(name[:48], [0, 0.0]); a[0] += 1; a[1] += v
tot = sum(v[1] for v in agg.values())
print("| kernel | launches | total us | share |\n|---|---|---|---|")
for k, (n, t) in sorted(agg.items(), key=lambda x: -x[1][1]):
    print("| `%s` | %d | %.1f | %.1f %% |" % (k, n, t, 100 * t / tot))
print("| all | %d | %.1f | |" % (sum(v[0] for v in agg.values()), tot))
