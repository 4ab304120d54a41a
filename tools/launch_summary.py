"""Aggregate an `ncu --metrics gpu__time_duration.sum --csv` launch list per kernel name (development aid)."""
import collections, csv, re, sys

rows = list(csv.reader(open(sys.argv[1])))
steps = float(sys.argv[2]) if len(sys.argv) > 2 else 1.0
hdr, agg = None, collections.OrderedDict()
for r in rows:
    if hdr is None:
        if "Kernel Name" in r:
            hdr = r
        continue
    if len(r) < len(hdr):
        continue
    d = dict(zip(hdr, r))
    if d.get("Metric Name") != "gpu__time_duration.sum":
        continue
    name = re.sub(r"\(.*", "", d["Kernel Name"])
    v = float(d["Metric Value"].replace(",", ""))
    v = v / 1e3 if d["Metric Unit"] == "ns" else v * 1e3 if d["Metric Unit"] == "ms" else v
    a = agg.setdefault(name, [0, 0.0])
    a[0] += 1
    a[1] += v
tot = sum(a[1] for a in agg.values())
print(f"total {tot / steps:.1f} us/step, {sum(a[0] for a in agg.values()) / steps:.0f} launches/step")
for k, a in sorted(agg.items(), key=lambda x: -x[1][1]):
    print(f"{a[1] / steps:10.1f} us {a[0] / steps:6.1f}  {a[1] / a[0]:8.1f} avg  {100 * a[1] / tot:5.1f}%  {k[:100]}")
