"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list per kernel (shares of the serialised total).
Usage: python tools/launch_summary.py launches.csv [first_fraction_to_skip]"""
import collections, csv, sys
lines = [l for l in open(sys.argv[1]) if l.startswith('"')]
rows = list(csv.DictReader(lines))
skip = float(sys.argv[2]) if len(sys.argv) > 2 else 0.0
rows = rows[int(len(rows) * skip):]
tot = sum(float(r["Metric Value"].replace(",", "")) for r in rows)
agg = collections.OrderedDict()
for r in rows:
    k = r["Kernel Name"].split("(")[0]
    a = agg.setdefault(k, [0, 0.0])
    a[0] += 1
    a[1] += float(r["Metric Value"].replace(",", ""))
print(f"total {tot / 1e6:.3f} ms over {len(rows)} launches")
for k, (n, t) in sorted(agg.items(), key=lambda x: -x[1][1]):
    print(f"{t / 1e6:9.3f} ms {100 * t / tot:5.1f}% n={n:3d} {k}")
