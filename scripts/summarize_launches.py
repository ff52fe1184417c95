"""Summarise an ncu `--metrics gpu__time_duration.sum --csv` launch list by kernel (shares of the step)."""
import collections, csv, sys
path = sys.argv[1]
lines = [l for l in open(path) if l.startswith('"')]
agg = collections.defaultdict(lambda: [0, 0.0])
for row in csv.DictReader(lines):
    name = row["Kernel Name"].split("(")[0].replace("void ", "").replace("madipm::", "").replace("(anonymous namespace)::", "")
    v = float(row["Metric Value"].replace(",", ""))
    us = v / 1000 if row["Metric Unit"] in ("ns", "nsecond") else v
    agg[name][0] += 1
    agg[name][1] += us
tot = sum(v[1] for v in agg.values())
print(f"# {path}: {sum(v[0] for v in agg.values())} launches, {tot:.1f} us of kernel time (ncu: cold-cache, serialised; compare SHARES)")
print(f"{'kernel':44s} {'launches':>8s} {'total_us':>10s} {'avg_us':>9s} {'share':>6s}")
for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    print(f"{k[:44]:44s} {v[0]:8d} {v[1]:10.1f} {v[1]/v[0]:9.1f} {100*v[1]/tot:5.1f}%")
