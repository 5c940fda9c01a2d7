"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list: per-kernel count, total and share."""
import csv, sys, collections, re
rows = []
with open(sys.argv[1]) as f:
    lines = [l for l in f if l.startswith('"')]
rd = csv.DictReader(lines)
tot = collections.defaultdict(lambda: [0, 0.0])
for r in rd:
    if r.get("Metric Name") != "gpu__time_duration.sum":
        continue
    name = re.sub(r"\(.*", "", r["Kernel Name"])
    v = float(r["Metric Value"].replace(",", ""))
    unit = r.get("Metric Unit", "ns")
    if unit in ("us", "usecond"): v *= 1e3
    elif unit in ("ms", "msecond"): v *= 1e6
    tot[name][0] += 1
    tot[name][1] += v
total = sum(v[1] for v in tot.values())
print(f"total {total/1e6:.3f} ms over {sum(v[0] for v in tot.values())} launches")
for k, (n, t) in sorted(tot.items(), key=lambda kv: -kv[1][1])[:40]:
    print(f"{t/1e6:10.3f} ms {100*t/total:6.2f}% {n:6d} x {t/n/1e3:9.2f} us  {k}")
