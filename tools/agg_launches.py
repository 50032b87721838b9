"""Aggregate an `ncu --metrics gpu__time_duration.sum --csv` launch list by kernel name."""
import csv, sys, re, collections
rows = []
with open(sys.argv[1]) as f:
    lines = [l for l in f if not l.startswith("==")]
rd = csv.DictReader(lines)
tot = collections.defaultdict(lambda: [0, 0.0])
for r in rd:
    if r.get("Metric Name") != "gpu__time_duration.sum":
        continue
    name = r["Kernel Name"]
    name = re.sub(r"\(.*", "", name)
    v = float(r["Metric Value"].replace(",", ""))
    unit = r.get("Metric Unit", "ns")
    if unit in ("us", "usecond"): v *= 1e3
    if unit in ("ms", "msecond"): v *= 1e6
    tot[name][0] += 1; tot[name][1] += v
total = sum(v[1] for v in tot.values())
print("total %.3f ms over %d launches" % (total / 1e6, sum(v[0] for v in tot.values())))
for k, (n, t) in sorted(tot.items(), key=lambda kv: -kv[1][1])[: int(sys.argv[2]) if len(sys.argv) > 2 else 40]:
    print("%8.3f ms %5.1f%% %6d x %8.1f us  %s" % (t / 1e6, 100 * t / total, n, t / n / 1e3, k[:110]))
