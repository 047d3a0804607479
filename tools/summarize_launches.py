"""Summarise an ncu `--metrics gpu__time_duration.sum --csv` launch list per kernel name (run here, no GPU)."""
import csv, sys, collections, re
rows = []
with open(sys.argv[1]) as f:
    lines = [l for l in f if not l.startswith("==")]
r = csv.DictReader(lines)
tot = collections.defaultdict(lambda: [0, 0.0])
for row in r:
    if row.get("Metric Name") != "gpu__time_duration.sum":
        continue
    v = float(row["Metric Value"].replace(",", ""))
    unit = row.get("Metric Unit", "ns")
    ns = v * {"ns": 1, "us": 1e3, "usecond": 1e3, "ms": 1e6, "msecond": 1e6, "nsecond": 1, "second": 1e9}.get(unit, 1)
    name = re.sub(r"\(.*", "", row["Kernel Name"])
    name = re.sub(r"^void |^.*::", "", name) if "at::" not in row["Kernel Name"] else "torch:" + name[:60]
    tot[name][0] += 1
    tot[name][1] += ns
total = sum(v[1] for v in tot.values())
print("total %.3f ms over %d launches" % (total / 1e6, sum(v[0] for v in tot.values())))
print("%-60s %6s %10s %6s" % ("kernel", "count", "ms", "share"))
for k, (c, ns) in sorted(tot.items(), key=lambda kv: -kv[1][1]):
    print("%-60s %6d %10.3f %5.1f%%" % (k[:60], c, ns / 1e6, 100 * ns / total))
