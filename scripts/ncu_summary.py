"""Summarise an ncu --csv launch list (gpu__time_duration.sum) by kernel name: count, total us, share."""
import csv, sys, re, collections
rows = []
with open(sys.argv[1]) as f:
    lines = [l for l in f if not l.startswith("==")]
r = csv.DictReader(lines)
tot = collections.OrderedDict()
for row in r:
    if row.get("Metric Name") != "gpu__time_duration.sum":
        continue
    name = row["Kernel Name"]
    name = re.sub(r"^void ", "", name)
    name = re.sub(r"\(.*\)$", "", name) if len(name) > 150 else name
    v = float(row["Metric Value"].replace(",", ""))
    unit = row["Metric Unit"]
    us = v / 1000.0 if unit in ("ns", "nsecond") else (v if unit in ("us", "usecond") else v * 1000.0)
    c, t = tot.get(name, (0, 0.0))
    tot[name] = (c + 1, t + us)
total = sum(t for _, t in tot.values())
print("total %.1f us over %d launches" % (total, sum(c for c, _ in tot.values())))
for name, (c, t) in sorted(tot.items(), key=lambda kv: -kv[1][1]):
    print("%6.2f%% %10.1f us %5d x  %s" % (100 * t / total, t, c, name[:160]))
