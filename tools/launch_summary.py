#!/usr/bin/env python
"""Summarise an ncu `--metrics gpu__time_duration.sum --csv` launch list per kernel."""
import collections, csv, re, sys
path = sys.argv[1]
lines = [l for l in open(path) if not l.startswith("==")]
agg, tot = collections.OrderedDict(), 0.0
for row in csv.DictReader(lines):
    try:
        t = float(row["Metric Value"].replace(",", ""))
    except Exception:
        continue
    u = row["Metric Unit"]
    t = t / 1000 if u == "ns" else (t * 1000 if u == "ms" else t)
    k = re.sub(r"\(.*", "", row["Kernel Name"]).replace("void ", "").replace("<unnamed>::", "")[:58]
    a = agg.setdefault(k, [0, 0.0]); a[0] += 1; a[1] += t; tot += t
print("total %.1f us over %d launches" % (tot, sum(a[0] for a in agg.values())))
for k, (c, t) in sorted(agg.items(), key=lambda kv: -kv[1][1])[: int(sys.argv[2]) if len(sys.argv) > 2 else 22]:
    print("%-60s n=%4d total=%9.1f us avg=%7.2f us share=%5.1f%%" % (k, c, t, t / c, 100 * t / tot))
