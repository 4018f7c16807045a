"""Aggregates an `ncu --metrics gpu__time_duration.sum --csv` launch list per kernel name."""
import collections, csv, re, sys
lines = [l for l in open(sys.argv[1]) if not l.startswith("==")]
agg, tot = collections.OrderedDict(), 0.0
for row in csv.DictReader(lines):
    v = float(row["Metric Value"].replace(",", ""))
    v = {"us": v, "ns": v / 1000, "ms": v * 1000}[row["Metric Unit"]]
    short = re.sub(r"\(.*", "", row["Kernel Name"])[:80]
    a = agg.setdefault(short, [0, 0.0]); a[0] += 1; a[1] += v; tot += v
print("total %.1f us over %d launches" % (tot, sum(a[0] for a in agg.values())))
for k, (n, t) in sorted(agg.items(), key=lambda kv: -kv[1][1])[: int(sys.argv[2]) if len(sys.argv) > 2 else 40]:
    print("%9.1f us %5.1f%% x%3d  %s" % (t, 100 * t / tot, n, k))
