"""Per-kernel breakdown of ONE training step from an `ncu --metrics gpu__time_duration.sum --csv` launch list of
`DGCNN_CUDA_GRAPH=0 python bench.py --steps 1 --warmup 1 --no-cpu-baseline` (steps are delimited by the Adam kernel)."""
import csv, re, sys
lines = [l for l in open(sys.argv[1]) if not l.startswith('==')]
rows = list(csv.DictReader(lines))
idx = [i for i, r in enumerate(rows) if 'adam_tf' in r['Kernel Name']]
which = int(sys.argv[3]) if len(sys.argv) > 3 else 1
a, b = idx[which] + 1, idx[which + 1] + 1
agg, tot = {}, 0.0
for r in rows[a:b]:
    n = re.sub(r'\(.*', '', r['Kernel Name'])[:70]
    v = float(r['Metric Value'].replace(',', '')); u = r['Metric Unit']
    v = {'us': v, 'ns': v / 1000, 'ms': v * 1000}[u]
    e = agg.setdefault(n, [0, 0.0]); e[0] += 1; e[1] += v; tot += v
print('one step: %.1f us over %d launches' % (tot, b - a))
for k, (n, t) in sorted(agg.items(), key=lambda kv: -kv[1][1])[: int(sys.argv[2]) if len(sys.argv) > 2 else 40]:
    print('%8.1f us %5.1f%% x%3d  %s' % (t, 100 * t / tot, n, k))
