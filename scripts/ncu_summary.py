"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list: time per kernel name and per (name, grid)."""
import csv, collections, sys
rows = list(csv.reader(open(sys.argv[1])))
top = int(sys.argv[2]) if len(sys.argv) > 2 else 40
hdr = None
agg = collections.defaultdict(lambda: [0, 0.0])
agg2 = collections.defaultdict(lambda: [0, 0.0])
for r in rows:
    if hdr is None:
        if 'Kernel Name' in r:
            hdr = r
        continue
    d = dict(zip(hdr, r))
    if d.get('Metric Name') != 'gpu__time_duration.sum':
        continue
    v = float(d['Metric Value'].replace(',', ''))
    u = d['Metric Unit']
    v = v / 1e3 if u == 'ns' else v * 1e3 if u == 'ms' else v
    k = d['Kernel Name'].split('(')[0].replace('void ', '')[:48]
    agg[k][0] += 1; agg[k][1] += v
    k2 = (k, d['Grid Size'], d['Block Size'])
    agg2[k2][0] += 1; agg2[k2][1] += v
tot = sum(a[1] for a in agg.values())
n = sum(a[0] for a in agg.values())
print(f'total {tot / 1e3:.2f} ms over {n} launches (cold-cache, serialised: compare shares)')
for k, a in sorted(agg.items(), key=lambda kv: -kv[1][1])[:top]:
    print(f'{a[1] / 1e3:8.2f} ms {100 * a[1] / tot:5.1f}%  n={a[0]:4d} avg {a[1] / a[0]:7.1f} us  {k}')
print()
for k, a in sorted(agg2.items(), key=lambda kv: -kv[1][1])[:top]:
    print(f'{a[1] / 1e3:8.2f} ms n={a[0]:4d} avg {a[1] / a[0]:7.1f} us  {k[0]} grid={k[1]} block={k[2]}')
