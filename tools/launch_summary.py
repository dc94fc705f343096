"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list: launches, total time and SHARE per kernel."""
import collections, csv, sys
rows = list(csv.reader(open(sys.argv[1])))
for i, r in enumerate(rows):
    if 'Kernel Name' in r:
        hdr, start = r, i + 1
        break
ix = {h: i for i, h in enumerate(hdr)}
agg = collections.defaultdict(lambda: [0, 0.0])
for r in rows[start:]:
    if len(r) < len(hdr) or r[ix['Metric Name']] != 'gpu__time_duration.sum':
        continue
    name = r[ix['Kernel Name']].split('(')[0].replace('void ', '').replace('digat::', '')
    v = float(r[ix['Metric Value']].replace(',', ''))
    unit = r[ix['Metric Unit']]
    v = v / 1e3 if unit.startswith('u') else v / 1e6 if unit.startswith('n') else v
    agg[name][0] += 1
    agg[name][1] += v
tot = sum(v[1] for v in agg.values())
print('kernel,launches,total_ms,share')
for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    print('%s,%d,%.3f,%.3f' % (k, v[0], v[1], v[1] / tot))
