"""Summarise an ncu `--metrics gpu__time_duration.sum --csv` launch list per kernel name."""
import csv, collections, re, sys
for f in sys.argv[1:]:
    rows = [r for r in csv.reader(l for l in open(f) if l.startswith('"'))]
    h = rows[0]; ki = h.index('Kernel Name'); vi = h.index('Metric Value'); ui = h.index('Metric Unit')
    d = collections.defaultdict(lambda: [0, 0.0])
    for r in rows[1:]:
        v = float(r[vi].replace(',', ''))
        v = v / 1000 if r[ui] == 'ns' else v * 1000 if r[ui] == 'ms' else v
        n = re.sub(r'\(.*', '', r[ki])
        d[n][0] += 1; d[n][1] += v
    tot = sum(v[1] for v in d.values())
    print(f, 'total us', round(tot, 1))
    for k, v in sorted(d.items(), key=lambda kv: -kv[1][1]):
        print(f'  {v[1]:10.1f} us {v[0]:4d} launches {v[1]/v[0]:8.1f} us/launch {100*v[1]/tot:5.1f}%  {k}')
