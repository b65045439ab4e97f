"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list: per-kernel totals, optional per-launch list."""
import csv, collections, sys
path = sys.argv[1]
keys = sys.argv[2:]
with open(path) as f:
    lines = [l for l in f if not l.startswith('==')]
rows = list(csv.DictReader(lines))
def us(row):
    v = float(row['Metric Value'].replace(',', '')); u = row['Metric Unit']
    return v / 1000 if u in ('ns', 'nsecond') else (v * 1000 if u in ('ms', 'msecond') else v)
agg = collections.defaultdict(lambda: [0, 0.0]); tot = 0
for r in rows:
    n = r['Kernel Name'][:70]; agg[n][0] += 1; agg[n][1] += us(r); tot += us(r)
print(f"total {tot:.0f} us over {len(rows)} launches")
for k, (n, t) in sorted(agg.items(), key=lambda kv: -kv[1][1])[:28]:
    print(f"{t:10.1f} us {n:5d}  {k}")
for key in keys:
    xs = [us(r) for r in rows if key in r['Kernel Name']]
    print(key, len(xs), f"total {sum(xs):.0f}", [f"{a:.0f}" for a in xs])
