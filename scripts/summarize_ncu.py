import csv, collections, re, sys
path = sys.argv[1]
with open(path) as f:
    lines = [l for l in f if not l.startswith('==')]
agg = collections.defaultdict(lambda: [0.0, 0]); tot = 0
for row in csv.DictReader(lines):
    if row.get('Metric Name') != 'gpu__time_duration.sum': continue
    v = float(row['Metric Value'].replace(',', '')); unit = row['Metric Unit']
    v = v / 1e3 if unit == 'ns' else (v * 1e3 if unit == 'ms' else v)
    name = re.sub(r'\(.*', '', row['Kernel Name'])[:80]
    agg[name][0] += v; agg[name][1] += 1; tot += v
print(f"total GPU time {tot/1e3:.2f} ms over {sum(c for _, c in agg.values())} launches (ncu: cold cache, serialised)")
for n, (us, c) in sorted(agg.items(), key=lambda kv: -kv[1][0])[:int(sys.argv[2]) if len(sys.argv) > 2 else 40]:
    print(f"{n:82s} {us/1e3:8.2f} ms {c:5d} calls {us/c:8.1f} us  {100*us/tot:5.1f}%")
