"""Summarise an ncu launch list (gpu__time_duration.sum CSV) of scratch/train_step.py: the last of `steps` steps."""
import collections
import csv
import re
import sys

path, steps = sys.argv[1], int(sys.argv[2]) if len(sys.argv) > 2 else 3
thresh = float(sys.argv[3]) if len(sys.argv) > 3 else 1e9
with open(path) as f:
    lines = [l for l in f if l.startswith('"')]
rows = [(int(x['ID']), x['Kernel Name'], float(x['Metric Value']) / 1000, x['Grid Size']) for x in csv.DictReader(lines)]
per = len(rows) // steps
last = rows[-per:]
clean = lambda k: re.sub(r'\(.*', '', k).replace('void ', '').replace('(anonymous namespace)::', '').replace('<unnamed>::', '')
agg = collections.defaultdict(lambda: [0, 0.0])
for i, k, t, g in last:
    agg[clean(k)][0] += 1
    agg[clean(k)][1] += t
tot = sum(v[1] for v in agg.values())
print(f"{len(last)} launches, {tot:.1f} us")
for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1])[:28]:
    print(f"{v[1]:9.1f} {v[0]:4d} {100 * v[1] / tot:5.1f}%  {k[:100]}")
for i, k, t, g in last:
    if t > thresh:
        print(f"{i:5d} {t:8.1f} {g:18s} {clean(k)[:70]}")
