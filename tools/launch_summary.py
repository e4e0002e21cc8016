"""Per-kernel summary of an ncu launch list (`--metrics gpu__time_duration.sum --csv`): python tools/launch_summary.py file.csv"""
import csv, collections, sys
rows = list(csv.reader(open(sys.argv[1])))
hi = [i for i, r in enumerate(rows) if r and r[0] == 'ID'][0]
h = rows[hi]; ik = h.index('Kernel Name'); iv = h.index('Metric Value')
agg = collections.defaultdict(list)
for r in rows[hi + 2:]:
    if len(r) > iv:
        agg[r[ik].split('(')[0]].append(float(r[iv].replace(',', '')))
tot = sum(sum(v) for v in agg.values())
for k, v in sorted(agg.items(), key=lambda x: -sum(x[1]))[:int(sys.argv[2]) if len(sys.argv) > 2 else 8]:
    print(f"{k[:44]:44s} n={len(v):4d} mean={sum(v)/len(v)/1e3:8.1f} us  total={sum(v)/1e6:8.3f} ms {100*sum(v)/tot:5.1f}%")
