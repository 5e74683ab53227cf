"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list: python scripts/launch_summary.py file.csv"""
import collections, csv, re, sys
rows = list(csv.reader(open(sys.argv[1])))
hi = next(i for i, r in enumerate(rows) if 'Kernel Name' in r)
hdr, data = rows[hi], rows[hi + 1:]
kn, mv, mn, mu = (hdr.index(k) for k in ('Kernel Name', 'Metric Value', 'Metric Name', 'Metric Unit'))
agg = collections.defaultdict(lambda: [0, 0.0])
tot = 0.0
for r in data:
    if len(r) <= mv or r[mn] != 'gpu__time_duration.sum':
        continue
    name = re.sub(r'\(.*', '', r[kn])[:80]
    v = float(r[mv].replace(',', ''))
    v = v / 1000 if r[mu] == 'ns' else v
    agg[name][0] += 1
    agg[name][1] += v
    tot += v
print(f"total {tot / 1000:.2f} ms over {sum(a[0] for a in agg.values())} launches (cold-cache, serialised: compare shares)")
print("| kernel | launches | total ms | share | avg us |\n|---|---|---|---|---|")
for k, (n, t) in sorted(agg.items(), key=lambda kv: -kv[1][1])[:14]:
    print(f"| {k} | {n} | {t / 1000:.3f} | {100 * t / tot:.1f}% | {t / n:.1f} |")
