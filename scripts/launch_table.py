"""Per-kernel totals of an ncu launch list (`ncu --metrics gpu__time_duration.sum --csv --log-file X`).  usage: python scripts/launch_table.py X.csv"""
import csv, collections, sys
rows = [r for r in csv.reader(open(sys.argv[1])) if len(r) > 10]
hdr = rows[0]; ki, vi, ui = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
agg = collections.OrderedDict()
for r in rows[1:]:
    n = r[ki].split('(')[0].replace("void ", ""); v = float(r[vi].replace(',', ''))
    v = v / 1e3 if r[ui] == 'ns' else (v * 1e3 if r[ui] == 'ms' else v)
    a = agg.setdefault(n, [0, 0.0]); a[0] += 1; a[1] += v
tot = sum(a[1] for a in agg.values())
for n, (c, t) in sorted(agg.items(), key=lambda x: -x[1][1]):
    print(f"{n:60s} {c:4d} {t:10.1f} us  avg {t / c:9.1f}  share {t / tot:.3f}")
print(f"{len(rows) - 1} launches, {tot / 1e3:.2f} ms")
