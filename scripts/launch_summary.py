"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list: python scripts/launch_summary.py csv "title" """
import collections, csv, sys
rows = [r for r in csv.reader(l for l in open(sys.argv[1]) if l.startswith('"'))]
h = rows[0]; k = h.index("Kernel Name"); v = h.index("Metric Value")
agg = collections.OrderedDict()
for r in rows[1:]:
    n = r[k].split("(")[0].replace("void ", "").replace("ttn::", "")
    a = agg.setdefault(n, [0, 0.0]); a[0] += 1; a[1] += float(r[v]) / 1e6
tot = sum(a[1] for a in agg.values())
print(f"# {sys.argv[2] if len(sys.argv) > 2 else sys.argv[1]}")
print(f"# {sum(a[0] for a in agg.values())} launches, {tot:.2f} ms total (cold-cache, serialised, clocks not boosted: compare shares)")
for n, a in sorted(agg.items(), key=lambda x: -x[1][1]):
    print(f"{n[:70]:70s} launches {a[0]:3d} {a[1]:10.3f} ms {100 * a[1] / tot:5.1f}%")
