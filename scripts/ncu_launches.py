"""Aggregate `ncu --metrics gpu__time_duration.sum --csv` launch lists by kernel: count, total, share."""
import csv, sys, collections
rows = list(csv.reader(open(sys.argv[1])))
hdr = next(i for i, r in enumerate(rows) if "Kernel Name" in r)
names = rows[hdr]
ki, vi, ui = names.index("Kernel Name"), names.index("Metric Value"), names.index("Metric Unit")
tot = collections.OrderedDict()
for r in rows[hdr + 1:]:
    if len(r) != len(names):
        continue
    v = float(r[vi].replace(",", ""))
    if r[ui] in ("usecond", "us"):
        v *= 1000.0
    elif r[ui] in ("msecond", "ms"):
        v *= 1e6
    k = r[ki].split("(")[0]
    c, t = tot.get(k, (0, 0.0))
    tot[k] = (c + 1, t + v)
total = sum(t for _, t in tot.values())
print(f"# {sys.argv[1]}: {sum(c for c, _ in tot.values())} launches, {total / 1e3:.1f} us total (ncu-serialised, cold cache)")
print(f"{'kernel':70s} {'launches':>8s} {'total_us':>10s} {'share':>7s}")
for k, (c, t) in sorted(tot.items(), key=lambda kv: -kv[1][1]):
    print(f"{k[:70]:70s} {c:8d} {t / 1e3:10.1f} {t / total * 100:6.1f}%")
