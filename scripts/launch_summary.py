"""Per-kernel totals from an `ncu --metrics gpu__time_duration.sum --csv` launch list."""
import csv, sys, collections
rows = list(csv.reader(open(sys.argv[1], errors="ignore")))
hi = next(i for i, r in enumerate(rows) if "Kernel Name" in r)
h = rows[hi]; ki = h.index("Kernel Name"); vi = h.index("Metric Value"); gi = h.index("Grid Size"); bi = h.index("Block Size")
d = collections.OrderedDict()
for r in rows[hi + 1:]:
    if len(r) <= vi: continue
    try: v = float(r[vi].replace(",", ""))
    except ValueError: continue
    d.setdefault((r[ki][:70], r[gi], r[bi]), []).append(v)
tot = sum(sum(v) for v in d.values())
for (k, g, b), v in d.items():
    print(f"{k:70s} {g:>14s} {b:>12s} n={len(v):4d} avg={sum(v)/len(v)/1000:9.1f}us tot={sum(v)/1000:10.1f}us {100*sum(v)/tot:5.1f}%")
print("total us", tot / 1000)
