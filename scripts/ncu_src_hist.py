"""Dynamic instruction histogram (by opcode and by hot SASS region) from `ncu --page source --csv` output."""
import csv, sys, collections
rows = list(csv.reader(open(sys.argv[1])))
hi = next(i for i, r in enumerate(rows) if "Instructions Executed" in r)
h = rows[hi]
ie = h.index("Instructions Executed"); isrc = h.index("Source"); isam = h.index("# Samples")
body = [r for r in rows[hi + 1:] if len(r) > ie and r[ie].isdigit()]
tot = sum(int(r[ie]) for r in body); stot = sum(int(r[isam]) for r in body)
print("total warp-inst", tot, "samples", stot)
op = collections.Counter(); sm = collections.Counter()
for r in body:
    m = r[isrc].split()
    o = m[0] if not m[0].startswith('@') else m[1]
    op[o.split('.')[0]] += int(r[ie]); sm[o.split('.')[0]] += int(r[isam])
for k, v in op.most_common(int(sys.argv[2]) if len(sys.argv) > 2 else 30):
    print(f"{k:10s} {v:10d} {100*v/tot:5.1f}%  samples {100*sm[k]/max(stot,1):5.1f}%")
