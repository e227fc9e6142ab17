"""Per-source-line stall samples / instruction counts from an .ncu-rep (needs -lineinfo and --import-source on).

    python scripts/ncu_lines.py report.ncu-rep <kernel-regex> [top N]
"""
import collections, csv, subprocess, sys
rep, kern = sys.argv[1], sys.argv[2]
top = int(sys.argv[3]) if len(sys.argv) > 3 else 40
txt = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass", "--kernel-name",
                      f"regex:{kern}", "--launch-count", "1"], capture_output=True, text=True).stdout
rows = list(csv.reader(txt.splitlines()))
cur = hdr = None
agg = collections.defaultdict(lambda: [0, 0, collections.Counter(), ""])
STALLS = ("stall_barrier", "stall_long_sb", "stall_short_sb", "stall_wait", "stall_lg", "stall_mio", "stall_math",
          "stall_not_selected", "stall_selected", "stall_sleep", "stall_membar", "stall_branch_resolving",
          "stall_dispatch", "stall_tex", "stall_drain")
for r in rows:
    if r and r[0] == "File Path":
        cur, hdr = r[1], None
        continue
    if r and r[0] == "Line No":
        hdr = r
        continue
    if hdr is None or cur is None or len(r) < len(hdr) // 2 or not cur.endswith((".inc", ".cu", ".cuh")):
        continue
    try:
        ln, s, e = int(r[0]), int(r[hdr.index("# Samples")]), int(r[hdr.index("Instructions Executed")])
    except ValueError:
        continue
    if s == 0 and e == 0:
        continue
    a = agg[(cur.split("/")[-1], ln)]
    a[0] += s
    a[1] += e
    a[3] = r[1].strip()[:70]
    for k in STALLS:
        try:
            a[2][k] += int(r[hdr.index(k)])
        except ValueError:
            pass
tot = sum(a[0] for a in agg.values())
tote = sum(a[1] for a in agg.values())
print("total samples", tot, "warp instructions", tote)
allst = collections.Counter()
for a in agg.values():
    allst.update(a[2])
print("stalls:", ", ".join(f"{k[6:]}:{100 * v / max(tot, 1):.1f}%" for k, v in allst.most_common(10)))
for (f, ln), a in sorted(agg.items(), key=lambda kv: -kv[1][0])[:top]:
    st = ", ".join(f"{k[6:]}:{v}" for k, v in a[2].most_common(3))
    print(f"{f:20s} {ln:4d} samp {100 * a[0] / tot:5.1f}% inst {100 * a[1] / tote:5.1f}%  {st:55s} | {a[3]}")
