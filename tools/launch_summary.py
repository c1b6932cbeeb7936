"""Summarise an ncu launch list (--metrics gpu__time_duration.sum --csv): per-kernel totals and the sequence."""
import csv, re, sys, collections
path = sys.argv[1]; pat = sys.argv[2] if len(sys.argv) > 2 else ""
lines = [l for l in open(path) if l.startswith('"')]
r = csv.reader(lines); hdr = next(r)
ki, vi, gi = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Grid Size")
seq = []
for row in r:
    name = re.sub(r"\(.*", "", row[ki]).replace("<unnamed>::", "").replace("void ", "")[:50]
    seq.append((name, row[gi], float(row[vi].replace(",", "")) / 1e3))
sel = [s for s in seq if pat in s[0]]
tot = collections.OrderedDict()
for s in sel:
    t = tot.setdefault(s[0], [0, 0.0]); t[0] += 1; t[1] += s[2]
all_t = sum(v[1] for v in tot.values())
for k, v in tot.items():
    print(f"{k:50s} n={v[0]:4d} total={v[1]:10.1f} us  share={100*v[1]/all_t:5.1f}%")
if "--seq" in sys.argv:
    for s in sel: print(f"  {s[0]:40s} {s[1]:16s} {s[2]:9.1f}")
