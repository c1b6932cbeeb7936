"""Per-kernel shares of the LAST device-resident Newton-KKT step in an ncu launch list of bench.py
(--metrics gpu__time_duration.sum --csv):  python tools/step_launches.py launches.csv [B]"""
import csv, re, sys, collections
path = sys.argv[1]; B = int(sys.argv[2]) if len(sys.argv) > 2 else 4096
lines = [l for l in open(path) if l.startswith('"')]
r = csv.reader(lines); hdr = next(r)
ki, vi, gi = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Grid Size")
seq = []
for row in r:
    name = re.sub(r"\(.*", "", row[ki]).replace("<unnamed>::", "").replace("void ", "")[:50]
    seq.append((name, row[gi], float(row[vi].replace(",", "")) / 1e3))
starts = [i for i, s in enumerate(seq) if s[0] == "dt_from_lamb_kernel" and s[1].startswith(f"({(B + 127) // 128},")]
lo = starts[-1]
hi = lo + 1
while hi < len(seq) and seq[hi][0] != "dt_from_lamb_kernel" and not seq[hi][0].startswith("cutlass"):
    hi += 1
# the step ends with the residual norm at the new point
while hi > lo and seq[hi - 1][0] != "residual_kernel":
    hi -= 1
step = seq[lo:hi]
tot = collections.OrderedDict()
for s in step:
    t = tot.setdefault(s[0], [0, 0.0]); t[0] += 1; t[1] += s[2]
all_t = sum(v[1] for v in tot.values())
print(f"# last device-resident Newton-KKT step over {B} instances: {len(step)} launches, sum of launch durations {all_t / 1e3:.2f} ms")
print("# (ncu serialises launches and runs them cold-cache: compare SHARES with bench.py's phase_ms, not absolutes)")
for k, v in sorted(tot.items(), key=lambda kv: -kv[1][1]):
    print(f"{k:50s} launches={v[0]:3d} total_us={v[1]:10.1f} share={100 * v[1] / all_t:5.1f}%")
