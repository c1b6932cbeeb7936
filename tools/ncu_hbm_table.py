"""Per-kernel DRAM bytes / duration / achieved GB/s from an ncu --csv metrics log (largest-grid launch per kernel)."""
import csv, json, sys
rows = [dict(zip(h, x)) for h, *xs in [list(csv.reader([l for l in open(sys.argv[1]) if l.startswith('"')]))] for x in xs]
scale = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "ns": 1e-3, "us": 1, "usecond": 1, "ms": 1e3, "msecond": 1e3, "nsecond": 1e-3, "%": 1}
byid = {}
for d in rows:
    e = byid.setdefault(int(d["ID"]), {"kernel": d["Kernel Name"].split("(")[0].replace("<unnamed>::", ""), "grid": d["Grid Size"]})
    e[d["Metric Name"]] = float(d["Metric Value"].replace(",", "")) * scale.get(d["Metric Unit"], 1)
best = {}
for e in byid.values():
    t = e.get("gpu__time_duration.sum", 0)
    if e["kernel"] not in best or t > best[e["kernel"]]["gpu__time_duration.sum"]:
        best[e["kernel"]] = e
out = []
for k, e in sorted(best.items(), key=lambda kv: -kv[1]["gpu__time_duration.sum"]):
    rd, wr, t = e.get("dram__bytes_read.sum", 0), e.get("dram__bytes_write.sum", 0), e["gpu__time_duration.sum"]
    out.append(dict(kernel=k, grid=e["grid"], us=round(t, 1), dram_read_GB=round(rd * 1e-9, 3), dram_write_GB=round(wr * 1e-9, 3),
                    dram_GBps=round((rd + wr) / t * 1e-3, 1), dram_pct_of_peak=round(e.get("dram__throughput.avg.pct_of_peak_sustained_elapsed", 0), 1)))
print(json.dumps(out, indent=1))
