"""Sum dram__bytes_read/write and durations over the ldlt launches of one factorisation from an ncu --csv log
(tools/profile_factor.py --reps 1 runs two factorisations: the second half of the launches is taken)."""
import csv, json, sys
path, out = sys.argv[1], sys.argv[2]
lines = [l for l in open(path) if l.startswith('"')]
r = csv.reader(lines); hdr = next(r)
rows = [dict(zip(hdr, x)) for x in r]
byid = {}
for d in rows:
    if "ldlt_solve" in d["Kernel Name"]:
        continue
    e = byid.setdefault(int(d["ID"]), {"name": d["Kernel Name"].split("(")[0].replace("<unnamed>::", ""), "grid": d["Grid Size"]})
    v = float(d["Metric Value"].replace(",", ""))
    unit = d["Metric Unit"]
    scale = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "ns": 1e-3, "us": 1, "usecond": 1, "ms": 1e3, "msecond": 1e3, "nsecond": 1e-3}.get(unit, 1)
    e[d["Metric Name"]] = v * scale
ids = sorted(byid)
half = ids[len(ids) // 2:]
tot_r = sum(byid[i].get("dram__bytes_read.sum", 0) for i in half)
tot_w = sum(byid[i].get("dram__bytes_write.sum", 0) for i in half)
tot_t = sum(byid[i].get("gpu__time_duration.sum", 0) for i in half)
res = {"launches": len(half), "dram_bytes_read": tot_r, "dram_bytes_write": tot_w,
       "dram_bytes_per_factorisation": tot_r + tot_w, "sum_launch_us": tot_t,
       "per_launch": [dict(id=i, **byid[i]) for i in half]}
json.dump(res, open(out, "w"), indent=1)
print(json.dumps({k: v for k, v in res.items() if k != "per_launch"}))
