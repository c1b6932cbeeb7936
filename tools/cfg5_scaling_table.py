"""Strong-scaling table of the cfg5 sweeps: python tools/cfg5_scaling_table.py  (reads profiles/r02_cfg5_sweep_{1,2,4,8}.json,
writes profiles/r02_cfg5_scaling.md)."""
import json
import os

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
rows = {}
for G in (1, 2, 4, 8):
    for r in json.load(open(os.path.join(ROOT, "profiles", f"r02_cfg5_sweep_{G}.json")))["rows"]:
        rows[(r["method"], r["N"], r["B"], G)] = r
md = ["cfg5 strong scaling (tools/sweep_cfg5.py under torchrun): the case's batch is block-sharded over the GPUs, time = max over ranks.",
      "factor+solve per second of the WHOLE job; efficiency = rate(G) / (G x rate(1)).  Cases that exist at all four GPU counts.", "",
      "| method | N | B | 1 GPU /s | 2 GPUs /s | 4 GPUs /s | 8 GPUs /s | eff 2 | eff 4 | eff 8 | factor TFLOP/s at 8 |",
      "|---|---|---|---|---|---|---|---|---|---|---|"]
for (meth, N, B, G) in sorted(k for k in rows if k[3] == 1):
    if not all((meth, N, B, g) in rows for g in (2, 4, 8)) or B < 1024:
        continue
    v = [rows[(meth, N, B, g)]["factor_solve_per_s"] for g in (1, 2, 4, 8)]
    md.append(f"| {meth} | {N} | {B} | {v[0]:.0f} | {v[1]:.0f} | {v[2]:.0f} | {v[3]:.0f} | {v[1] / v[0] / 2:.2f} | "
              f"{v[2] / v[0] / 4:.2f} | {v[3] / v[0] / 8:.2f} | {rows[(meth, N, B, 8)]['factor_gflops'] * 1e-3:.1f} |")
extra = [k for k in rows if k[3] == 8 and (k[0], k[1], k[2], 1) not in rows]
md += ["", "Cases that only fit with 8 GPUs (B N^2 8 bytes <= 64 GB per GPU):", ""]
for k in sorted(extra):
    r = rows[k]
    md.append(f"* {k[0]} N={k[1]} B={k[2]}: factor {r['factor_ms']:.1f} ms ({r['factor_gflops'] * 1e-3:.1f} TFLOP/s), solve "
              f"{r['solve_ms']:.2f} ms ({r['solve_gbs']:.0f} GB/s), {r['factor_solve_per_s']:.0f} factor+solve/s")
open(os.path.join(ROOT, "profiles", "r02_cfg5_scaling.md"), "w").write("\n".join(md) + "\n")
print("\n".join(md[:8]))
