"""Time the full-order assembly modes (gf_kkt_assemble_full: Asymmetric / Extended step formulations) at the cfg3 shape
and report achieved HBM GB/s against the algorithmic bytes: read 8 (nI n + m n) (the inactive rows of H, and J once --
its second, transposed read hits L2), write 8 (n + m)^2.

    python tools/bench_assemble_full.py [--B 1024] [--n 512] [--m 256]
"""
import argparse, json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from pygradflow_b200 import kernels as K
from pygradflow_b200.engine import KKTEngine
from pygradflow_b200.params import StepSolverType

ap = argparse.ArgumentParser()
ap.add_argument("--B", type=int, default=1024)
ap.add_argument("--n", type=int, default=512)
ap.add_argument("--m", type=int, default=256)
args = ap.parse_args()
B, n, m = args.B, args.n, args.m
f64 = dict(dtype=torch.float64, device="cuda")
g = torch.Generator(device="cuda"); g.manual_seed(0)
H = torch.randn((B, n, n), generator=g, **f64); H = H + H.transpose(1, 2)
J = torch.randn((B, m, n), generator=g, **f64)
dt = torch.full((B,), 0.5, **f64); rho = torch.full((B,), 1e-2, **f64)
active = (torch.rand((B, n), generator=g, device="cuda") < 0.2).to(torch.uint8)
w = K.WorkList.all(B)
out = {}
for kind in ("Asymmetric", "Extended", "Symmetric"):
    eng = KKTEngine(B, n, m, "cuda", formulation=StepSolverType[kind]) if kind != "Symmetric" else KKTEngine(B, n, m, "cuda")
    eng.active.copy_(active); eng.update_active_set(w)
    nI = eng.nI.double().mean().item()
    ts = []
    for r in range(4):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); eng.assemble(H, J, dt, rho, w); e1.record(); torch.cuda.synchronize()
        if r: ts.append(e0.elapsed_time(e1))
    ms = min(ts)
    if kind == "Symmetric":
        N = nI + m
        bytes_ = B * 8 * (nI * nI / 2 + m * nI + N * N / 2)   # lower triangle only
    else:
        bytes_ = B * 8 * (nI * n + m * n + (n + m) ** 2)
    out[kind] = dict(ms=ms, algorithmic_GB=bytes_ / 1e9, GBs=bytes_ / ms / 1e6)
print(json.dumps(dict(B=B, n=n, m=m, mean_inactive=nI, **out)))
