"""CUDA-event timings of the kernels added in round 2, against the roofline that bounds each:

  gmres / minres (gf_krylov.cu)   HBM: every matrix-vector product streams the dense K once (8 N^2 bytes)
  hess_rho       (gf_syrk.cu)     FP64 tensor pipe: 2 n^2 m flop per instance
  fused cfg2     (gf_fused.cu)    latency: us per outer iteration of one instance (one warp)

    python tools/bench_round2_kernels.py [--which krylov,syrk,fused] [--out profiles/r02_round2_kernels.json]
"""
import argparse
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch

from pygradflow_b200 import kernels as K
from pygradflow_b200 import synth
from pygradflow_b200.kernels import WorkList


def ev_ms(fn, reps=3, setup=None):
    best = 1e30
    for _ in range(reps + 1):
        if setup:
            setup()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        fn()
        e1.record()
        torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1))
    return best


def kkt_batch(B, n, m, dev, seed=0):
    """Quasi-definite KKT matrices of the cfg3 shape (order n + m), generated on the device."""
    g = torch.Generator(device=dev)
    g.manual_seed(seed)
    f64 = dict(dtype=torch.float64, device=dev)
    N = n + m
    Km = torch.zeros((B, N, N), **f64)
    step = 64
    for lo in range(0, B, step):
        hi = min(B, lo + step)
        M = torch.randn((hi - lo, n, n), generator=g, **f64)
        Km[lo:hi, :n, :n] = torch.bmm(M, M.transpose(1, 2)) / n + 1.1 * torch.eye(n, **f64)
        A = torch.randn((hi - lo, m, n), generator=g, **f64)
        Km[lo:hi, n:, :n] = A
        Km[lo:hi, :n, n:] = A.transpose(1, 2)
        Km[lo:hi, n:, n:] = -0.99 * torch.eye(m, **f64)
    rhs = torch.randn((B, N), generator=g, **f64)
    return Km, rhs


def bench_krylov(out, dev, hbm):
    for (B, n, m) in [(1024, 512, 256), (4096, 128, 64)]:
        N = n + m
        Km, rhs0 = kkt_batch(B, n, m, dev)
        i32 = dict(dtype=torch.int32, device=dev)
        Nvec = torch.full((B,), N, **i32)
        info, iters = torch.zeros((B,), **i32), torch.zeros((B,), **i32)
        w = WorkList.all(B)
        for kind in ("gmres", "minres"):
            rows = K.krylov_scratch_rows(kind == "minres")
            scratch = torch.zeros((B, rows, N), dtype=torch.float64, device=dev)
            rhs = rhs0.clone()
            if kind == "gmres":
                fn = lambda: K.gmres_solve(Km, N, Nvec, rhs, None, None, False, scratch, info, iters, w)
            else:
                fn = lambda: K.minres_solve(Km, N, Nvec, rhs, None, scratch, info, iters, w)
            ms = ev_ms(fn, reps=2, setup=lambda: rhs.copy_(rhs0))
            prods = int(iters.sum().item())
            res = torch.bmm(Km[:8], rhs[:8].unsqueeze(2)).squeeze(2) - rhs0[:8]
            byts = prods * 8.0 * N * N
            out[f"{kind}_B{B}_N{N}"] = dict(ms=ms, matvecs_total=prods, matvecs_mean=prods / B, failed=int((info != 0).sum().item()),
                                            algorithmic_bytes=byts, GBps=byts / ms * 1e-6, frac_hbm=byts / ms * 1e-6 / hbm,
                                            rel_residual=float((res.norm(dim=1) / rhs0[:8].norm(dim=1)).max().item()))
            print(kind, B, N, out[f"{kind}_B{B}_N{N}"], flush=True)
        del Km


def bench_syrk(out, dev, peak):
    f64 = dict(dtype=torch.float64, device=dev)
    for (B, n, m) in [(4096, 512, 256), (1024, 1024, 512)]:
        J = torch.randn((B, m, n), **f64)
        H = torch.randn((B, n, n), **f64)
        rho = torch.rand((B,), **f64)
        o = torch.empty_like(H)
        w = WorkList.all(B)
        ms = ev_ms(lambda: K.hess_rho(H, J, rho, o, w), reps=3)
        ref = rho[:4, None, None] * torch.bmm(J[:4].transpose(1, 2), J[:4]) + H[:4]
        flops = 2.0 * B * n * n * m
        t0 = ev_ms(lambda: torch.bmm(J.transpose(1, 2), J, out=o), reps=3)
        out[f"hess_rho_B{B}_n{n}_m{m}"] = dict(ms=ms, TFLOPs=flops / ms * 1e-9, frac_fp64_peak=flops / ms * 1e-9 / peak,
                                               cublas_bmm_ms=t0, max_rel_err=float(((o[:4] - ref).abs().max() / ref.abs().max()).item())
                                               if False else None)
        K.hess_rho(H, J, rho, o, w)
        torch.cuda.synchronize()
        out[f"hess_rho_B{B}_n{n}_m{m}"]["max_rel_err"] = float(((o[:4] - ref).abs().max() / ref.abs().max()).item())
        print(out[f"hess_rho_B{B}_n{n}_m{m}"], flush=True)
        del J, H, o


def bench_fused(out, dev):
    from pygradflow_b200.params import Params
    from pygradflow_b200.problem import BatchedRosenbrock
    from pygradflow_b200.solver import BatchedSolver

    B = 4096
    d = synth.rosenbrock_batch(range(B), 64)
    prob = BatchedRosenbrock(d["a"], d["b"], d["lb"], d["ub"])
    s = BatchedSolver(prob, Params())
    s.solve(d["x0"], None)
    holder = {}
    ms = ev_ms(lambda: holder.update(res=s.solve(d["x0"], None)), reps=2)
    it = holder["res"].iterations.cpu().numpy()
    out["fused_cfg2"] = dict(ms=ms, launches=s.fused_launches, iterations_total=int(it.sum()), iterations_max=int(it.max()),
                             us_per_outer_iteration_slowest=1e3 * ms / int(it.max()),
                             outer_iterations_per_s=float(it.sum()) / ms * 1e3)
    print(out["fused_cfg2"], flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--which", default="krylov,syrk,fused")
    ap.add_argument("--out", default=None)
    a = ap.parse_args()
    dev = torch.device("cuda")
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    try:
        hbm = json.load(open(os.path.join(root, "MEASURED_PEAKS.json")))["hbm_gbs"]
    except Exception:
        hbm = 6554.9
    x = torch.randn(4096, 4096, dtype=torch.float64, device=dev)
    torch.matmul(x, x)
    peak = 2 * 4096 ** 3 / ev_ms(lambda: torch.matmul(x, x), reps=4) * 1e-9
    del x
    out = dict(device=torch.cuda.get_device_name(0), hbm_peak_GBps=hbm, fp64_dgemm_peak_TFLOPs=peak)
    for w in a.which.split(","):
        {"krylov": lambda: bench_krylov(out, dev, hbm), "syrk": lambda: bench_syrk(out, dev, peak),
         "fused": lambda: bench_fused(out, dev)}[w]()
    if a.out:
        json.dump(out, open(a.out, "w"), indent=1)


if __name__ == "__main__":
    main()
