#!/usr/bin/env python
"""Benchmark of the batched Newton-KKT step (BASELINE.json metric) on B200.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference]

One step = one full Newton-KKT step (residual + active set, KKT assembly, factorisation, substitution, step
finish, residual norm at the new point: SURVEY.md 8a a1-a17) over a batch of 4096 random dense convex QPs
with n=512, m=256 (cfg3 of BASELINE.json) per GPU.  Prints ONE JSON line (rank 0).
"""

from __future__ import annotations

import argparse
import json
import math
import os
import subprocess
import sys
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402

N_VARS, N_CONS, BATCH = 512, 256, 4096
METRIC = "batched Newton-KKT steps/sec at n=512, batch 4096"
UNIT = "steps/s"
LAMB, RHO = 1.0, 1e-2
CPU_SAMPLE = 128


# ----------------------------------------------------------------------------------------------- data
def make_batch_torch(B, n, m, device, seed):
    """cfg3 generator with the torch CUDA generator: H = sym(MM'/n) + 0.1 I, A ~ N(0,1),
    b = -A x_f, g = 0.3 N(0,1), bounds [-1, 1]."""
    import torch

    gen = torch.Generator(device=device)
    gen.manual_seed(2000 + seed)
    f64 = dict(dtype=torch.float64, device=device)
    H = torch.empty((B, n, n), **f64)
    for lo in range(0, B, 256):
        hi = min(B, lo + 256)
        M = torch.randn((hi - lo, n, n), generator=gen, **f64)
        G = torch.bmm(M, M.transpose(1, 2)) / n
        H[lo:hi] = 0.5 * (G + G.transpose(1, 2)) + 0.1 * torch.eye(n, **f64)
        del M, G
    A = torch.randn((B, m, n), generator=gen, **f64)
    xf = torch.rand((B, n), generator=gen, **f64) - 0.5
    b = -torch.bmm(A, xf.unsqueeze(2)).squeeze(2)
    g = 0.3 * torch.randn((B, n), generator=gen, **f64)
    lb = torch.full((B, n), -1.0, **f64)
    ub = torch.full((B, n), 1.0, **f64)
    return dict(H=H, A=A, g=g, b=b, lb=lb, ub=ub)


def make_sample_numpy(S, n, m):
    """Same workload generated on the host (used when no GPU data is at hand: --impl reference)."""
    from pygradflow_b200 import synth

    d = synth.qp_batch(range(S), n, m)
    rng = np.random.default_rng(7)
    d["x"] = np.clip(0.3 * rng.uniform(-1, 1, (S, n)), -1.0, 1.0)
    d["y"] = 0.1 * rng.standard_normal((S, m))
    d["lamb"] = np.full(S, LAMB)
    d["rho"] = np.full(S, RHO)
    return d


# ----------------------------------------------------------------------------------------------- CPU arm
def _cpu_step(args):
    """One full Newton-KKT step of one instance through the oracle (the reference's algorithm, SuperLU)."""
    from threadpoolctl import threadpool_limits

    from oracle import gradflow_oracle as orc

    H, A, g, b, lb, ub, x, y, lamb, rho = args
    with threadpool_limits(limits=1):
        t0 = time.perf_counter()
        prob = orc.DenseQP(H, A, g, b, lb, ub)
        prm = orc.OracleParams(newton_type="full", linear_solver="splu")
        it = orc.Iterate(prob, prm, x, y)
        dt = 1.0 / float(np.asarray(lamb).reshape(()))
        rho = float(np.asarray(rho).reshape(()))
        res = orc.newton_method(prob, prm, it, dt, rho).step(it)
        nxt = res.iterate
        fn = float(np.linalg.norm(orc.ImplicitFunc(prob, it, dt).value_at(nxt, rho)))
        t1 = time.perf_counter()
    return nxt.x, nxt.y, res.diff, fn, t1 - t0


def cpu_steps(sample, cores):
    """Run the sample through a process pool (one instance per task, like the reference's runner:
    pygradflow/runners/runner.py:107-153).  Returns (results, wall seconds)."""
    import multiprocessing as mp

    S = sample["x"].shape[0]
    tasks = [tuple(np.ascontiguousarray(sample[k][i]) for k in ("H", "A", "g", "b", "lb", "ub", "x", "y", "lamb", "rho"))
             for i in range(S)]
    ctx = mp.get_context("fork")
    with ctx.Pool(cores) as pool:
        pool.map(_cpu_noop, range(cores))  # start the workers before timing
        t0 = time.perf_counter()
        res = pool.map(_cpu_step, tasks, chunksize=1)
        wall = time.perf_counter() - t0
    return res, wall


def _cpu_noop(i):
    import scipy.sparse.linalg  # noqa: F401

    return i


def host_cores():
    try:
        return len(os.sched_getaffinity(0))
    except Exception:
        return os.cpu_count() or 1


def run_reference_arm(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    cores = host_cores()
    sample = make_sample_numpy(CPU_SAMPLE, N_VARS, N_CONS)
    for _ in range(min(args.warmup, 1)):
        cpu_steps({k: v[: 2 * cores] for k, v in sample.items()}, cores)
    walls = []
    for _ in range(args.steps):
        _, wall = cpu_steps(sample, cores)
        walls.append(wall)
    total = sum(walls)
    value = CPU_SAMPLE * args.steps / total
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * total / args.steps,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": "cfg3: random dense convex QPs n=512 m=256, equality + bound constraints, "
                               "one full Newton-KKT step per instance", "n": N_VARS, "m": N_CONS,
                   "batch_per_step": CPU_SAMPLE, "lamb": LAMB, "rho": RHO},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": "port",
                         "sample": f"{CPU_SAMPLE} instances of the cfg3 workload per step (of 4096), oracle port of "
                                   "the reference path (dense NumPy assembly + scipy splu = SuperLU, as "
                                   "lu_solver.py:14), one process per core"},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


# ----------------------------------------------------------------------------------------------- GPU arm
class ClockSampler:
    FIELDS = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
              "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
              "clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.proc = None
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", f"--id={index}", f"--query-gpu={self.FIELDS}", "--format=csv,noheader,nounits",
                 "-lms", "100"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
        except Exception:
            self.proc = None

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            out, _ = self.proc.communicate(timeout=5)
        except Exception:
            self.proc.kill()
            out = ""
        sm, mx, reasons, power = [], [], set(), []
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for ln in out.strip().splitlines():
            parts = [p.strip() for p in ln.split(",")]
            if len(parts) < 7:
                continue
            try:
                sm.append(float(parts[0]))
                mx.append(float(parts[1]))
                power.append(float(parts[2]))
            except ValueError:
                continue
            for nm, v in zip(names, parts[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(nm)
        # keep the samples taken under load (upper half by power)
        if sm:
            order = np.argsort(power)[len(power) // 2:]
            load = [sm[i] for i in order]
        else:
            load = []
        return {"sm_mhz": float(np.median(load)) if load else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


def measure_fp64_peak(device):
    import torch

    N = 4096
    a = torch.randn(N, N, device=device, dtype=torch.float64)
    b = torch.randn(N, N, device=device, dtype=torch.float64)
    torch.matmul(a, b)
    best = 1e9
    for _ in range(5):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        torch.matmul(a, b)
        e1.record()
        torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1))
    return 2 * N ** 3 / best * 1e-9  # TFLOP/s


def ncu_traffic():
    """DRAM bytes (read + write) of one whole factorisation of the bench batch, summed over its launches, from the
    committed ncu capture (profiles/r01_ncu_ldlt_factor_dram.json); None when the capture is not at hand."""
    try:
        with open(os.path.join(ROOT, "profiles", "r01_ncu_ldlt_factor_dram.json")) as f:
            return float(json.load(f)["dram_bytes_per_factorisation"])
    except Exception:
        return None


def run_gpu_arm(args):
    import torch
    import torch.distributed as dist

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device (there is no CPU fallback)")
    torch.cuda.set_device(local_rank)
    device = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=device)

    from pygradflow_b200 import kernels as K
    from pygradflow_b200.host_step import HostNewtonKKT
    from pygradflow_b200.newton import NewtonKKTStepper
    from pygradflow_b200.params import LinearSolverType
    from pygradflow_b200.problem import BatchedQP

    B, n, m = BATCH, N_VARS, N_CONS
    data = make_batch_torch(B, n, m, device, seed=rank)
    prob = object.__new__(BatchedQP)
    prob.var_lb, prob.var_ub = data["lb"], data["ub"]
    prob.B, prob.n, prob.m, prob.device = B, n, m, device
    prob.H, prob.A, prob.g, prob.b = data["H"], data["A"], data["g"], data["b"]
    f64 = dict(dtype=torch.float64, device=device)
    lamb = torch.full((B,), LAMB, **f64)
    rho = torch.full((B,), RHO, **f64)
    linear = LinearSolverType[args.linear]
    # State of the timed step: every instance's iterate after `--state-iters` outer iterations of the batched
    # solver from the cfg3 start (x0 = 0, y0 = 0, default Params) -- a mid-solve point with its own lambda, rho and
    # a developed active set -- instead of an arbitrary random point.
    from pygradflow_b200.params import Params
    from pygradflow_b200.solver import BatchedSolver

    bs = BatchedSolver(prob, Params(linear_solver_type=linear))
    bs.solve(None, None, max_outer=args.state_iters)
    x, y = bs.cur[0].clone(), bs.cur[1].clone()
    lamb, rho = bs.lamb.clone(), bs.rho.clone()
    running = int((bs.status == 0).sum().item())
    del bs
    torch.cuda.empty_cache()
    data["x"], data["y"] = x, y
    stepper = NewtonKKTStepper(prob, linear)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for _ in range(args.warmup):
        stepper.step(x, y, lamb, rho)
    barrier()
    fp64_peak = measure_fp64_peak(device) if rank == 0 else None
    stepper.enable_timing()
    sampler = ClockSampler(local_rank) if rank == 0 else None
    launches0 = K.LAUNCHES
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(args.steps):
        out = stepper.step(x, y, lamb, rho)
    e1.record()
    barrier()
    ms = e0.elapsed_time(e1)
    launches = K.LAUNCHES - launches0
    clocks = sampler.stop() if sampler else None
    t = torch.tensor([ms], dtype=torch.float64, device=device)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms_max = float(t.item())
    value = B * world * args.steps / (ms_max * 1e-3)
    phases = stepper.phase_ms()
    eng = stepper.engine
    info = out[4]
    Nvec = eng.Nvec.to(torch.float64)
    flops_ldlt = float((Nvec ** 3 / 3.0).sum().item())
    n_fallback = int((eng.fbkey != 0).sum().item()) if linear != LinearSolverType.LU else 0
    nfail = int((info != 0).sum().item())
    xn_gpu, yn_gpu, diff_gpu, fn_gpu = (o.clone() for o in out[:4])

    # ---- end to end through host buffers (pinned), chunked + double buffered
    e2e = None
    try:
        hk = HostNewtonKKT(n, m, chunk=256, device=device, linear=linear)
        host = {k: HostNewtonKKT.pinned_like(v) for k, v in
                dict(H=data["H"], A=data["A"], g=data["g"], b=data["b"], lb=data["lb"], ub=data["ub"], x=x, y=y,
                     lamb=lamb, rho=rho).items()}
        outp = {"xn": torch.empty((B, n), dtype=torch.float64, pin_memory=True),
                "yn": torch.empty((B, m), dtype=torch.float64, pin_memory=True),
                "diff": torch.empty((B,), dtype=torch.float64, pin_memory=True),
                "fnorm": torch.empty((B,), dtype=torch.float64, pin_memory=True),
                "info": torch.empty((B,), dtype=torch.int32, pin_memory=True)}
        esteps = max(1, min(args.steps, 3))
        hk.step(host, outp)
        barrier()
        t0 = time.perf_counter()
        a0, a1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a0.record()
        for _ in range(esteps):
            hk.step(host, outp)
        a1.record()
        barrier()
        wall = time.perf_counter() - t0
        ems = torch.tensor([a0.elapsed_time(a1)], dtype=torch.float64, device=device)
        if world > 1:
            dist.all_reduce(ems, op=dist.ReduceOp.MAX)
        h2d, d2h = hk.bytes_per_step(B)
        e2e = {"value": B * world * esteps / (float(ems.item()) * 1e-3), "unit": UNIT,
               "h2d_bytes_per_step": h2d * world, "d2h_bytes_per_step": d2h * world, "steps": esteps,
               "wall_s": wall, "chunk": 256,
               "max_abs_diff_vs_device_path": float((outp["xn"].to(device) - xn_gpu).abs().max().item())}
        del hk, host
    except Exception as exc:  # pragma: no cover
        e2e = {"value": None, "unit": UNIT, "error": repr(exc)}

    cpu_baseline = None
    parity = None
    if rank == 0 and world == 1 and not args.no_cpu:
        cores = host_cores()
        S = CPU_SAMPLE
        sample = {k: data[k][:S].cpu().numpy() for k in ("H", "A", "g", "b", "lb", "ub", "x", "y")}
        sample["lamb"], sample["rho"] = lamb[:S].cpu().numpy(), rho[:S].cpu().numpy()
        res, wall = cpu_steps(sample, cores)
        cpu_value = S / wall
        xr = np.stack([r[0] for r in res])
        yr = np.stack([r[1] for r in res])
        dr = np.array([r[2] for r in res])
        rel = lambda a, b_: float(np.max(np.abs(a - b_)) / max(1.0, np.max(np.abs(b_))))
        parity = {"instances": S, "x_rel": rel(xn_gpu[:S].cpu().numpy(), xr), "y_rel": rel(yn_gpu[:S].cpu().numpy(), yr),
                  "diff_rel": rel(diff_gpu[:S].cpu().numpy(), dr)}
        cpu_baseline = {"value": cpu_value, "unit": UNIT, "cores": cores, "kind": "port",
                        "sample": f"first {S} of the {B} instances (identical inputs, copied back from the GPU), one "
                                  "full Newton-KKT step each through the oracle port (dense NumPy assembly + scipy "
                                  "splu = SuperLU as lu_solver.py:14), one process per core",
                        "sum_cpu_seconds": float(sum(r[4] for r in res)), "wall_s": wall}

    if rank == 0:
        # the roofline is that of the factorisation kernels: events around gf_ldlt_factor alone (the `factor` phase
        # of the step additionally holds the bookkeeping of the pivoted-LU fallback list)
        factor_ms = stepper.ldlt_ms() or phases["factor"]
        achieved = flops_ldlt / (factor_ms * 1e-3) * 1e-12 if linear != LinearSolverType.LU else \
            2.0 * flops_ldlt / (factor_ms * 1e-3) * 1e-12
        hbm_peak = None
        try:
            hbm_peak = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"]
        except Exception:
            hbm_peak = 6650.0
        nIbar = float(eng.nI.to(torch.float64).mean().item())
        Nbar = float(Nvec.mean().item())
        # algorithmic HBM bytes of the HBM-bound phases, per batched step
        nId = eng.nI.double()
        if eng.linear == LinearSolverType.LDLT:  # lower triangle only: half of H_II read, half of K written
            asm_bytes = float(((nId ** 2 / 2 + m * nId) * 8 + n + (Nvec ** 2) * 4).sum().item())
        else:
            asm_bytes = float(((nId ** 2 + m * nId) * 8 + n + (Nvec ** 2) * 8).sum().item())
        solve_bytes = float(((Nvec ** 2) * 8 + 16 * Nvec).sum().item())
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms_max / args.steps, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": "cfg3: batch of 4096 random dense convex QPs n=512 m=256 per GPU, equality + "
                                   "bound constraints (active-set KKT); one full Newton-KKT step per instance per "
                                   "step", "n": n, "m": m, "batch_per_gpu": B, "global_batch": B * world,
                       "state": f"iterates after {args.state_iters} outer iterations of the batched solver from x0=0, "
                                f"y0=0 (per-instance lambda, rho); {running} of {B} instances still running",
                       "linear_solver": eng.linear.name,
                       "mean_inactive": nIbar, "mean_kkt_order": Nbar, "lu_fallback_instances": n_fallback,
                       "failed_instances": nfail,
                       "l2": "inputs (13 GB of H, A per step) are far larger than the 126 MB L2; no flush needed"},
            "roofline": {"bound": "tensor", "kernel": "batched LDL' factorisation = gf_ldlt_factor: ldlt_diag0_kernel + one "
                                                      "ldlt_column_kernel (DMMA) per 64-wide block column, issued as "
                                                      "two interleaved half-batches",
                         "achieved": achieved, "peak": fp64_peak, "unit": "TFLOP/s",
                         "frac": achieved / fp64_peak if fp64_peak else None, "traffic": ncu_traffic(),
                         "peak_source": "FP64 cuBLAS DGEMM 4096^3 measured in this run (MEASURED_PEAKS.json has no "
                                        "FP64 entry; profiles/r01_fp64_peak.json: 35.4 TFLOP/s)",
                         "algorithmic_flops_per_step": flops_ldlt if linear != LinearSolverType.LU else 2 * flops_ldlt,
                         "ms": factor_ms},
            "roofline_hbm": {
                "peak": hbm_peak, "unit": "GB/s",
                "assemble": {"bytes": asm_bytes, "ms": phases["assemble"],
                             "achieved": asm_bytes / (phases["assemble"] * 1e-3) * 1e-9,
                             "frac": asm_bytes / (phases["assemble"] * 1e-3) * 1e-9 / hbm_peak},
                "solve": {"bytes": solve_bytes, "ms": phases["solve"],
                          "achieved": solve_bytes / (phases["solve"] * 1e-3) * 1e-9,
                          "frac": solve_bytes / (phases["solve"] * 1e-3) * 1e-9 / hbm_peak},
            },
            "phase_ms": phases,
            "clocks": clocks,
            "e2e": e2e,
            "gpu_launches": launches,
            "cpu_baseline": cpu_baseline,
            "parity_vs_cpu_sample": parity,
        }
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--linear", default="Auto", choices=["Auto", "LU", "LDLT"])
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg")
    ap.add_argument("--state-iters", type=int, default=6, help="outer iterations run to reach the timed state")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference_arm(args)
    else:
        run_gpu_arm(args)


if __name__ == "__main__":
    main()
