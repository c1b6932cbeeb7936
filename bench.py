#!/usr/bin/env python
"""Benchmark of the batched Newton-KKT step (BASELINE.json metric) on B200.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference] [--scaling weak|strong]

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
STATE_ITERS = 6


def bench_config(world: int, scaling: str):
    """`config` of the JSON line -- the SAME dict in both arms (the reference arm times a bounded sample of it)."""
    per_gpu = BATCH if scaling == "weak" else max(1, BATCH // world)
    return {
        "workload": "cfg3: batch of 4096 random dense convex QPs n=512 m=256, equality + bound constraints (active-set "
                    "KKT), SURVEY 8d generator rng=default_rng(2000+k) for instance k; one step = one full Newton-KKT "
                    "step of every instance (residual + active set, KKT assembly, factorisation, substitution, step "
                    "finish, residual norm at the new point) at that instance's iterate after 6 outer iterations of "
                    "Solver.solve from x0=0, y0=0 (default Params: per-instance lambda, rho)",
        "n": N_VARS, "m": N_CONS, "batch_per_gpu": per_gpu, "global_batch": per_gpu * world, "state_iters": STATE_ITERS,
        "l2": "inputs (13 GB of H, A per 4096 instances) are far larger than the 126 MB L2; no flush needed",
    }


def generate_host(k0: int, B: int, n: int, m: int, pinned: bool, threads: int):
    """Instances k0 .. k0+B-1 of the cfg3 generator (pygradflow_b200.synth.qp_instance, seeded per instance) written
    straight into (pinned) host tensors; one thread per core (NumPy's generators and BLAS release the GIL)."""
    import torch
    from concurrent.futures import ThreadPoolExecutor
    from threadpoolctl import threadpool_limits

    from pygradflow_b200 import synth

    def buf(*shape):
        return torch.empty(shape, dtype=torch.float64, pin_memory=pinned)

    host = dict(H=buf(B, n, n), A=buf(B, m, n), g=buf(B, n), b=buf(B, m), lb=buf(B, n), ub=buf(B, n))
    views = {k: v.numpy() for k, v in host.items()}

    def one(i):
        d = synth.qp_instance(k0 + i, n, m)
        for key in views:
            views[key][i] = d[key]

    with threadpool_limits(limits=1):
        with ThreadPoolExecutor(max(1, threads)) as ex:
            list(ex.map(one, range(B)))
    return host


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


def _cpu_state(k):
    """Instance k of the workload and its state after STATE_ITERS outer iterations of the oracle's Solver.solve."""
    from threadpoolctl import threadpool_limits

    from oracle import gradflow_oracle as orc
    from pygradflow_b200 import synth

    with threadpool_limits(limits=1):
        d = synth.qp_instance(k, N_VARS, N_CONS)
        prob = orc.DenseQP(d["H"], d["A"], d["g"], d["b"], d["lb"], d["ub"])
        res = orc.Solver(prob, orc.OracleParams(iteration_limit=STATE_ITERS)).solve(d["x0"], d["y0"])
    return (d["H"], d["A"], d["g"], d["b"], d["lb"], d["ub"], res.x, res.y, np.float64(res.lamb), np.float64(res.rho))


def cpu_steps(sample, cores):
    """Run the sample through a process pool (one instance per task, like the reference's runner:
    pygradflow/runners/runner.py:107-153).  Returns (results, wall seconds)."""
    import multiprocessing as mp

    if isinstance(sample, dict):
        S = sample["x"].shape[0]
        tasks = [tuple(np.ascontiguousarray(sample[k][i]) for k in ("H", "A", "g", "b", "lb", "ub", "x", "y", "lamb", "rho"))
                 for i in range(S)]
    else:
        tasks = sample
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
    """The reference's CPU path (oracle port, kind "port": the reference itself cannot travel to the GPU box) on all
    host cores: per step, the first CPU_SAMPLE instances of the SAME workload as the GPU arm (same generator seeds,
    same mid-solve state), one full Newton-KKT step each."""
    import multiprocessing as mp

    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    cores = host_cores()
    with mp.get_context("fork").Pool(cores) as pool:   # untimed set-up: data + state of the sampled instances
        tasks = pool.map(_cpu_state, range(CPU_SAMPLE), chunksize=1)
    for _ in range(min(args.warmup, 1)):
        cpu_steps(tasks[: 2 * cores], cores)
    walls = []
    for _ in range(args.steps):
        _, wall = cpu_steps(tasks, cores)
        walls.append(wall)
    total = sum(walls)
    value = CPU_SAMPLE * args.steps / total
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * total / args.steps,
        "higher_is_better": True, "scaling": args.scaling, "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": bench_config(args.gpus, args.scaling),
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": "port",
                         "sample": f"instances k=0..{CPU_SAMPLE - 1} of the workload per step (the GPU arm's first "
                                   f"{CPU_SAMPLE}: same seeds, same state after {STATE_ITERS} outer iterations), oracle "
                                   "port of the reference path (dense NumPy assembly + scipy splu = SuperLU, as "
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
    newest committed `ncu --set full` capture under profiles/ (a profiler run cannot happen inside the timed bench);
    (None, None) when no capture is at hand."""
    for name in ("r02g_ncu_ldlt_factor_dram.json", "r02_ncu_ldlt_factor_dram.json", "r01_ncu_ldlt_factor_dram.json"):
        try:
            with open(os.path.join(ROOT, "profiles", name)) as f:
                return float(json.load(f)["dram_bytes_per_factorisation"]), "profiles/" + name
        except Exception:
            continue
    return None, None


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

    from pygradflow_b200 import dist as gdist
    from pygradflow_b200 import kernels as K
    from pygradflow_b200.host_step import HostNewtonKKT
    from pygradflow_b200.newton import NewtonKKTStepper
    from pygradflow_b200.params import LinearSolverType, Params
    from pygradflow_b200.problem import BatchedQP
    from pygradflow_b200.solver import BatchedSolver

    n, m = N_VARS, N_CONS
    cfg = bench_config(world, args.scaling)
    B = cfg["batch_per_gpu"]
    k0 = rank * B
    t_gen = time.perf_counter()
    host = generate_host(k0, B, n, m, pinned=True, threads=max(1, host_cores() // min(world, 8)))
    gen_s = time.perf_counter() - t_gen
    f64 = dict(dtype=torch.float64, device=device)

    def upload(hostd):
        return {k: v.to(device, non_blocking=True) for k, v in hostd.items()}

    def make_prob(dd):
        q = object.__new__(BatchedQP)
        q.var_lb, q.var_ub = dd["lb"], dd["ub"]
        q.B, q.n, q.m, q.device = B, n, m, device
        q.H, q.A, q.g, q.b = dd["H"], dd["A"], dd["g"], dd["b"]
        return q

    data = upload(host)
    prob = make_prob(data)
    linear = LinearSolverType[args.linear]
    # State of the timed step: every instance's iterate after STATE_ITERS outer iterations of the batched solver from
    # the cfg3 start (x0 = 0, y0 = 0, default Params) -- a mid-solve point with its own lambda, rho and a developed
    # active set; the reference arm computes the same state with the oracle's Solver.
    bs = BatchedSolver(prob, Params(linear_solver_type=linear))
    bs.solve(None, None, max_outer=STATE_ITERS)
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

    def max_over_ranks(v):
        t = torch.tensor([v], dtype=torch.float64, device=device)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

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
    ms_max = max_over_ranks(ms)
    value = B * world * args.steps / (ms_max * 1e-3)
    phases = stepper.phase_ms()
    eng = stepper.engine
    info = out[4]
    Nvec = eng.Nvec.to(torch.float64)
    flops_ldlt = float((Nvec ** 3 / 3.0).sum().item())
    n_fallback = int((eng.fbkey != 0).sum().item()) if linear != LinearSolverType.LU else 0
    nfail = int((info != 0).sum().item())
    xn_gpu, yn_gpu, diff_gpu, fn_gpu = (o.clone() for o in out[:4])
    factor_ms = stepper.ldlt_ms() or phases["factor"]
    nIbar = float(eng.nI.to(torch.float64).mean().item())
    Nbar = float(Nvec.mean().item())
    nId = eng.nI.double()
    if eng.linear == LinearSolverType.LDLT:  # lower triangle only: half of H_II read, half of K written
        asm_bytes = float(((nId ** 2 / 2 + m * nId) * 8 + n + (Nvec ** 2) * 4).sum().item())
    else:
        asm_bytes = float(((nId ** 2 + m * nId) * 8 + n + (Nvec ** 2) * 8).sum().item())
    solve_bytes = float(((Nvec ** 2) * 8 + 16 * Nvec).sum().item())
    eng_linear = eng.linear

    # ---- end to end, per step, through host buffers (pinned).  Two shapes of the same public call:
    #   e2e           : the problem is registered once (its data stays in HBM, like the reference keeps its Problem object
    #                   across the steps of a solve); every step copies this step's inputs -- x, y, lambda, rho -- from
    #                   pinned host memory and reads xn, yn, diff, |F|, info back (ResidentNewtonKKT)
    #   e2e_streaming : nothing is resident: H, A, g, b, bounds cross PCIe again every step, chunked and double
    #                   buffered (HostNewtonKKT) -- for data that changes every step or exceeds HBM
    e2e, e2e_streaming = None, None
    hostx = {k: HostNewtonKKT.pinned_like(v) for k, v in dict(x=x, y=y, lamb=lamb, rho=rho).items()}
    outp = {"xn": torch.empty((B, n), dtype=torch.float64, pin_memory=True),
            "yn": torch.empty((B, m), dtype=torch.float64, pin_memory=True),
            "diff": torch.empty((B,), dtype=torch.float64, pin_memory=True),
            "fnorm": torch.empty((B,), dtype=torch.float64, pin_memory=True),
            "info": torch.empty((B,), dtype=torch.int32, pin_memory=True)}
    try:
        from pygradflow_b200.host_step import ResidentNewtonKKT

        stepper.events = None
        stepper.engine.ldlt_events = None
        rk = ResidentNewtonKKT(prob, linear, stepper=stepper)
        rk.step(hostx, outp)
        barrier()
        t0 = time.perf_counter()
        a0, a1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a0.record()
        for _ in range(args.steps):
            rk.step(hostx, outp)
        a1.record()
        barrier()
        wall = time.perf_counter() - t0
        ems = max_over_ranks(a0.elapsed_time(a1))
        h2d, d2h = rk.bytes_per_step()
        e2e = {"value": B * world * args.steps / (ems * 1e-3), "unit": UNIT,
               "h2d_bytes_per_step": h2d * world, "d2h_bytes_per_step": d2h * world, "steps": args.steps,
               "wall_s": wall, "max_abs_diff_vs_device_path": float((outp["xn"].to(device) - xn_gpu).abs().max().item()),
               "what": "problem registered once (H, A, g, b, bounds resident in HBM, uploaded outside the timed region, "
                       "as the reference arm keeps its Problem objects in host memory); per step: x, y, lambda, rho "
                       "pinned host -> device, the full Newton-KKT step, xn, yn, diff, |F(next)|, info device -> pinned "
                       "host, all inside the timed region (CUDA events on the launching stream, max over ranks)"}
        del rk
    except Exception as exc:  # pragma: no cover
        e2e = {"value": None, "unit": UNIT, "error": repr(exc)}
    try:
        hk = HostNewtonKKT(n, m, chunk=256, device=device, linear=linear)
        hin = dict(host)
        hin.update(hostx)
        esteps = max(1, min(args.steps, 3))
        hk.step(hin, outp)
        barrier()
        t0 = time.perf_counter()
        a0, a1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a0.record()
        for _ in range(esteps):
            hk.step(hin, outp)
        a1.record()
        barrier()
        wall = time.perf_counter() - t0
        ems = max_over_ranks(a0.elapsed_time(a1))
        h2d, d2h = hk.bytes_per_step(B)
        e2e_streaming = {"value": B * world * esteps / (ems * 1e-3), "unit": UNIT,
               "h2d_bytes_per_step": h2d * world, "d2h_bytes_per_step": d2h * world, "steps": esteps,
               "wall_s": wall, "chunk": 256, "host_gb_per_s": (h2d + d2h) * world * esteps / (ems * 1e-3) * 1e-9,
               "max_abs_diff_vs_device_path": float((outp["xn"].to(device) - xn_gpu).abs().max().item()),
               "limit": "PCIe / host DRAM: 2.25 MB per instance cross the bus every step; the box exposes one NUMA "
                        "node for all GPUs (nvidia-smi topo), so there is no NUMA-local staging to choose"}
        del hk, hin
    except Exception as exc:  # pragma: no cover
        e2e_streaming = {"value": None, "unit": UNIT, "error": repr(exc)}

    # ---- end to end, whole solves: upload the problem data once from pinned host memory, BatchedSolver.solve,
    # download (x, y, status, iterations) -- H and A are constant over the Newton steps of a solve
    del stepper, eng
    torch.cuda.empty_cache()
    e2e_solve, gather = None, None
    try:
        barrier()
        s0, s1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        t0 = time.perf_counter()
        s0.record()
        d2 = upload(host)
        solver = BatchedSolver(make_prob(d2), Params(linear_solver_type=linear))
        res = solver.solve(None, None)
        hx = res.x.to("cpu", non_blocking=True)
        hy = res.y.to("cpu", non_blocking=True)
        hs = res.status.to("cpu", non_blocking=True)
        hi = res.iterations.to("cpu", non_blocking=True)
        s1.record()
        barrier()
        wall = time.perf_counter() - t0
        sms = max_over_ranks(s0.elapsed_time(s1))
        wall = max_over_ranks(wall)
        up = sum(v.numel() * 8 for v in host.values())
        down = (hx.numel() + hy.numel()) * 8 + (hs.numel() + hi.numel()) * 4
        counts = torch.stack([(res.status == 1).sum(), res.iterations.sum(), res.iterations.max()]).to(torch.int64)
        steps_t = torch.tensor([res.newton_steps], dtype=torch.int64, device=device)
        if world > 1:
            dist.all_reduce(counts[:2], op=dist.ReduceOp.SUM)
            dist.all_reduce(counts[2:], op=dist.ReduceOp.MAX)
            dist.all_reduce(steps_t, op=dist.ReduceOp.SUM)
        e2e_solve = {"value": B * world / (wall), "unit": "solves/s", "wall_s": wall, "device_ms": sms,
                     "h2d_bytes": up * world, "d2h_bytes": down * world, "optimal": int(counts[0].item()),
                     "instances": B * world, "max_iterations": int(counts[2].item()),
                     "newton_kkt_steps": int(steps_t.item()),
                     "newton_kkt_steps_per_s": float(steps_t.item()) / wall,
                     "what": "upload H, A, g, b, bounds once (pinned host -> HBM), BatchedSolver.solve from x0=0, y0=0 "
                             "(default Params), download x, y, status, iterations; wall clock, max over ranks"}
        # ---- the single collective of a sharded solve: all-gather of the converged iterates / status
        if world > 1:
            local = dict(x=res.x, y=res.y, status=res.status, iterations=res.iterations, accepted_steps=res.accepted_steps)
            barrier()
            g0, g1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            g0.record()
            full = gdist.gather_result(local, B * world, (rank * B, (rank + 1) * B))
            g1.record()
            barrier()
            gather = {"gather_ms": max_over_ranks(g0.elapsed_time(g1)), "rows": int(full.x.shape[0]),
                      "bytes": int(full.x.numel() * 8 + full.y.numel() * 8 + 3 * 4 * full.status.numel()),
                      "optimal_after_gather": int((full.status == 1).sum().item()), "collective": "all_gather_into_tensor (NCCL)"}
        del solver, d2, res
    except Exception as exc:  # pragma: no cover
        e2e_solve = {"value": None, "unit": "solves/s", "error": repr(exc)}
    torch.cuda.empty_cache()

    cpu_baseline = None
    parity = None
    parity_solve = None
    if rank == 0 and world == 1 and not args.no_cpu:
        cores = host_cores()
        S = min(CPU_SAMPLE, B)
        sample = {k: host[k][:S].numpy() for k in ("H", "A", "g", "b", "lb", "ub")}
        sample["x"], sample["y"] = hostx["x"][:S].numpy(), hostx["y"][:S].numpy()
        sample["lamb"], sample["rho"] = hostx["lamb"][:S].numpy(), hostx["rho"][:S].numpy()
        res_c, wall = cpu_steps(sample, cores)
        cpu_value = S / wall
        xr = np.stack([r[0] for r in res_c])
        yr = np.stack([r[1] for r in res_c])
        dr = np.array([r[2] for r in res_c])
        rel = lambda a, b_: float(np.max(np.abs(a - b_)) / max(1.0, np.max(np.abs(b_))))
        parity = {"instances": S, "x_rel": rel(xn_gpu[:S].cpu().numpy(), xr), "y_rel": rel(yn_gpu[:S].cpu().numpy(), yr),
                  "diff_rel": rel(diff_gpu[:S].cpu().numpy(), dr)}
        cpu_baseline = {"value": cpu_value, "unit": UNIT, "cores": cores, "kind": "port",
                        "sample": f"first {S} of the {B} instances (identical inputs and state: the GPU arm's own "
                                  "buffers), one full Newton-KKT step each through the oracle port (dense NumPy "
                                  "assembly + scipy splu = SuperLU as lu_solver.py:14), one process per core",
                        "sum_cpu_seconds": float(sum(r[4] for r in res_c)), "wall_s": wall}
        if not args.no_parity_solves:
            # whole solves of the first instances against the oracle's Solver.solve: decisions must be identical up to
            # the rounding-noise horizon (tools/parity_sweep.py; 128-instance tables under profiles/r02_parity_sweep.json)
            try:
                sys.path.insert(0, os.path.join(ROOT, "tools"))
                import parity_sweep

                sm = parity_sweep.sweep(3, args.parity_solves, workers=cores)["summary"]
                parity_solve = {k: sm[k] for k in ("instances", "status_equal", "identical", "never_hit_horizon",
                                                   "never_hit_horizon_identical", "hit_horizon", "diverged_before_horizon",
                                                   "abs_iter_diff_hist", "max_x_rel_identical", "post_horizon_share")}
            except Exception as exc:  # pragma: no cover
                parity_solve = {"error": repr(exc)}

    if rank == 0:
        # the roofline is that of the factorisation kernels: events around gf_ldlt_factor alone (the `factor` phase
        # of the step additionally holds the bookkeeping of the pivoted-LU fallback list)
        achieved = flops_ldlt / (factor_ms * 1e-3) * 1e-12 if linear != LinearSolverType.LU else \
            2.0 * flops_ldlt / (factor_ms * 1e-3) * 1e-12
        hbm_peak = None
        try:
            hbm_peak = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"]
        except Exception:
            hbm_peak = 6650.0
        traffic, traffic_src = ncu_traffic()
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms_max / args.steps, "higher_is_better": True,
            "scaling": args.scaling, "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": cfg,
            "workload_stats": {"running_instances": running, "linear_solver": eng_linear.name, "mean_inactive": nIbar,
                               "mean_kkt_order": Nbar, "lu_fallback_instances": n_fallback, "failed_instances": nfail,
                               "data_generation_s": gen_s},
            "roofline": {"bound": "tensor", "kernel": "batched LDL' factorisation = gf_ldlt_factor: ldlt_diag0_kernel + one "
                                                      "ldlt_column_kernel (DMMA) per 64-wide block column, issued as "
                                                      "two interleaved half-batches",
                         "achieved": achieved, "peak": fp64_peak, "unit": "TFLOP/s",
                         "frac": achieved / fp64_peak if fp64_peak else None, "traffic": traffic,
                         "traffic_source": traffic_src,
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
            "e2e_streaming": e2e_streaming,
            "e2e_solve": e2e_solve,
            "gather": gather,
            "gpu_launches": launches,
            "cpu_baseline": cpu_baseline,
            "parity_vs_cpu_sample": parity,
            "parity_solves_vs_cpu": parity_solve,
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
    ap.add_argument("--scaling", default="weak", choices=["weak", "strong"],
                    help="weak: 4096 instances per GPU; strong: a global batch of 4096 sharded over the GPUs")
    ap.add_argument("--no-parity-solves", action="store_true", help="skip the whole-solve parity sample")
    ap.add_argument("--parity-solves", type=int, default=32, help="instances of the whole-solve parity sample")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference_arm(args)
    else:
        run_gpu_arm(args)


if __name__ == "__main__":
    main()
