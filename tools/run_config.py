"""Full-size solves of the BASELINE.json configurations through the batched driver, with a parity sample.

    python tools/run_config.py --cfg 2 [--B 4096] [--linear Auto|LU|LDLT] [--check 4] [--out profiles/x.json]

cfg2: chained Rosenbrock n=64, bounds only;  cfg3: random dense convex QPs n=512, m=256.
Reports solves/s, Newton-KKT steps/s, status counts, iteration statistics and, for the first `--check` instances,
the comparison with the CPU oracle (status, iteration count, final iterate).
"""
import argparse
import json
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch

from pygradflow_b200 import synth
from pygradflow_b200.params import LinearSolverType, NewtonType, Params
from pygradflow_b200.problem import BatchedOCP, BatchedQP, BatchedRosenbrock
from pygradflow_b200.solver import BatchedSolver


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--cfg", type=int, default=2)
    ap.add_argument("--B", type=int, default=4096)
    ap.add_argument("--n", type=int, default=None)
    ap.add_argument("--m", type=int, default=None)
    ap.add_argument("--stages", type=int, default=128)
    ap.add_argument("--linear", default="Auto")
    ap.add_argument("--newton", default="Simplified")
    ap.add_argument("--step-solver", default="Symmetric", help="Symmetric | Asymmetric | Extended (full-order LU)")
    ap.add_argument("--check", type=int, default=4)
    ap.add_argument("--no-graph", action="store_true", help="run every outer iteration eagerly (no CUDA-graph replay)")
    ap.add_argument("--out", default=None)
    args = ap.parse_args()
    B = args.B
    t0 = time.perf_counter()
    if args.cfg == 2:
        n, m = args.n or 64, 0
        d = synth.rosenbrock_batch(range(B), n)
        prob = BatchedRosenbrock(d["a"], d["b"], d["lb"], d["ub"])
        x0, y0 = d["x0"], None
    elif args.cfg == 4:
        d = synth.ocp_batch(range(B), stages=args.stages)
        prob = BatchedOCP(d["A"], d["B"], d["Q"], d["R"], d["xinit"], d["umax"], d["h"])
        n, m = prob.n, prob.m
        x0, y0 = d["x0"], d["y0"]
    else:
        n, m = args.n or 512, args.m if args.m is not None else 256
        d = synth.qp_batch(range(B), n, m)
        prob = BatchedQP(d["H"], d["A"], d["g"], d["b"], d["lb"], d["ub"])
        x0, y0 = d["x0"], d["y0"]
    gen_s = time.perf_counter() - t0
    params = Params(linear_solver_type=LinearSolverType[args.linear], newton_type=NewtonType[args.newton],
                    step_solver_type=args.step_solver)
    solver = BatchedSolver(prob, params, use_graph=not args.no_graph)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    res = solver.solve(x0, y0)
    torch.cuda.synchronize()
    wall = time.perf_counter() - t0
    st = res.status.cpu().numpy()
    it = res.iterations.cpu().numpy()
    out = dict(cfg=args.cfg, B=B, n=n, m=m, linear=solver.engine.linear.name, newton=args.newton,
               step_solver=args.step_solver, wall_s=wall,
               solves_per_s=B / wall, newton_steps=res.newton_steps, newton_steps_per_s=res.newton_steps / wall,
               outer_iterations=res.outer_iterations, ms_per_outer=1e3 * wall / max(1, res.outer_iterations),
               status_counts={int(k): int(v) for k, v in zip(*np.unique(st, return_counts=True))},
               iterations=dict(min=int(it.min()), median=float(np.median(it)), max=int(it.max())),
               max_total_res=float(res.total_res.max().item()), data_gen_s=gen_s, device=torch.cuda.get_device_name(0))
    if args.check > 0:
        from oracle import gradflow_oracle as orc

        chk = []
        for b in range(min(args.check, B)):
            if args.cfg == 4:
                p = orc.OCP(d["A"][b], d["B"][b], d["Q"][b], d["R"][b], d["xinit"][b], d["umax"], d["h"])
                t1 = time.perf_counter()
                ref = orc.Solver(p, orc.OracleParams()).solve(d["x0"][b], d["y0"][b])
            elif args.cfg == 2:
                p = orc.ChainedRosenbrock(d["a"][b], d["b"][b], d["lb"][b], d["ub"][b])
                t1 = time.perf_counter()
                ref = orc.Solver(p, orc.OracleParams()).solve(d["x0"][b], np.zeros(0))
            else:
                p = orc.DenseQP(d["H"][b], d["A"][b], d["g"][b], d["b"][b], d["lb"][b], d["ub"][b])
                t1 = time.perf_counter()
                ref = orc.Solver(p, orc.OracleParams(step_solver_type=args.step_solver.lower())).solve(d["x0"][b], d["y0"][b])
            cpu_s = time.perf_counter() - t1
            xg = res.x[b].cpu().numpy()
            chk.append(dict(instance=b, status_gpu=int(st[b]), status_cpu=int(ref.status), iters_gpu=int(it[b]),
                            iters_cpu=int(ref.iterations), x_rel=float(np.max(np.abs(xg - ref.x)) / max(1.0, np.max(np.abs(ref.x)))),
                            cpu_seconds=cpu_s))
        out["oracle_check"] = chk
    print(json.dumps(out))
    if args.out:
        with open(args.out, "w") as f:
            json.dump(out, f, indent=1)


if __name__ == "__main__":
    main()
