"""CUDA-event timings of the stage-structured (cfg4) kernels: compact Jacobian / Hessian evaluators, J'v, Schur-complement
factorisation by block cyclic reduction, solve.  python tools/bench_stage.py [--B 128] [--stages 128] [--out x.json]"""
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
from pygradflow_b200.newton import NewtonKKTStepper
from pygradflow_b200.params import LinearSolverType
from pygradflow_b200.problem import BatchedOCP


def timeit(fn, reps=50):
    for _ in range(5):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps * 1e3  # us


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--B", type=int, default=128)
    ap.add_argument("--stages", type=int, default=128)
    ap.add_argument("--out", default=None)
    a = ap.parse_args()
    B, S = a.B, a.stages
    d = synth.ocp_batch(range(B), stages=S)
    prob = BatchedOCP(d["A"], d["B"], d["Q"], d["R"], d["xinit"], d["umax"], d["h"])
    st = NewtonKKTStepper(prob, LinearSolverType.Auto)
    eng = st.engine
    n, m = prob.n, prob.m
    f64 = dict(dtype=torch.float64, device="cuda")
    rng = np.random.default_rng(0)
    z = torch.as_tensor(np.clip(0.5 * rng.standard_normal((B, n)), -0.4, 0.4), **f64)
    y = torch.as_tensor(0.3 * rng.standard_normal((B, m)), **f64)
    lamb, rho = torch.full((B,), 2.0, **f64), torch.full((B,), 1e-3, **f64)
    st.step(z, y, lamb, rho)
    w = WorkList.all(B)
    J, H = st.Jbuf[0], st.Hbuf
    S_, nx, nu = eng.stage
    out = dict(B=B, S=S, n=n, m=m, device=torch.cuda.get_device_name(0))
    out["jac_us"] = timeit(lambda: prob.jac(z, J, w))
    out["hess_us"] = timeit(lambda: prob.lag_hess(z, y, H, w))
    out["eval_us"] = timeit(lambda: prob.eval(z, st.grad, st.cons, st.obj, w))
    out["aug_lag_grad_us"] = timeit(lambda: prob.aug_lag_grad(J, st.grad, st.cons, y, rho, st.dL, None, None, w))
    out["factor_us"] = timeit(lambda: eng.factor_assembled(H, J, st.dt, rho, w))
    out["solve_us"] = timeit(lambda: K.stage_kkt_solve(S_, nx, nu, J, H, eng.active, st.F, st.dt, rho, eng.Tinv, eng.Pf,
                                                       eng.Qf, eng.rhs, w))
    out["step_us"] = timeit(lambda: st.step(z, y, lamb, rho), reps=20)
    # algorithmic bytes: compact J (8 (nx + w) per row) + H diagonal + factors
    jb = B * m * (2 * nx + nu) * 8
    fb = 3 * B * S * nx * nx * 8
    out["factor_GBps"] = (jb + B * n * 9 + fb) / (out["factor_us"] * 1e-6) * 1e-9
    out["solve_GBps"] = (jb + B * n * 9 + fb + B * (n + m) * 16) / (out["solve_us"] * 1e-6) * 1e-9
    print(json.dumps(out))
    if a.out:
        json.dump(out, open(a.out, "w"), indent=1)


if __name__ == "__main__":
    main()
