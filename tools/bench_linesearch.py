"""HBM throughput of the globalized-Newton kernels at the cfg3 shape (n=512, m=256, B=4096): merit gradient
(F'^T F matrix-free: H once, J twice), trial point, fused residual-norm + Armijo test."""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from pygradflow_b200 import kernels as K

B, n, m = 4096, 512, 256
f64 = dict(dtype=torch.float64, device="cuda")
g = torch.Generator(device="cuda"); g.manual_seed(0)
R = lambda *s: torch.randn(*s, generator=g, **f64)
H, J, F = R(B, n, n), R(B, m, n), R(B, n + m)
active = (torch.rand(B, n, generator=g, device="cuda") < 0.1).to(torch.uint8)
dt, rho = torch.full((B,), 0.5, **f64), torch.full((B,), 1e-2, **f64)
dx, dy, x, y, x0, y0, dL, cons = R(B, n), R(B, m), R(B, n), R(B, m), R(B, n), R(B, m), R(B, n), R(B, m)
lb, ub = -torch.ones(B, n, **f64), torch.ones(B, n, **f64)
res, inner, alpha, nres = (torch.zeros(B, **f64) for _ in range(4))
xt, yt = torch.zeros(B, n, **f64), torch.zeros(B, m, **f64)
trials = torch.zeros(B, dtype=torch.int32, device="cuda"); state = torch.zeros_like(trials)
w = K.WorkList.all(B)


def timeit(fn, reps=5):
    fn(); torch.cuda.synchronize()
    best = 1e9
    for _ in range(reps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record(); torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1))
    return best


out = {}
ms = timeit(lambda: K.merit_grad(H, J, F, active, dt, rho, dx, dy, res, inner, w))
by = B * 8.0 * (n * n + 2 * m * n + 3 * (n + m))
out["merit_grad"] = dict(ms=ms, algorithmic_bytes=by, gbs=by / ms * 1e-6)
alpha.fill_(1.0)
ms = timeit(lambda: K.ls_trial(x, y, dx, dy, alpha, xt, yt, w))
by = B * 8.0 * 3 * (n + m)
out["ls_trial"] = dict(ms=ms, algorithmic_bytes=by, gbs=by / ms * 1e-6)


def arm():
    state.zero_(); trials.zero_(); alpha.fill_(1.0)
    K.armijo_residual(xt, yt, x0, y0, dL, cons, lb, ub, dt, res, inner, 1e-8, 30, alpha, trials, state, nres, w)


ms = timeit(arm)
by = B * 8.0 * (5 * n + 3 * m)
out["armijo_residual (+3 tiny fills)"] = dict(ms=ms, algorithmic_bytes=by, gbs=by / ms * 1e-6)
out["hbm_peak_gbs"] = 6554.9
print(json.dumps(out, indent=1))
