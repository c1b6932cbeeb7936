"""Synthetic problem data for the configurations named in BASELINE.json.

Pure data generation (NumPy, seeded per instance) -- no algorithm lives here.  The same
arrays feed the reference (golden fixtures), the CPU oracle and the B200 path.
"""

from __future__ import annotations

import numpy as np


def qp_instance(k: int, n: int, m: int):
    """cfg3 generator: random dense convex QP with equality + bound constraints.

    rng = default_rng(2000 + k); H = sym(M M'/n) + 0.1 I; A ~ N(0,1); b = -A x_f (feasible);
    g = 0.3 N(0,1); bounds [-1, 1]; x0 = 0, y0 = 0.
    """
    rng = np.random.default_rng(2000 + k)
    M = rng.standard_normal((n, n))
    G = (M @ M.T) / n
    H = 0.5 * (G + G.T) + 0.1 * np.eye(n)
    A = rng.standard_normal((m, n))
    xf = rng.uniform(-0.5, 0.5, n)
    b = -(A @ xf)
    g = 0.3 * rng.standard_normal(n)
    return dict(
        H=H, A=A, g=g, b=b, lb=np.full(n, -1.0), ub=np.full(n, 1.0), x0=np.zeros(n), y0=np.zeros(m)
    )


def qp_batch(ks, n: int, m: int):
    """Stack qp_instance over ``ks`` -> dict of [B, ...] arrays."""
    items = [qp_instance(int(k), n, m) for k in ks]
    return {key: np.stack([it[key] for it in items]) for key in items[0]}


def rosenbrock_instance(k: int, n: int):
    """cfg2 generator: chained Rosenbrock with perturbed coefficients, bounds only.

    rng = default_rng(1000 + k); a_i = 1 + 0.1 U(-1,1); b_i = 100 (1 + 0.1 U(-1,1));
    lb = -1.5; ub_i = 0.9 (i even) / 2.0 (i odd); x0_i = -1.2 / 1.0 + 0.05 U(-1,1), clipped.
    """
    rng = np.random.default_rng(1000 + k)
    a = 1.0 + 0.1 * rng.uniform(-1, 1, n - 1)
    b = 100.0 * (1.0 + 0.1 * rng.uniform(-1, 1, n - 1))
    lb = np.full(n, -1.5)
    ub = np.where(np.arange(n) % 2 == 0, 0.9, 2.0)
    x0 = np.where(np.arange(n) % 2 == 0, -1.2, 1.0) + 0.05 * rng.uniform(-1, 1, n)
    x0 = np.clip(x0, lb, ub)
    return dict(a=a, b=b, lb=lb, ub=ub, x0=x0, y0=np.zeros(0))


def rosenbrock_batch(ks, n: int):
    items = [rosenbrock_instance(int(k), n) for k in ks]
    return {key: np.stack([it[key] for it in items]) for key in items[0]}


def kkt_instance(N: int, k: int = 0, lamb: float = 1.0, rho: float = 1e-2):
    """cfg5 generator: a quasi-definite KKT matrix of order N = n_I + m with m = N // 3.

    rng = default_rng(4000 + N + 7919 k).  K = [[H + lamb I, A'], [A, -lamb/(1+lamb rho) I]].
    """
    rng = np.random.default_rng(4000 + N + 7919 * k)
    m = N // 3
    nI = N - m
    M = rng.standard_normal((nI, nI))
    G = (M @ M.T) / max(nI, 1)
    H = 0.5 * (G + G.T) + 0.1 * np.eye(nI)
    A = rng.standard_normal((m, nI))
    K = np.zeros((N, N))
    K[:nI, :nI] = H + lamb * np.eye(nI)
    K[:nI, nI:] = A.T
    K[nI:, :nI] = A
    K[nI:, nI:] = (-lamb / (1.0 + lamb * rho)) * np.eye(m)
    rhs = rng.standard_normal(N)
    return K, rhs, m


def ocp_instance(k: int, stages: int = 128, nx: int = 8, nu: int = 8, h: float = 0.05):
    """cfg4 generator: discretised nonlinear optimal-control problem.

    Variables z = (x_1, u_0, x_2, u_1, ..., x_S, u_{S-1}) stage-interleaved, n = S (nx + nu).
    Dynamics x_{j+1} = x_j + h (A_j x_j + B_j u_j + 0.1 sin(x_j)), x_0 fixed  => m = S nx.
    Cost sum 1/2 (x'Qx + u'Ru), Q, R diagonal positive.  Bounds on u only.
    rng = default_rng(3000 + k).
    """
    rng = np.random.default_rng(3000 + k)
    Aj = 0.5 * rng.standard_normal((stages, nx, nx)) / np.sqrt(nx)
    Bj = rng.standard_normal((stages, nx, nu)) / np.sqrt(nu)
    Q = rng.uniform(0.5, 2.0, (stages, nx))
    R = rng.uniform(0.1, 0.5, (stages, nu))
    xinit = rng.uniform(-1.0, 1.0, nx)
    umax = 0.4
    n = stages * (nx + nu)
    return dict(A=Aj, B=Bj, Q=Q, R=R, xinit=xinit, umax=umax, h=h, stages=stages, nx=nx, nu=nu,
                x0=np.zeros(n), y0=np.zeros(stages * nx))


def ocp_batch(ks, stages: int = 128, nx: int = 8, nu: int = 8, h: float = 0.05):
    """Stack ocp_instance over ``ks`` -> dict of [B, ...] arrays (scalars umax, h, stages, nx, nu unchanged)."""
    items = [ocp_instance(int(k), stages, nx, nu, h) for k in ks]
    out = {key: np.stack([it[key] for it in items]) for key in ("A", "B", "Q", "R", "xinit", "x0", "y0")}
    out.update({key: items[0][key] for key in ("umax", "h", "stages", "nx", "nu")})
    return out


def general_qp_instance(k: int, n: int, m: int):
    """QP with general constraint bounds (the input form of the reference's slack transform): qp_instance's data,
    the first half of the rows equalities A x + b = e_i with small non-zero right-hand sides, the second half
    ranges -r_i <= A x + b <= r_i (one of them one-sided)."""
    d = qp_instance(k, n, m)
    rng = np.random.default_rng(5000 + k)
    me = m // 2
    e = 0.05 * rng.standard_normal(me)
    r = rng.uniform(0.05, 0.3, m - me)
    cl = np.concatenate([e, -r])
    cu = np.concatenate([e, r])
    if m - me > 0:
        cu[-1] = np.inf
    d.update(cons_lb=cl, cons_ub=cu)
    return d


def general_qp_batch(ks, n: int, m: int):
    items = [general_qp_instance(int(k), n, m) for k in ks]
    return {key: np.stack([it[key] for it in items]) for key in items[0]}
