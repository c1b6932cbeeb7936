"""Kernel-level parity of the CUDA path (through the C ABI) against the CPU oracle and the golden
fixtures produced by the real reference.  Run on the B200 box: pytest -m gpu.

Tolerances: copies / gathers / thresholds bit-exact; floating-point results 1e-10 relative
(north star), most are checked far tighter.
"""

import numpy as np
import pytest
import scipy.linalg

torch = pytest.importorskip("torch")
pytestmark = pytest.mark.gpu

from helpers import rel_err  # noqa: E402
from oracle import gradflow_oracle as orc  # noqa: E402
from pygradflow_b200 import synth  # noqa: E402


def _need_gpu():
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")


@pytest.fixture(scope="module")
def K():
    _need_gpu()
    from pygradflow_b200 import kernels

    return kernels


def dev(a, dtype=torch.float64):
    return torch.as_tensor(np.ascontiguousarray(a), dtype=dtype).cuda().contiguous()


def allw(K, B):
    return K.WorkList.all(B)


# ------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("n,m", [(16, 8), (64, 32), (50, 0), (130, 67), (512, 256)])
def test_qp_eval_and_aug_lag_grad(K, n, m):
    B = 5
    d = synth.qp_batch(range(B), n, m)
    rng = np.random.default_rng(0)
    x = rng.uniform(-1, 1, (B, n))
    y = rng.standard_normal((B, m))
    rho = rng.uniform(0.01, 2.0, B)
    grad = torch.zeros(B, n, dtype=torch.float64, device="cuda")
    cons = torch.zeros(B, m, dtype=torch.float64, device="cuda")
    obj = torch.zeros(B, dtype=torch.float64, device="cuda")
    H, A, g, b = dev(d["H"]), (dev(d["A"]) if m else None), dev(d["g"]), (dev(d["b"]) if m else None)
    K.qp_eval(H, A, g, b, dev(x), grad, cons if m else None, obj, allw(K, B))
    dL = torch.zeros_like(grad)
    jty = torch.zeros_like(grad)
    jtc = torch.zeros_like(grad)
    K.aug_lag_grad(A, grad, cons if m else None, dev(y) if m else None, dev(rho), dL, jty if m else None,
                   jtc if m else None, allw(K, B))
    for k in range(B):
        p = orc.DenseQP(d["H"][k], d["A"][k], d["g"][k], d["b"][k], d["lb"][k], d["ub"][k])
        it = orc.Iterate(p, orc.OracleParams(), x[k], y[k])
        assert rel_err(grad[k].cpu().numpy(), it.obj_grad) <= 1e-13
        assert rel_err(cons[k].cpu().numpy(), it.cons) <= 1e-13
        assert abs(obj[k].item() - it.obj) <= 1e-12 * max(1.0, abs(it.obj))
        assert rel_err(dL[k].cpu().numpy(), it.aug_lag_deriv_x(rho[k])) <= 1e-12
        if m:
            assert rel_err(jty[k].cpu().numpy(), it.cons_jac.T @ it.y) <= 1e-12
            assert rel_err(jtc[k].cpu().numpy(), it.cons_jac.T @ it.cons) <= 1e-12


@pytest.mark.parametrize("n", [2, 8, 64, 100])
def test_rosenbrock_eval_bit_exact(K, n):
    B = 4
    d = synth.rosenbrock_batch(range(B), n)
    rng = np.random.default_rng(1)
    x = rng.uniform(-1.5, 2.0, (B, n))
    grad = torch.zeros(B, n, dtype=torch.float64, device="cuda")
    obj = torch.zeros(B, dtype=torch.float64, device="cuda")
    Hd = torch.zeros(B, n, n, dtype=torch.float64, device="cuda")
    K.rosen_eval(dev(d["a"]), dev(d["b"]), dev(x), grad, obj, allw(K, B))
    K.rosen_hess(dev(d["a"]), dev(d["b"]), dev(x), Hd, allw(K, B))
    for k in range(B):
        p = orc.ChainedRosenbrock(d["a"][k], d["b"][k], d["lb"][k], d["ub"][k])
        # same operation order as the NumPy expressions => identical bits
        assert np.array_equal(grad[k].cpu().numpy(), p.obj_grad(x[k]))
        assert np.array_equal(Hd[k].cpu().numpy(), p.lag_hess(x[k], None))
        assert abs(obj[k].item() - p.obj(x[k])) <= 1e-12 * max(1.0, abs(p.obj(x[k])))


# ------------------------------------------------------------------------------------------------
def _random_state(n, m, k, seed):
    d = synth.qp_instance(k, n, m)
    rng = np.random.default_rng(seed)
    x0 = np.clip(rng.uniform(-1.3, 1.3, n), d["lb"], d["ub"])
    y0 = 0.2 * rng.standard_normal(m)
    x = np.clip(x0 + 0.1 * rng.standard_normal(n), d["lb"], d["ub"])
    y = y0 + 0.1 * rng.standard_normal(m)
    return d, x0, y0, x, y


@pytest.mark.parametrize("n,m", [(16, 8), (64, 32), (48, 0), (200, 70)])
@pytest.mark.parametrize("dt,rho", [(1.0, 1e-8), (0.05, 1.0), (20.0, 1e-2)])
def test_residual_active_assemble_rhs_finish(K, n, m, dt, rho):
    """a4-a12, a15 against the oracle on identical inputs: active set and K bit-exact."""
    from pygradflow_b200.engine import KKTEngine
    from pygradflow_b200.params import LinearSolverType

    B = 3
    f64 = dict(dtype=torch.float64, device="cuda")
    states = [_random_state(n, m, k, 10 + k) for k in range(B)]
    stack = lambda i: np.stack([s[i] for s in states])
    data = {key: np.stack([s[0][key] for s in states]) for key in states[0][0]}
    x0, y0, x, y = stack(1), stack(2), stack(3), stack(4)
    H, A = dev(data["H"]), (dev(data["A"]) if m else None)
    lb, ub = dev(data["lb"]), dev(data["ub"])
    dtv, rhov = torch.full((B,), dt, **f64), torch.full((B,), rho, **f64)
    grad, cons, obj = torch.zeros(B, n, **f64), torch.zeros(B, m, **f64), torch.zeros(B, **f64)
    K.qp_eval(H, A, dev(data["g"]), dev(data["b"]) if m else None, dev(x), grad, cons if m else None, obj, allw(K, B))
    dL = torch.zeros(B, n, **f64)
    K.aug_lag_grad(A, grad, cons if m else None, dev(y) if m else None, rhov, dL, None, None, allw(K, B))
    eng = KKTEngine(B, n, m, "cuda", LinearSolverType.LU)
    F = torch.zeros(B, n + m, **f64)
    nrm = torch.zeros(B, **f64)
    yd, y0d = (dev(y), dev(y0)) if m else (None, None)
    K.residual(dev(x), yd, dev(x0), y0d, dL, cons if m else None, lb, ub, dtv, True, 0, eng.active, F, nrm, allw(K, B))
    Fu = torch.zeros(B, n + m, **f64)
    nrmu = torch.zeros(B, **f64)
    act_u = torch.zeros(B, n, dtype=torch.uint8, device="cuda")
    K.residual(dev(x), yd, dev(x0), y0d, dL, cons if m else None, lb, ub, dtv, False, 0, act_u, Fu, nrmu, allw(K, B))
    eng.update_active_set(allw(K, B))
    K.kkt_assemble(H, A, eng.perm, eng.nI, dtv, rhov, eng.K, 1, False, allw(K, B))
    Kfull = eng.K.cpu().numpy().copy()
    K.kkt_rhs(H, A, eng.perm, eng.nI, F, dtv, rhov, eng.rhs, allw(K, B))
    rhs = eng.rhs.cpu().numpy().copy()
    for k in range(B):
        d = states[k][0]
        prob = orc.DenseQP(d["H"], d["A"], d["g"], d["b"], d["lb"], d["ub"])
        prm = orc.OracleParams()
        it0 = orc.Iterate(prob, prm, x0[k], y0[k])
        it = orc.Iterate(prob, prm, x[k], y[k])
        sf = orc.ScaledImplicitFunc(prob, it0, dt)
        uf = orc.ImplicitFunc(prob, it0, dt)
        act = sf.compute_active_set(it, rho)
        assert np.array_equal(eng.active[k].cpu().numpy().astype(bool), act)
        assert np.array_equal(act_u[k].cpu().numpy().astype(bool), uf.compute_active_set(it, rho))
        Fo = sf.value_at(it, rho)
        assert rel_err(F[k].cpu().numpy(), Fo) <= 1e-12
        assert abs(nrm[k].item() - np.linalg.norm(Fo)) <= 1e-12 * max(1.0, np.linalg.norm(Fo))
        Fuo = uf.value_at(it, rho)
        assert rel_err(Fu[k].cpu().numpy(), Fuo) <= 1e-12
        assert abs(nrmu[k].item() - np.linalg.norm(Fuo)) <= 1e-12 * max(1.0, np.linalg.norm(Fuo))
        nI = int((~act).sum())
        assert int(eng.nI[k].item()) == nI
        perm = eng.perm[k].cpu().numpy()
        assert np.array_equal(perm[:nI], np.where(~act)[0]) and np.array_equal(perm[nI:], np.where(act)[0])
        lamb = 1.0 / dt
        fact = 1.0 / (1.0 + lamb * rho)
        Fk = F[k].cpu().numpy()
        b0, b1, b2 = dt * Fk[:n][act], Fk[:n][~act], Fk[n:]
        Ko, ro = orc.kkt_system(d["H"], d["A"], act, lamb, rho, b0, b1, fact * b2)
        N = nI + m
        assert np.array_equal(Kfull[k, :N, :N], Ko)  # gathers + one add on the diagonal: bit-exact
        assert rel_err(rhs[k, :N], ro) <= 1e-12
    # lower-only + identity padding used by the LDL' path
    eng2 = KKTEngine(B, n, m, "cuda", LinearSolverType.LDLT)
    eng2.active.copy_(eng.active)
    eng2.update_active_set(allw(K, B))
    eng2.K.fill_(float("nan"))
    K.kkt_assemble(H, A, eng2.perm, eng2.nI, dtv, rhov, eng2.K, 64, True, allw(K, B))
    Klow = eng2.K.cpu().numpy()
    for k in range(B):
        N = int(eng.nI[k].item()) + m
        Np = ((N + 63) // 64) * 64
        ref = np.eye(Np)
        ref[:N, :N] = Kfull[k, :N, :N]
        assert np.array_equal(np.tril(Klow[k, :Np, :Np]), np.tril(ref))


# ------------------------------------------------------------------------------------------------
def _lu_solve_gpu(K, mats, rhss, trans=False):
    B = len(mats)
    ld = max(m.shape[0] for m in mats)
    Kd = np.zeros((B, ld, ld))
    rd = np.zeros((B, ld))
    Nv = np.array([m.shape[0] for m in mats], dtype=np.int32)
    for k, (m_, r_) in enumerate(zip(mats, rhss)):
        N = m_.shape[0]
        Kd[k, :N, :N] = m_
        rd[k, :N] = r_
    Kt, rt, Nt = dev(Kd), dev(rd), dev(Nv, torch.int32)
    piv = torch.zeros(B, ld, dtype=torch.int32, device="cuda")
    info = torch.zeros(B, dtype=torch.int32, device="cuda")
    K.lu_factor(Kt, ld, Nt, piv, info, allw(K, B))
    K.lu_solve(Kt, ld, Nt, piv, rt, trans, allw(K, B))
    return rt.cpu().numpy(), info.cpu().numpy(), piv.cpu().numpy(), Kt.cpu().numpy()


@pytest.mark.parametrize("name", ["indef", "posdef", "negdef"])
def test_lu_reference_fixtures(K, golden, name):
    """tests/pygradflow/test_linear_solver.py:94-136 matrices, solution of the reference's LUSolver."""
    g = golden("linear_solver")
    mat, rhs = g[f"{name}/mat"], g["rhs"]
    sol, info, _, _ = _lu_solve_gpu(K, [mat], [rhs])
    assert info[0] == 0
    assert np.allclose(mat @ sol[0, :5] - rhs, 0.0)
    assert rel_err(sol[0, :5], g[f"{name}/sol"]) <= 1e-10
    solt, _, _, _ = _lu_solve_gpu(K, [mat], [rhs], trans=True)
    assert rel_err(solt[0, :5], g[f"{name}/sol_trans"]) <= 1e-10


@pytest.mark.parametrize("N", [12, 48, 96, 200])
def test_lu_golden_kkt(K, golden, N):
    g = golden("linear_solver")
    Km, rhs, _ = synth.kkt_instance(N)
    sol, info, _, _ = _lu_solve_gpu(K, [Km], [rhs])
    assert info[0] == 0 and rel_err(sol[0, :N], g[f"kkt{N}/sol"]) <= 1e-10
    solt, _, _, _ = _lu_solve_gpu(K, [Km], [rhs], trans=True)
    assert rel_err(solt[0, :N], g[f"kkt{N}/sol_trans"]) <= 1e-10


@pytest.mark.parametrize("sizes", [[1, 2, 3, 5, 17, 31, 32], [32] * 19 + [7, 1], [1, 2, 3, 5, 17, 31, 32, 33],
                                   [64, 63, 40, 33, 7], [64, 63, 65, 100, 112], [113, 150, 257], [768, 700, 333], [1000, 530, 513, 40],
                                   [1100, 1030], [2100, 90]])
def test_lu_ragged_batch_vs_lapack(K, sizes):
    """Ragged orders in one batch (warp-per-matrix register kernel, shared-memory kernel, blocked panel kernel +
    DMMA trailing update); unsymmetric matrices,
    pivot sequence identical to LAPACK getrf on the transposed view, K x = r and K' x = r."""
    rng = np.random.default_rng(5)
    mats = [rng.standard_normal((N, N)) + 0.1 * np.eye(N) for N in sizes]
    rhss = [rng.standard_normal(N) for N in sizes]
    sol, info, piv, _ = _lu_solve_gpu(K, mats, rhss)
    solt, _, _, _ = _lu_solve_gpu(K, mats, rhss, trans=True)
    for k, N in enumerate(sizes):
        assert info[k] == 0
        ref = np.linalg.solve(mats[k], rhss[k])
        scale = np.linalg.cond(mats[k])
        assert rel_err(sol[k, :N], ref) <= 1e-14 * scale * 10
        assert rel_err(solt[k, :N], np.linalg.solve(mats[k].T, rhss[k])) <= 1e-14 * scale * 10
        _, lap_piv = scipy.linalg.lu_factor(mats[k].T)  # the kernel factors the column-major view = K'
        assert np.array_equal(piv[k, :N], lap_piv)


def test_lu_singular_and_nonfinite_info(K):
    sing = np.array([[1.0, 2.0, 3.0], [0.0, 0.0, 0.0], [1.0, 0.0, 1.0]])  # exactly singular (lu_solver.py:15-17)
    zero = np.zeros((4, 4))
    nanm = np.eye(3)
    nanm[1, 1] = np.nan
    _, info, _, _ = _lu_solve_gpu(K, [sing, zero, nanm, np.eye(2)], [np.ones(3), np.ones(4), np.ones(3), np.ones(2)])
    assert info[0] > 0 and info[1] == 1 and info[2] != 0 and info[3] == 0


@pytest.mark.parametrize("N", [50, 64, 150, 300, 600, 1100])
def test_lu_blocked_singular_and_nonfinite_info(K, N):
    """info of the multi-launch path (register-resident panels): a zero column reports its 1-based index like
    getrf, a NaN reports -1, a regular matrix in the same batch is untouched by its neighbours."""
    rng = np.random.default_rng(N)
    good = rng.standard_normal((N, N))
    sing = good.copy()
    col = N // 2 + 3
    sing[col, :] = 0.0  # the kernels factor the transposed view: storage row = column of M
    nanm = good.copy()
    nanm[N // 3, N // 3] = np.nan
    rhs = rng.standard_normal(N)
    sol, info, _, _ = _lu_solve_gpu(K, [good, sing, nanm], [rhs, rhs, rhs])
    assert info[0] == 0 and info[1] == col + 1 and info[2] == -1
    assert rel_err(sol[0], np.linalg.solve(good, rhs)) <= 1e-14 * np.linalg.cond(good) * 10


def test_lu_empty_system(K):
    """N = 0 (everything active, m = 0: tests/pygradflow/test_newton.py:176-214) is a no-op."""
    Kt = torch.zeros(2, 1, 1, dtype=torch.float64, device="cuda")
    piv = torch.zeros(2, 1, dtype=torch.int32, device="cuda")
    info = torch.zeros(2, dtype=torch.int32, device="cuda")
    Nv = torch.zeros(2, dtype=torch.int32, device="cuda")
    rhs = torch.ones(2, 1, dtype=torch.float64, device="cuda")
    K.lu_factor(Kt, 1, Nv, piv, info, allw(K, 2))
    K.lu_solve(Kt, 1, Nv, piv, rhs, False, allw(K, 2))
    assert info.cpu().tolist() == [0, 0]


# ------------------------------------------------------------------------------------------------
def _ldlt_gpu(K, mats, rhss, npos=None):
    B = len(mats)
    Nmax = max(m.shape[0] for m in mats)
    ld = max(((Nmax + 63) // 64) * 64, 64)
    Kd = np.tile(np.eye(ld), (B, 1, 1))
    rd = np.zeros((B, ld))
    Nv = np.array([m.shape[0] for m in mats], dtype=np.int32)
    for k, (m_, r_) in enumerate(zip(mats, rhss)):
        N = m_.shape[0]
        Kd[k, :N, :N] = np.tril(m_) + np.triu(np.full((N, N), np.nan), 1)  # upper triangle must not be read
        rd[k, :N] = r_
    Kt, rt, Nt = dev(Kd), dev(rd), dev(Nv, torch.int32)
    dvec = torch.zeros(B, ld, dtype=torch.float64, device="cuda")
    info = torch.zeros(B, dtype=torch.int32, device="cuda")
    nneg = torch.zeros(B, dtype=torch.int32, device="cuda")
    K.ldlt_factor(Kt, Nmax, Nt, dvec, info, nneg, None if npos is None else dev(npos, torch.int32), allw(K, B))
    K.ldlt_solve(Kt, Nmax, Nt, rt, allw(K, B))
    return rt.cpu().numpy(), info.cpu().numpy(), nneg.cpu().numpy(), Kt.cpu().numpy(), dvec.cpu().numpy()


@pytest.mark.parametrize("N", [12, 48, 96, 200])
def test_ldlt_golden_kkt(K, golden, N):
    g = golden("linear_solver")
    Km, rhs, m = synth.kkt_instance(N)
    sol, info, nneg, _, _ = _ldlt_gpu(K, [Km], [rhs], npos=[N - m])
    assert info[0] == 0
    assert nneg[0] == m == int(g[f"kkt{N}/neg"])
    assert rel_err(sol[0, :N], g[f"kkt{N}/sol"]) <= 1e-10


@pytest.mark.parametrize("sizes", [[1, 2, 5, 63, 64, 65], [128, 129, 200, 256, 300], [768, 512, 705, 640]])
def test_ldlt_ragged_batch_factor_identity(K, sizes):
    """L D L' reproduces K (lower triangle), the solve matches LAPACK, inertia = (nI, m, 0)."""
    mats, rhss, ms = [], [], []
    for i, N in enumerate(sizes):
        Km, r, m = synth.kkt_instance(N, k=i + 1)
        mats.append(Km)
        rhss.append(r)
        ms.append(m)
    sol, info, nneg, fac, dvec = _ldlt_gpu(K, mats, rhss, npos=[N - m for N, m in zip(sizes, ms)])
    for k, N in enumerate(sizes):
        assert info[k] == 0 and nneg[k] == ms[k]
        L = np.tril(fac[k, :N, :N], -1) + np.eye(N)
        D = np.diag(fac[k, :N, :N]).copy()
        assert np.array_equal(D, dvec[k, :N])
        assert rel_err(np.tril((L * D) @ L.T), np.tril(mats[k])) <= 1e-12
        assert rel_err(sol[k, :N], np.linalg.solve(mats[k], rhss[k])) <= 1e-11


def test_ldlt_reference_fixtures_and_inertia(K, golden):
    """The 5x5 fixtures of tests/pygradflow/test_linear_solver.py: residual criterion and eigenvalue count."""
    g = golden("linear_solver")
    for name in ("indef", "posdef", "negdef"):
        mat, rhs = g[f"{name}/mat"], g["rhs"]
        sol, info, nneg, _, _ = _ldlt_gpu(K, [mat], [rhs])
        assert info[0] == 0
        assert np.allclose(mat @ sol[0, :5] - rhs, 0.0)
        assert nneg[0] == int(g[f"{name}/neg"])
        assert rel_err(sol[0, :5], g[f"{name}/sol"]) <= 1e-10


def test_ldlt_flags_non_quasidefinite(K):
    """A KKT matrix whose H block is indefinite must be flagged (info = -2) when the expected sign
    pattern is supplied -- the engine then refactors that instance with pivoted LU."""
    Km, r, m = synth.kkt_instance(96)
    bad = Km.copy()
    bad[3, 3] = -5.0
    _, info, _, _, _ = _ldlt_gpu(K, [Km, bad], [r, r], npos=[96 - m, 96 - m])
    assert info[0] == 0 and info[1] == -2


def test_ldlt_empty_system_resets_info(K):
    """N = 0 (everything active, m = 0): nothing to factorise, but a stale info word must not survive."""
    Kt = torch.zeros(2, 64, 64, dtype=torch.float64, device="cuda")
    dvec = torch.zeros(2, 64, dtype=torch.float64, device="cuda")
    info = torch.full((2,), 7, dtype=torch.int32, device="cuda")
    nneg = torch.full((2,), 5, dtype=torch.int32, device="cuda")
    Nv = torch.zeros(2, dtype=torch.int32, device="cuda")
    rhs = torch.ones(2, 64, dtype=torch.float64, device="cuda")
    K.ldlt_factor(Kt, 0, Nv, dvec, info, nneg, None, allw(K, 2))
    K.ldlt_solve(Kt, 0, Nv, rhs, allw(K, 2))
    K.ldlt_factor(Kt, 5, Nv, dvec, info, nneg, None, allw(K, 2))  # Nmax > 0 but every instance empty
    assert info.cpu().tolist() == [0, 0] and nneg.cpu().tolist() == [0, 0]


def test_ldlt_split_mode_matches_default(K):
    """The opt-in split launch scheme (chain kernel + 64-row panel kernel on separate streams, GF_LDLT_P64=1) runs in
    a subprocess (the switch is read once) and must give the same factors as the default scheme."""
    import os
    import subprocess
    import sys

    code = (
        "import os, sys, torch, numpy as np\n"
        "sys.path.insert(0, os.getcwd())\n"
        "from pygradflow_b200 import kernels as K, synth\n"
        "B, N = 1100, 200\n"
        "Km, r, m = synth.kkt_instance(N)\n"
        "ld = 256\n"
        "Kd = np.tile(np.eye(ld), (B, 1, 1)); Kd[:, :N, :N] = np.tril(Km)\n"
        "Kt = torch.as_tensor(Kd, device='cuda')\n"
        "i32 = dict(dtype=torch.int32, device='cuda')\n"
        "Nv = torch.full((B,), N, **i32); info = torch.zeros(B, **i32); nneg = torch.zeros(B, **i32)\n"
        "dvec = torch.zeros(B, ld, dtype=torch.float64, device='cuda')\n"
        "K.ldlt_factor(Kt, N, Nv, dvec, info, nneg, None, K.WorkList.all(B))\n"
        "torch.cuda.synchronize()\n"
        "assert int(info.abs().sum()) == 0 and int((nneg != m).sum()) == 0\n"
        "print(repr(float(torch.tril(Kt[:, :N, :N]).double().sum())), repr(float(dvec.sum())))\n"
    )
    outs = []
    for flag in ("0", "1"):
        env = dict(os.environ, GF_LDLT_P64=flag)
        res = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, env=env,
                             cwd=os.path.dirname(os.path.dirname(os.path.abspath(__file__))), timeout=300)
        assert res.returncode == 0, res.stderr[-2000:]
        outs.append(res.stdout.strip())
    assert outs[0] == outs[1]


@pytest.mark.parametrize("n,m", [(150, 40), (512, 256), (200, 0)])
def test_fused_assembly_factor_is_bit_identical(K, n, m):
    """gf_kkt_ldlt_factor (K gathered inside the factorisation kernels) == gf_kkt_assemble + gf_ldlt_factor, bit for
    bit, on ragged active sets."""
    B = 6
    rng = np.random.default_rng(n + m)
    d = synth.qp_batch(range(B), n, m)
    active = rng.uniform(size=(B, n)) < rng.uniform(0.0, 0.5, size=(B, 1))
    active[0] = False
    f64 = dict(dtype=torch.float64, device="cuda")
    i32 = dict(dtype=torch.int32, device="cuda")
    H, J = dev(d["H"]), (dev(d["A"]) if m else None)
    act = torch.as_tensor(active.astype(np.uint8), device="cuda")
    perm, nI, Nvec = torch.zeros((B, n), **i32), torch.zeros(B, **i32), torch.zeros(B, **i32)
    w = allw(K, B)
    K.index_sets(act, m, perm, nI, Nvec, w)
    dt, rho = dev(10.0 ** rng.uniform(-1, 1, B)), dev(10.0 ** rng.uniform(-4, 0, B))
    ld = ((n + m + 63) // 64) * 64
    out = []
    for fused in (False, True):
        Kt = torch.full((B, ld, ld), float("nan"), **f64)
        dvec, info, nneg = torch.zeros((B, ld), **f64), torch.zeros(B, **i32), torch.zeros(B, **i32)
        if fused:
            K.kkt_ldlt_factor(H, J, perm, nI, dt, rho, Nvec, Kt, dvec, info, nneg, w)
        else:
            K.kkt_assemble(H, J, perm, nI, dt, rho, Kt, 64, True, w)
            K.ldlt_factor(Kt, n + m, Nvec, dvec, info, nneg, nI, w)
        out.append((torch.tril(Kt).cpu().numpy(), dvec.cpu().numpy(), info.cpu().numpy(), nneg.cpu().numpy()))
    Nv = Nvec.cpu().numpy()
    for b in range(B):
        Np = ((Nv[b] + 63) // 64) * 64
        assert np.array_equal(out[0][0][b, :Np, :Np], out[1][0][b, :Np, :Np])
        assert np.array_equal(out[0][1][b, :Np], out[1][1][b, :Np])
    assert np.array_equal(out[0][2], out[1][2]) and np.array_equal(out[0][3], out[1][3])
    assert (out[0][2] == 0).all()


def test_count_newton_steps_matches_the_phase_expression(K):
    """gf_count_newton_steps: one step for phase 2..4, a second one for phase 3 / 4, added to the device counter."""
    gen = torch.Generator(device="cuda").manual_seed(5)
    for B in (1, 31, 256, 4097):
        phase = torch.randint(0, 7, (B,), generator=gen, device="cuda", dtype=torch.int32)
        cnt = torch.full((1,), 11, dtype=torch.int64, device="cuda")
        K.count_newton_steps(phase, cnt)
        want = 11 + int((((phase >= 2) & (phase <= 4)).sum() + ((phase == 3) | (phase == 4)).sum()).item())
        assert int(cnt.item()) == want


@pytest.mark.parametrize("sizes", [[1, 2, 5, 63, 64], [65, 66, 100, 127, 128, 9, 64], [128] * 5])
def test_ldlt_small_order_kernel_is_bit_identical_to_the_column_kernels(K, sizes, monkeypatch):
    """Padded order <= 128: one CTA per matrix with inv(L)' kept in the tile's own upper triangle (ldlt_small_kernel).
    Same products in the same order as the per-column kernels: factors, pivots, the scratch triangle, info and the
    inertia count must come out bit for bit the same (GF_LDLT_SMALL=0 selects the per-column kernels)."""
    mats, rhss, ms = [], [], []
    for i, N in enumerate(sizes):
        Km, r, m = synth.kkt_instance(N, k=i + 3)
        mats.append(Km)
        rhss.append(r)
        ms.append(m)
    npos = [N - m for N, m in zip(sizes, ms)]
    monkeypatch.setenv("GF_LDLT_SMALL", "1")
    sol1, info1, nneg1, fac1, d1 = _ldlt_gpu(K, mats, rhss, npos=npos)
    monkeypatch.setenv("GF_LDLT_SMALL", "0")
    sol0, info0, nneg0, fac0, d0 = _ldlt_gpu(K, mats, rhss, npos=npos)
    assert np.array_equal(info1, info0) and np.array_equal(nneg1, nneg0) and (info1 == 0).all()
    assert np.array_equal(d1, d0)
    assert np.array_equal(fac1, fac0)
    assert np.array_equal(sol1, sol0)
    for k, N in enumerate(sizes):
        assert rel_err(sol1[k, :N], np.linalg.solve(mats[k], rhss[k])) <= 1e-11
