"""Newton-KKT steps on HOST buffers: the end-to-end shape of the plug-in boundary.

The reference hands its step solver host arrays (``Iterate.x``, ``Problem`` callbacks returning
NumPy / SciPy objects: pygradflow/step/solver/scaled_step_solver.py:76-79).  This module is the batched
equivalent: problem data and iterates live in pinned host memory, are streamed to the GPU in chunks on a
copy stream while the previous chunk computes on another, and the step results stream back.  It is what
bench.py times as ``e2e`` (host-to-device and device-to-host copies inside the timed region).
"""

from __future__ import annotations

from typing import Dict

import torch

from . import kernels as K
from .newton import NewtonKKTStepper
from .params import LinearSolverType
from .problem import BatchedQP


class _Slot:
    def __init__(self, C, n, m, device, linear):
        f64 = dict(dtype=torch.float64, device=device)
        q = object.__new__(BatchedQP)
        q.var_lb = torch.zeros((C, n), **f64)
        q.var_ub = torch.zeros((C, n), **f64)
        q.B, q.n, q.m, q.device = C, n, m, torch.device(device)
        q.H = torch.zeros((C, n, n), **f64)
        q.g = torch.zeros((C, n), **f64)
        q.A = torch.zeros((C, m, n), **f64) if m > 0 else None
        q.b = torch.zeros((C, m), **f64) if m > 0 else None
        self.qp = q
        self.x = torch.zeros((C, n), **f64)
        self.y = torch.zeros((C, m), **f64)
        self.lamb = torch.zeros((C,), **f64)
        self.rho = torch.zeros((C,), **f64)
        self.stepper = NewtonKKTStepper(q, linear)
        self.ready = torch.cuda.Event()
        self.done = torch.cuda.Event()


class HostNewtonKKT:
    """Chunked, double-buffered ``newton_kkt_step`` for the QP family with host-resident inputs."""

    INPUT_KEYS = ("H", "A", "g", "b", "lb", "ub", "x", "y", "lamb", "rho")

    def __init__(self, n: int, m: int, chunk: int = 256, device="cuda",
                 linear: LinearSolverType = LinearSolverType.Auto, nbuf: int = 2):
        self.n, self.m, self.chunk, self.device = n, m, chunk, device
        self.slots = [_Slot(chunk, n, m, device, linear) for _ in range(nbuf)]
        self.copy_stream = torch.cuda.Stream(device=device)
        self.compute_stream = torch.cuda.Stream(device=device)

    @staticmethod
    def pinned_like(t: torch.Tensor) -> torch.Tensor:
        out = torch.empty(t.shape, dtype=t.dtype, device="cpu", pin_memory=True)
        out.copy_(t)
        return out

    def bytes_per_step(self, B: int):
        """Bytes that cross PCIe per step (H: lower block triangle only, it is symmetric)."""
        n, m = self.n, self.m
        h2d = K.h2d_sym_lower_bytes(B, n) + 8 * B * (m * n + 4 * n + 2 * m + 2)
        d2h = 8 * B * (n + m + 2) + 4 * B
        return h2d, d2h

    def step(self, host: Dict[str, torch.Tensor], out: Dict[str, torch.Tensor]) -> None:
        """host: pinned tensors keyed by INPUT_KEYS ([B, ...]); out: pinned xn, yn, diff, fnorm, info."""
        B = host["x"].shape[0]
        C, m = self.chunk, self.m
        sc, sk = self.copy_stream, self.compute_stream
        cur = torch.cuda.current_stream()
        sc.wait_stream(cur)
        sk.wait_stream(cur)
        for i, lo in enumerate(range(0, B, C)):
            hi = min(B, lo + C)
            cnt = hi - lo
            s = self.slots[i % len(self.slots)]
            with torch.cuda.stream(sc):
                sc.wait_event(s.done)  # the slot's previous chunk has been consumed
                q = s.qp
                K.h2d_sym_lower(q.H, host["H"][lo:hi], cnt)
                pairs = [(q.g, "g"), (q.var_lb, "lb"), (q.var_ub, "ub"), (s.x, "x"), (s.lamb, "lamb"), (s.rho, "rho")]
                if m > 0:
                    pairs += [(q.A, "A"), (q.b, "b"), (s.y, "y")]
                for dst, key in pairs:
                    dst[:cnt].copy_(host[key][lo:hi], non_blocking=True)
                s.ready.record(sc)
            with torch.cuda.stream(sk):
                sk.wait_event(s.ready)
                K.symmetrize_lower(s.qp.H, cnt)
                s.stepper.work.nwork = cnt
                xn, yn, diff, fnorm, info = s.stepper.step(s.x, s.y, s.lamb, s.rho)
                out["xn"][lo:hi].copy_(xn[:cnt], non_blocking=True)
                if m > 0:
                    out["yn"][lo:hi].copy_(yn[:cnt], non_blocking=True)
                out["diff"][lo:hi].copy_(diff[:cnt], non_blocking=True)
                out["fnorm"][lo:hi].copy_(fnorm[:cnt], non_blocking=True)
                out["info"][lo:hi].copy_(info[:cnt], non_blocking=True)
                s.done.record(sk)
        cur.wait_stream(sk)
        cur.wait_stream(sc)


class ResidentNewtonKKT:
    """``newton_kkt_step`` on host iterates for a problem registered ONCE.

    The reference builds its ``Problem`` once and hands the step solver a new ``Iterate`` every step
    (pygradflow/solver.py:305-316, step/solver/__init__.py:18-19): the problem data (H, A, g, b, bounds of the QP
    family) does not change between the Newton steps of a solve.  ``register`` therefore uploads it once; ``step`` copies
    this step's inputs -- the iterates x, y and the per-instance lambda, rho -- from pinned host memory, runs the
    batched step on the resident data and copies (xn, yn, diff, fnorm, info) back to pinned host memory.
    ``HostNewtonKKT`` above is the variant for data that changes every step or does not fit into HBM."""

    def __init__(self, problem: BatchedQP, linear: LinearSolverType = LinearSolverType.Auto, stepper=None):
        self.problem = problem
        B, n, m, dev = problem.B, problem.n, problem.m, problem.device
        f64 = dict(dtype=torch.float64, device=dev)
        self.stepper = stepper if stepper is not None else NewtonKKTStepper(problem, linear)
        self.x = torch.zeros((B, n), **f64)
        self.y = torch.zeros((B, m), **f64)
        self.lamb = torch.zeros((B,), **f64)
        self.rho = torch.zeros((B,), **f64)

    @classmethod
    def register(cls, host: Dict[str, torch.Tensor], device="cuda", linear: LinearSolverType = LinearSolverType.Auto):
        """host: (pinned) H [B,n,n], A [B,m,n] or None, g, b, lb, ub."""
        prob = BatchedQP(host["H"], host.get("A"), host["g"], host.get("b"), host["lb"], host["ub"], device=device)
        return cls(prob, linear)

    def bytes_per_step(self):
        B, n, m = self.problem.B, self.problem.n, self.problem.m
        return 8 * B * (n + m + 2), 8 * B * (n + m + 2) + 4 * B

    def step(self, host: Dict[str, torch.Tensor], out: Dict[str, torch.Tensor]) -> None:
        """host: pinned x [B,n], y [B,m], lamb [B], rho [B]; out: pinned xn, yn, diff, fnorm, info.  Asynchronous on
        the current stream (synchronise before reading `out`)."""
        m = self.problem.m
        self.x.copy_(host["x"], non_blocking=True)
        if m > 0:
            self.y.copy_(host["y"], non_blocking=True)
        self.lamb.copy_(host["lamb"], non_blocking=True)
        self.rho.copy_(host["rho"], non_blocking=True)
        xn, yn, diff, fnorm, info = self.stepper.step(self.x, self.y, self.lamb, self.rho)
        out["xn"].copy_(xn, non_blocking=True)
        if m > 0:
            out["yn"].copy_(yn, non_blocking=True)
        out["diff"].copy_(diff, non_blocking=True)
        out["fnorm"].copy_(fnorm, non_blocking=True)
        out["info"].copy_(info, non_blocking=True)
