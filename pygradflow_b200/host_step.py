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
