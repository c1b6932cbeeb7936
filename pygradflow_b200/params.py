"""Solver parameters of the hot path -- field names and defaults of the reference's ``Params``
(pygradflow/params.py:197-265) for every field the Newton/KKT path reads, so a reference ``Params``
object can be passed in unchanged (duck typing) and this one can be handed to reference-style code.
"""

from __future__ import annotations

import enum
import math
from dataclasses import dataclass
from typing import Any, Callable, Optional

import numpy as np


class NewtonType(enum.Enum):
    """pygradflow/params.py:19-46."""

    Simplified = enum.auto()
    Full = enum.auto()
    ActiveSet = enum.auto()
    Globalized = enum.auto()


class LinearSolverType(enum.Enum):
    """LU is the reference default (params.py:236).  LDLT is the B200 symmetric factorisation with
    inertia (the contract of the reference's MA57 / MUMPS / SSIDS / Cholesky wrappers); Auto picks LDLT
    for quasi-definite systems of order > 112 with a per-instance pivoted-LU fallback, LU otherwise -- and
    Banded (banded LDLT in the family's own KKT ordering) when the problem family provides one."""

    LU = enum.auto()
    LDLT = enum.auto()
    Auto = enum.auto()
    Banded = enum.auto()
    BlockTri = enum.auto()  # stage-structured families (cfg4): Schur complement on the multipliers + block cyclic reduction
    GMRES = enum.auto()   # linear_solver/gmres_solver.py: scipy's restarted GMRES(20), atol 1e-8, at most n restarts
    MINRES = enum.auto()  # linear_solver/minres_solver.py: scipy's MINRES (symmetric formulation only)


class StepSolverType(enum.Enum):
    """pygradflow/params.py:50-70: formulation of the Newton step system (step/solver/__init__.py:12-31).  Symmetric
    is the reduced quasi-definite KKT system (LDL' or LU); Asymmetric and Extended are the full-order unsymmetric
    systems of asymmetric_step_solver.py / extended_step_solver.py, Standard the derivative of the unscaled implicit
    function with H_rho = H + rho J'J (standard_step_solver.py) -- all three through the pivoted LU."""

    Standard = enum.auto()
    Extended = enum.auto()
    Symmetric = enum.auto()
    Asymmetric = enum.auto()


class ScalingType(enum.Enum):
    """pygradflow/params.py:166-195: how create_scaling (scale.py:234-280) chooses the power-of-two weights."""

    NoScaling = enum.auto()
    GradJac = enum.auto()
    KKT = enum.auto()
    Nominal = enum.auto()
    Custom = enum.auto()


class ActiveSetType(enum.Enum):
    """pygradflow/params.py:14-18: how the point that decides the active set is chosen (newton_control.py:60-88)."""

    Standard = enum.auto()
    Explicit = enum.auto()
    SmallestActiveSet = enum.auto()
    LargestActiveSet = enum.auto()


class StepControlType(enum.Enum):
    """pygradflow/params.py:113-130, the Newton-based controllers (step_control.py:123-150).  Optimizing / BoxReduced
    solve the proximal sub-problem with Ipopt / a box solver instead of the Newton-KKT path and are out of scope."""

    DistanceRatio = enum.auto()
    ResiduumRatio = enum.auto()
    Exact = enum.auto()
    Fixed = enum.auto()


class PenaltyUpdate(enum.Enum):
    """pygradflow/params.py:122-128, penalty.py:36-255.  The two filters can veto an accepted step (solver.py:357-378)."""

    Constant = enum.auto()
    DualNorm = enum.auto()
    DualEquilibration = enum.auto()
    ParetoDecrease = enum.auto()
    ObjectiveFilter = enum.auto()
    LagrangianFilter = enum.auto()


def _enum_name(value) -> str:
    return value.name if isinstance(value, enum.Enum) else str(value)


@dataclass
class Params:
    rho: float = 1e-8
    theta_max: float = 0.9
    theta_ref: float = 0.5
    lamb_init: float = 1.0
    lamb_min: float = 1e-12
    lamb_max: float = 1e12
    lamb_inc: float = 2.0
    lamb_red: float = 0.5
    K_P: float = 0.2
    K_I: float = 0.005
    opt_tol: float = 1e-6
    active_tol: float = 1e-8
    local_infeas_tol: float = 1e-8
    newton_type: NewtonType = NewtonType.Simplified
    newton_tol: float = 1e-8
    step_control_type: StepControlType = StepControlType.DistanceRatio
    active_set_type: ActiveSetType = ActiveSetType.Standard
    active_set_tau: Optional[float] = None
    step_solver: Optional[Callable[..., Any]] = None
    step_solver_type: StepSolverType = StepSolverType.Symmetric
    linear_solver_type: LinearSolverType = LinearSolverType.Auto
    penalty_update: PenaltyUpdate = PenaltyUpdate.DualNorm
    scaling_type: ScalingType = ScalingType.NoScaling
    scaling_primal: Optional[Any] = None
    scaling_dual: Optional[Any] = None
    scaling: Optional[Any] = None  # Custom: a scale.BatchedScaling or (var_weights, cons_weights[, obj_weight])
    iteration_limit: Optional[int] = None
    obj_lower_limit: float = -1e10
    inertia_correction: bool = False
    report_rcond: bool = False
    # entries each instance's penalty filter can hold (the reference's list is unbounded, penalty.py:176; the entries
    # are mutually non-dominated, so the list stays far shorter than the number of accepted steps); exceeding it raises
    penalty_filter_capacity: int = 1024
    # families with a fused persistent solver (problem.fused_family(); cfg2: one warp runs a whole instance) use it when
    # the other parameters are the defaults it implements; False forces the lock-step driver
    fused: bool = True

    def __post_init__(self):
        for key, cls in (("newton_type", NewtonType), ("linear_solver_type", LinearSolverType),
                         ("penalty_update", PenaltyUpdate), ("step_control_type", StepControlType),
                         ("active_set_type", ActiveSetType), ("step_solver_type", StepSolverType),
                         ("scaling_type", ScalingType)):
            v = getattr(self, key)
            if not isinstance(v, cls):
                setattr(self, key, cls[_enum_name(v)])  # accepts strings and the reference's own enums

    @property
    def dtype(self):
        return np.float64

    @property
    def log_theta_ref(self) -> float:
        return math.log(self.theta_ref)

    @staticmethod
    def from_reference(ref) -> "Params":
        """Copy the hot-path fields out of a reference ``pygradflow.params.Params`` (or any look-alike)."""
        kw = {}
        for f in Params.__dataclass_fields__:
            if f in ("step_solver", "linear_solver_type", "scaling"):
                continue
            if hasattr(ref, f):
                v = getattr(ref, f)
                if f == "penalty_update" and _enum_name(v) not in PenaltyUpdate.__members__:
                    raise ValueError(f"penalty_update={_enum_name(v)} is not a PenaltyUpdate")
                if f == "step_solver_type" and _enum_name(v) not in StepSolverType.__members__:
                    raise ValueError(f"step_solver_type={_enum_name(v)} is outside the B200 path")
                if f == "step_control_type" and _enum_name(v) not in StepControlType.__members__:
                    raise ValueError(f"step_control_type={_enum_name(v)} is outside the B200 path (Newton-based only)")
                kw[f] = v
        # the reference's direct solvers (LU, MA57, ...) map to the engine's own choice; its iterative ones carry over
        lin = _enum_name(getattr(ref, "linear_solver_type", "Auto"))
        if lin in ("GMRES", "MINRES"):
            kw["linear_solver_type"] = LinearSolverType[lin]
        return Params(**kw)
