"""ctypes binding of libgradflow_b200.so (the C ABI declared in include/gradflow_b200.h).

There is no CPU fallback: if the shared library is missing or fails to load, every entry point raises.
"""

from __future__ import annotations

import ctypes
import os
from ctypes import c_double, c_int, c_void_p

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("GF_LIB_PATH") or os.path.join(_HERE, "csrc", "libgradflow_b200.so")  # GF_LIB_PATH: A/B of builds

_P = c_void_p  # device pointer
_I = c_int
_D = c_double
_WORK = [_P, _P, _I, _P]  # work, nwork_dev, nwork, stream

# name -> argtypes (all return int).  Must list every symbol include/gradflow_b200.h declares.
SIGNATURES = {
    "gf_version": [],
    "gf_qp_eval": [_I, _I, _I, _P, _P, _P, _P, _P, _P, _P, _P] + _WORK,
    "gf_rosen_eval": [_I, _I, _P, _P, _P, _P, _P] + _WORK,
    "gf_rosen_hess": [_I, _I, _P, _P, _P, _P] + _WORK,
    "gf_ocp_eval": [_I, _I, _I, _I, _D, _P, _P, _P, _P, _P, _P, _P, _P, _P] + _WORK,
    "gf_ocp_jac": [_I, _I, _I, _I, _D, _P, _P, _P, _P] + _WORK,
    "gf_ocp_hess": [_I, _I, _I, _I, _D, _P, _P, _P, _P, _P] + _WORK,
    "gf_aug_lag_grad": [_I, _I, _I, _P, _P, _P, _P, _P, _P, _P, _P] + _WORK,
    "gf_residual": [_I, _I, _I, _P, _P, _P, _P, _P, _P, _P, _P, _P, _I, _I, _P, _P, _P] + _WORK,
    "gf_residual_tau": [_I, _I, _I, _P, _P, _P, _P, _P, _P, _P, _P, _P, _P, _I, _I, _P, _P, _P] + _WORK,
    "gf_index_sets": [_I, _I, _I, _P, _P, _P, _P] + _WORK,
    "gf_kkt_assemble": [_I, _I, _I, _I, _I, _I, _P, _P, _P, _P, _P, _P, _P] + _WORK,
    "gf_kkt_rhs": [_I, _I, _I, _I, _P, _P, _P, _P, _P, _P, _P, _P] + _WORK,
    "gf_kkt_assemble_full": [_I, _I, _I, _I, _I, _P, _P, _P, _P, _P, _P, _P, _P] + _WORK,
    "gf_kkt_rhs_full": [_I, _I, _I, _I, _I, _P, _P, _P, _P, _P, _P, _P] + _WORK,
    "gf_lu_factor": [_I, _I, _I, _P, _P, _P, _P] + _WORK,
    "gf_lu_solve": [_I, _I, _I, _P, _P, _P, _P, _I, _I] + _WORK,
    "gf_ldlt_factor": [_I, _I, _I, _P, _P, _P, _P, _P, _P] + _WORK,
    "gf_ldlt_solve": [_I, _I, _I, _P, _P, _P, _I] + _WORK,
    "gf_step_finish": [_I, _I, _I, _I, _P, _P, _P, _P, _P, _P, _P, _P, _P, _P, _P, _P, _P, _P, _P] + _WORK,
    "gf_check_terminate": [_I, _I, _I, _P, _P, _P, _P, _P, _P, _P, _P, _D, _D, _D, _D, _I, _P, _P, _P] + _WORK,
    "gf_dr_first": [_I, _P, _P, _P, _P, _P, _D, _D, _D, _P, _P, _P],
    "gf_dr_second": [_I, _P, _P, _P, _D, _D, _D, _D, _D, _D, _P, _P, _P, _P, _P],
    "gf_count_newton_steps": [_I, _P, _P, _P],
    "gf_single_control": [_I, _I, _P, _P, _P, _P, _P, _D, _D, _D, _D, _D, _D, _D, _D, _D, _P, _P, _P, _P, _P, _P],
    "gf_exact_control": [_I, _I, _I, _I, _P, _P, _P, _P, _P, _D, _D, _P, _P, _P, _P, _P, _P],
    "gf_commit": [_I, _I, _I, _P, _P, _D, _I] + [_P] * 10 + [_P] * 10 + [_P],
    "gf_merit_grad": [_I, _I, _I, _P, _P, _P, _P, _P, _P, _P, _P, _P, _P] + _WORK,
    "gf_ls_begin": [_I, _P, _D, _P, _P, _P] + _WORK,
    "gf_ls_trial": [_I, _I, _I, _P, _P, _P, _P, _P, _P, _P] + _WORK,
    "gf_armijo_residual": [_I, _I, _I, _P, _P, _P, _P, _P, _P, _P, _P, _P, _P, _P, _D, _I, _P, _P, _P, _P] + _WORK,
    "gf_kkt_ldlt_factor": [_I, _I, _I, _I, _P, _P, _P, _P, _P, _P, _P, _P, _P, _P, _P] + _WORK,
    "gf_band_assemble": [_I, _I, _I, _I, _P, _P, _P, _P, _P, _P, _P] + _WORK,
    "gf_band_factor": [_I, _I, _I, _P, _P, _P] + _WORK,
    "gf_band_solve": [_I, _I, _I, _P, _P] + _WORK,
    "gf_band_permute": [_I, _I, _I, _I, _P, _P, _P, _P, _P, _I] + _WORK,
    "gf_ocp_jac_banded": [_I, _I, _I, _I, _D, _P, _P, _P, _P, _I] + _WORK,
    "gf_ocp_hess_diag": [_I, _I, _I, _I, _D, _P, _P, _P, _P, _P] + _WORK,
    "gf_stage_aug_lag_grad": [_I, _I, _I, _I, _P, _P, _P, _P, _P, _P, _P, _P] + _WORK,
    "gf_stage_kkt_factor": [_I, _I, _I, _I, _P, _P, _P, _P, _P, _P, _P, _P, _P, _P] + _WORK,
    "gf_stage_kkt_solve": [_I, _I, _I, _I, _P, _P, _P, _P, _P, _P, _P, _P, _P, _P, _I] + _WORK,
    "gf_ldexp": [_I, _I, _I, _P, _P, _I, _P, _I, _P, _I, _P] + _WORK,
    "gf_h2d_sym_lower": [_P, _P, _I, _I, _I, _P],
    "gf_symmetrize_lower": [_P, _I, _I, _I, _P],
    "gf_krylov_scratch_rows": [_I, _I],
    "gf_gmres_solve": [_I, _I, _I, _P, _P, _P, _I, _P, _P, _I, _I, _I, _D, _D, _P, _P, _P] + _WORK,
    "gf_minres_solve": [_I, _I, _I, _P, _P, _P, _I, _P, _D, _P, _P, _P] + _WORK,
    "gf_pareto_update": [_I, _I, _I, _P, _P, _P, _P, _P, _P, _D, _D, _P] + _WORK,
    "gf_filter_update": [_I, _I, _I, _I, _P, _P, _P, _P, _P, _P, _P, _P, _P, _P, _P, _P, _I, _P] + _WORK,
    "gf_rosen_fused_solve": [_I, _I] + [_P] * 15 + [_P, _I, _I, _I] + _WORK,
    "gf_hess_rho": [_I, _I, _I, _P, _P, _P, _P] + _WORK,
    "gf_build_worklist": [_I, _P, _I, _I, _I, _P, _P, _P, _P, _P],
    "gf_dt_from_lamb": [_I, _P, _P, _P],
}

_lib = None


class NativeError(RuntimeError):
    pass


def load() -> ctypes.CDLL:
    """Load the CUDA library (built in-tree by pygradflow_b200.build).  Fails loudly if absent."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise NativeError(
                f"{LIB_PATH} is missing: build it with `python -m pygradflow_b200.build` "
                "(pygradflow_b200 has no CPU fallback)"
            )
        lib = ctypes.CDLL(LIB_PATH)
        for name, argtypes in SIGNATURES.items():
            fn = getattr(lib, name)  # AttributeError if the library does not export it
            fn.argtypes = argtypes
            fn.restype = c_int
        _lib = lib
    return _lib


def check(rc: int, name: str) -> None:
    if rc != 0:
        if rc >= 1000:
            raise NativeError(f"{name}: CUDA error {rc - 1000}")
        raise NativeError(f"{name}: {'bad argument' if rc == -1 else 'unsupported shape' if rc == -2 else rc}")


def ptr(t) -> int | None:
    """Device pointer of a torch tensor (None -> NULL)."""
    if t is None:
        return None
    assert t.is_cuda and t.is_contiguous(), "native kernels need contiguous CUDA tensors"
    return t.data_ptr()


def call(name: str, *args) -> None:
    lib = load()
    check(getattr(lib, name)(*args), name)
