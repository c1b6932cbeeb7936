"""The C-ABI shared library builds for sm_100a, loads, and exports every symbol include/gradflow_b200.h
declares (no compute calls: this runs without a GPU)."""

import ctypes
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def header_symbols():
    hdr = open(os.path.join(ROOT, "include", "gradflow_b200.h")).read()
    decls = {}
    for m in re.finditer(r"^int (gf_\w+)\((.*?)\);", hdr, re.M | re.S):
        args = [a for a in m.group(2).split(",") if a.strip() and a.strip() != "void"]
        decls[m.group(1)] = len(args)
    return decls


@pytest.fixture(scope="module")
def lib_path():
    from pygradflow_b200 import build

    return build.build()


def test_library_exports_every_header_symbol(lib_path):
    lib = ctypes.CDLL(lib_path)
    decls = header_symbols()
    assert len(decls) >= 20
    for name in decls:
        assert hasattr(lib, name), f"{name} declared in the header but not exported"
    assert lib.gf_version() >= 100


def test_binding_matches_header(lib_path):
    from pygradflow_b200 import native

    decls = header_symbols()
    assert set(decls) == set(native.SIGNATURES)
    for name, nargs in decls.items():
        assert len(native.SIGNATURES[name]) == nargs, name
    native.load()


def test_header_cites_reference_interfaces():
    hdr = open(os.path.join(ROOT, "include", "gradflow_b200.h")).read()
    for needle in ("params.py:234", "step_solver.py:66-130", "linear_solver.py:18-31", "lu_solver.py"):
        assert needle in hdr


def test_bad_arguments_are_rejected_without_a_gpu(lib_path):
    """Argument validation happens before any launch, so it is testable on the CPU box."""
    from pygradflow_b200 import native

    lib = native.load()
    assert lib.gf_lu_factor(0, 4, 4, None, None, None, None, None, None, 0, None) == -1
    assert lib.gf_ldlt_factor(1, 65, 10, None, None, None, None, None, None, None, None, 1, None) == -1  # ld % 64


def test_product_package_never_imports_the_oracle():
    pkg = os.path.join(ROOT, "pygradflow_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith(".py"):
                src = open(os.path.join(dirpath, f)).read()
                assert "import oracle" not in src and "from oracle" not in src, f
