"""Build libgradflow_b200.so in-tree with nvcc for sm_100a (no torch headers: the boundary is a plain C ABI).

Each translation unit is compiled to an object file under csrc/build/ (in parallel, only when its source or a shared
header changed) and the objects are linked into the shared library."""

from __future__ import annotations

import glob
import os
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor

CSRC = os.path.join(os.path.dirname(os.path.abspath(__file__)), "csrc")
OBJDIR = os.path.join(CSRC, "build")
LIB = os.path.join(CSRC, "libgradflow_b200.so")
SOURCES = ["gf_api.cu", "gf_eval.cu", "gf_step.cu", "gf_lu.cu", "gf_ldlt.cu", "gf_linesearch.cu", "gf_xfer.cu",
           "gf_band.cu", "gf_scale.cu", "gf_blocktri.cu", "gf_krylov.cu", "gf_penalty.cu", "gf_fused.cu",
           "gf_syrk.cu"]
NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-O3", "-lineinfo", "-std=c++17", "-Xptxas=-v",
    "-Xcompiler", "-fPIC",
]


def _headers():
    return glob.glob(os.path.join(CSRC, "*.cuh")) + [
        os.path.join(os.path.dirname(CSRC), "..", "include", "gradflow_b200.h")]


def _sources():
    return [os.path.join(CSRC, s) for s in SOURCES if os.path.exists(os.path.join(CSRC, s))]


def needs_build() -> bool:
    if not os.path.exists(LIB):
        return True
    t = os.path.getmtime(LIB)
    deps = _sources() + _headers()
    return any(os.path.getmtime(d) > t for d in deps if os.path.exists(d))


def _compile(nvcc, src, obj):
    cmd = [nvcc] + NVCC_FLAGS + ["-c", "-o", obj, src]
    res = subprocess.run(cmd, capture_output=True, text=True)
    return cmd, res


def build(force: bool = False, verbose: bool = False) -> str:
    if not force and not needs_build():
        return LIB
    nvcc = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
    os.makedirs(OBJDIR, exist_ok=True)
    hdr_t = max(os.path.getmtime(h) for h in _headers() if os.path.exists(h))
    jobs, objs = [], []
    for src in _sources():
        obj = os.path.join(OBJDIR, os.path.basename(src)[:-3] + ".o")
        objs.append(obj)
        stale = force or not os.path.exists(obj) or os.path.getmtime(obj) < max(os.path.getmtime(src), hdr_t)
        if stale:
            jobs.append((src, obj))
    log = []
    failed = False
    with ThreadPoolExecutor(max_workers=max(1, min(len(jobs), os.cpu_count() or 4))) as ex:
        for cmd, res in ex.map(lambda j: _compile(nvcc, *j), jobs):
            log.append(" ".join(cmd) + "\n" + res.stdout + res.stderr)
            if verbose or res.returncode != 0:
                sys.stderr.write(res.stdout + res.stderr)
            failed |= res.returncode != 0
    if failed:
        raise RuntimeError("nvcc failed building libgradflow_b200.so")
    cmd = [nvcc, "-gencode", "arch=compute_100a,code=sm_100a", "-shared", "-o", LIB] + objs
    res = subprocess.run(cmd, capture_output=True, text=True)
    log.append(" ".join(cmd) + "\n" + res.stdout + res.stderr)
    if res.returncode != 0:
        sys.stderr.write(res.stdout + res.stderr)
        raise RuntimeError("nvcc failed linking libgradflow_b200.so")
    with open(os.path.join(CSRC, "build.log"), "a" if not force else "w") as f:
        f.write("\n".join(log))
    return LIB


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose=True))
