"""Shared helpers of the parity tests."""

import numpy as np

THETA_NOISE = 1e-8


def rel_err(a, b):
    a = np.asarray(a, dtype=np.float64)
    b = np.asarray(b, dtype=np.float64)
    assert a.shape == b.shape, (a.shape, b.shape)
    if a.size == 0:
        return 0.0
    return float(np.max(np.abs(a - b)) / max(1.0, float(np.max(np.abs(b)))))


def noise_horizon(trace):
    """Index of the first outer iteration whose contraction ratio theta = |d2|/|d1| is rounding
    noise (the second simplified-Newton step of a QP whose frozen active set was already solved
    exactly is ~1e-16).  The reference feeds log(theta) into its PI step-size controller
    (distance_ratio_control.py:57-63), so from there on lambda -- and the trajectory -- depends on
    the last bits of the linear solve; strict 1e-10 parity is only meaningful before it."""
    for i, t in enumerate(trace):
        th = t["theta"]
        if th == th and th < THETA_NOISE:
            return i
    return len(trace)
