"""Shared comparison helpers for the parity tests."""
import numpy as np

REL_TOL = 1e-5      # BASELINE.json north_star: "within 1e-5 relative for cutouts, attention features and scores"


def rel_err(got, want):
    """Max absolute error relative to the reference tensor's magnitude."""
    got = np.asarray(got, dtype=np.float64)
    want = np.asarray(want, dtype=np.float64)
    scale = max(float(np.abs(want).max()), 1e-30)
    return float(np.abs(got - want).max()) / scale


def assert_rel(got, want, tol=REL_TOL, what=""):
    e = rel_err(got, want)
    assert e <= tol, "%s relative error %.3e > %.1e" % (what, e, tol)
