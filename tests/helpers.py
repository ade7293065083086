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


def cutout_mismatch_report(got, want, diag, tol=REL_TOL, idx_eps=2e-4):
    """Compare a GPU cutout with the oracle's.

    Returns (n_bad, n_excused, worst_unexcused).  A sample may differ by more than
    `tol` only where the oracle's own index arithmetic sits within `idx_eps` of a
    discontinuity (rint boundary in area mode, the outbound thresholds, the
    area/linear decision, or the call-global s_area): there a 1-ulp difference in
    the float32 arctangent legitimately flips the result (SURVEY.md §7 hard part 1).
    got/want: [M, S, P]; diag from oracle.cutout.cutout_diagnostics ([S, M, ...]).
    """
    err = np.abs(got.astype(np.float64) - want.astype(np.float64))
    scale = max(float(np.abs(want).max()), 1e-30)
    bad = err > tol * scale                                    # [M, S, P]
    near = (diag["rint_margin"] < idx_eps) | (diag["edge_margin"] < idx_eps)       # [S, M, P]
    near = near | (diag["span_margin"] < idx_eps)[..., None]
    near = near.transpose(1, 0, 2)
    if diag.get("s_area", 0) and diag.get("s_area_margin", 1.0) < idx_eps:
        near = np.ones_like(near)
    unexcused = bad & ~near
    worst = float(err[unexcused].max()) if unexcused.any() else 0.0
    return int(bad.sum()), int((bad & near).sum()), worst
