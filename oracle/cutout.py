"""CPU oracle for the distance-adaptive polar cutout.  TEST INFRASTRUCTURE ONLY.

A NumPy restatement of the reference's `scans_to_cutout`
(/root/reference/src/utils/utils.py:259-334).  Only `tests/`,
`__graft_entry__.smoke()` and `bench.py`'s CPU-baseline legs may import it; the
product path (planar_optical_flow_b200) never does.

Pinning: the reference holds no golden vectors for this path, so
the oracle is pinned against the reference itself, imported unmodified in the
build container (`oracle/ref_shim.py`); `tests/test_oracle_vs_reference.py`
asserts BIT-EQUAL outputs and `oracle/make_golden.py` freezes reference outputs
into `tests/golden/*.npz`.

The restatement keeps the reference's *mixed* precision exactly, because the
outputs are discontinuous in the sample index (SURVEY.md §7 hard part 1):

  * the window half-angle and the angular step are computed in the dtype of
    `scans` (float32 in every reference caller)              utils.py:279,282
  * `phi - half_alpha` is evaluated in promote(phi, scans)    utils.py:283-285
  * the sample angle, the fractional index and the interpolation ratio are
    float64 (int64 arange * float32 step promotes to float64) utils.py:286-294
  * the neighbour difference `hi - lo` is formed in the scan dtype before it
    meets the float64 ratio                                   utils.py:300
  * area-mode means are accumulated in the scan dtype         utils.py:319-322
  * the depth clip uses bounds rounded to the scan dtype      utils.py:327

Layout differs from the reference on purpose: work arrays here are
[S, M, P] (sample axis last) so the result needs no final transpose; every
per-element operation and its rounding is the same.
"""
import math

import numpy as np


def scans_to_cutout(scans, scan_phi, stride=1, centered=True, fixed=False,
                    window_width=1.66, window_depth=1.0, num_cutout_pts=48,
                    padding_val=29.99, area_mode=False, half_alpha=None):
    """Return the cutout tensor [M, S, P] float32, M = ceil(N / stride).

    Same signature, defaults and result as utils.py:259-334.  `half_alpha`
    ([S, M], test-only) overrides the arctangent of :279 so that the one
    platform-dependent step (NumPy's SIMD float32 arctan) can be held fixed or
    perturbed by an ulp when comparing implementations.
    """
    scans = np.asarray(scans)
    scan_phi = np.asarray(scan_phi)
    n_scans, n_pts = scans.shape
    P = int(num_cutout_pts)

    centre_r = scans[:, ::stride]                                   # utils.py:274-278
    if not fixed:
        centre_r = np.broadcast_to(scans[-1, ::stride], centre_r.shape)
    centre_phi = scan_phi[::stride]
    last = n_pts - 1

    # window half-angle, scan dtype                                   utils.py:279
    half = window_half_angle(scans, stride, fixed, window_width) if half_alpha is None \
        else np.asarray(half_alpha, dtype=centre_r.dtype)
    origin = scan_phi[0]
    pitch = scan_phi[1] - scan_phi[0]

    def fractional_index(n_samples):
        """Index of each of `n_samples` equi-angular samples, [S, M, n] f64."""
        step = 2.0 * half / (n_samples - 1)                         # utils.py:282,310
        start = centre_phi - half                                   # utils.py:284-285
        k = np.arange(n_samples)
        ang = start[..., None] + k * step[..., None]                # utils.py:286,314
        return (ang - origin) / pitch                               # utils.py:288,317

    idx = fractional_index(P)
    outside = (idx < 0) | (idx > last)                              # utils.py:289

    # two-tap linear resampling                                       utils.py:292-300
    lo = np.clip(np.floor(idx), 0, last).astype(np.int64)
    hi = np.minimum(lo + 1, last)
    frac = np.clip(idx - lo, 0.0, 1.0)
    row = np.arange(n_scans).reshape(n_scans, 1, 1)
    v_lo = scans[row, lo]
    v_hi = scans[row, hi]
    ct = v_lo + frac * (v_hi - v_lo)

    if area_mode:                                                   # utils.py:303-323
        span = idx[..., -1] - idx[..., 0]
        dense = span > P
        if dense.any():
            # one oversampling factor for the WHOLE call            utils.py:308
            s_area = int(math.ceil(np.max(span) / P))
            idx_a = fractional_index(s_area * P)
            nearest = np.rint(np.clip(idx_a, 0, last)).astype(np.int64)  # utils.py:318
            taps = scans[row, nearest].reshape(n_scans, -1, P, s_area)
            # sequential accumulation in the scan dtype, tap 0 first (what
            # np.mean over a non-trailing axis does at utils.py:320-322)
            acc = taps[..., 0].copy()
            for t in range(1, s_area):
                acc += taps[..., t]
            acc /= acc.dtype.type(s_area)
            ct[dense] = acc[dense]                                  # utils.py:323

    ct[outside] = padding_val                                       # utils.py:326
    r3 = centre_r[..., None]
    ct = np.clip(ct, r3 - window_depth, r3 + window_depth)          # utils.py:327
    if centered:                                                    # utils.py:328-330
        ct = ct - r3
        ct = ct / window_depth

    return np.ascontiguousarray(ct.transpose(1, 0, 2), dtype=np.float32)


def window_half_angle(scans, stride=1, fixed=False, window_width=1.66):
    """atan(0.5 * window_width / max(d, 1e-2)) in the scan dtype, [S, M]   (utils.py:274-279)."""
    scans = np.asarray(scans)
    centre_r = scans[:, ::stride]
    if not fixed:
        centre_r = np.broadcast_to(scans[-1, ::stride], centre_r.shape)
    return np.arctan(0.5 * window_width / np.maximum(centre_r, 1e-2))


def cutout_diagnostics(scans, scan_phi, stride=1, window_width=1.66,
                       num_cutout_pts=48, fixed=False, half_alpha=None, **_unused):
    """Per-sample rounding margins used by the parity tests.

    Returns dict with
      `s_area`        oversampling factor the call would use (0 = no area rows)
      `dense`         [S, M] bool, rows resampled in area mode
      `rint_margin`   [S, M, P] f64, min over a row-sample's taps of the distance
                      of the clipped area index from a .5 rounding boundary
                      (inf for non-area rows)
      `edge_margin`   [S, M, P] f64, distance of the linear index from the
                      outbound thresholds 0 and N-1
      `span_margin`   [S, M] f64, |span - P| (area-mode decision boundary)
    A GPU/oracle mismatch is only excusable where one of these margins is below
    the ~1e-4 index perturbation a 1-ulp float32 arctan difference can cause
    (SURVEY.md §7 hard part 1).
    """
    scans = np.asarray(scans)
    scan_phi = np.asarray(scan_phi)
    n_scans, n_pts = scans.shape
    P = int(num_cutout_pts)
    centre_r = scans[:, ::stride]
    if not fixed:
        centre_r = np.broadcast_to(scans[-1, ::stride], centre_r.shape)
    centre_phi = scan_phi[::stride]
    last = n_pts - 1
    half = window_half_angle(scans, stride, fixed, window_width) if half_alpha is None \
        else np.asarray(half_alpha, dtype=centre_r.dtype)
    origin, pitch = scan_phi[0], scan_phi[1] - scan_phi[0]

    def fractional_index(n):
        step = 2.0 * half / (n - 1)
        start = centre_phi - half
        return ((start[..., None] + np.arange(n) * step[..., None]) - origin) / pitch

    idx = fractional_index(P)
    span = idx[..., -1] - idx[..., 0]
    dense = span > P
    out = {
        "s_area": 0,
        "dense": dense,
        "edge_margin": np.minimum(np.abs(idx), np.abs(idx - last)),
        "span_margin": np.abs(span - P),
        "rint_margin": np.full(idx.shape, np.inf),
        "span_max": float(np.max(span)),
    }
    if dense.any():
        s_area = int(math.ceil(np.max(span) / P))
        out["s_area"] = s_area
        ia = np.clip(fractional_index(s_area * P), 0, last)
        m = np.abs((ia - np.floor(ia)) - 0.5).reshape(n_scans, -1, P, s_area).min(-1)
        out["rint_margin"] = np.where(dense[..., None], m, np.inf)
        # distance of span_max/P from an integer decides s_area itself
        q = np.max(span) / P
        out["s_area_margin"] = float(abs(q - round(q))) if abs(q - round(q)) < 0.5 else 0.5
    return out
