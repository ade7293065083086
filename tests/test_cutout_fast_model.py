"""POF_CUTOUT_FAST, restated in NumPy and bounded against the oracle on the CPU.

The scan kernel (csrc/pof_cutout.cu::cutout_scan_kernel) replaces the reference's float64 index / blend arithmetic by
  * a 32.32 fixed-point sample index (base and step rounded once, then a running 64-bit add; + 2^-24 so that the
    23-bit fraction is rounded rather than truncated),
  * blend pairs (V, D) = (v*scale, (v' - v)*scale) in float32, (V - d*scale) first, then ONE fma on the fraction,
  * float32 clip bounds.
This model evaluates exactly that arithmetic for two-tap rows and checks the error budget DESIGN.md states (5e-6 of the
output range; the reference's bar is 1e-5), so the bound is a property of the arithmetic, not only of the inputs the GPU
tests happen to use."""
import numpy as np
import pytest

from oracle import cutout as ocut
from planar_optical_flow_b200 import synth

CFG = dict(fixed=True, centered=True, window_width=1.0, window_depth=0.5, num_cutout_pts=56, padding_val=29.99, area_mode=False)


def _fma32(a, b, c):
    """float32 fused multiply-add: the product of two float32 is exact in float64."""
    return (a.astype(np.float64) * b.astype(np.float64) + c.astype(np.float64)).astype(np.float32)


def fast_model(scan, phi, window_width, window_depth, num_cutout_pts, padding_val, half_alpha, **_):
    scan = np.asarray(scan, np.float32)
    n, P = len(scan), int(num_cutout_pts)
    scale = np.float32(1.0 / window_depth)
    v0 = np.minimum(scan, np.float32(1e6))
    v1 = np.minimum(np.append(scan[1:], scan[-1]), np.float32(1e6))
    D = ((v1 - v0) * scale).astype(np.float32)
    V = (v0 * scale).astype(np.float32)
    V, D = np.append(V, V[-1]), np.append(D, D[-1])                    # entry N repeats beam N - 1
    ha = np.asarray(half_alpha, np.float32)
    step = ((np.float32(2.0) * ha) / np.float32(P - 1)).astype(np.float32)
    start = (phi - ha.astype(phi.dtype)).astype(np.float64)
    inv_pitch = 1.0 / float(phi.dtype.type(phi[1] - phi[0]))
    base = np.rint((start - float(phi[0])) * inv_pitch * 4294967296.0).astype(np.int64)
    slope = np.rint(step.astype(np.float64) * inv_pitch * 4294967296.0).astype(np.int64)
    fx = base[:, None] + 0x100 + slope[:, None] * np.arange(P, dtype=np.int64)[None, :]
    limit = ((n - 1) << 32) + 0x100
    inside = (fx >= 0) & (fx <= limit)
    beam = np.clip(fx >> 32, 0, n - 1)
    w = (((fx & 0xffffffff) >> 9).astype(np.uint32) | np.uint32(0x3f800000)).view(np.float32)          # 1.fraction
    d = scan[:, None]
    bias = (-d * scale).astype(np.float32)
    depth = np.float32(window_depth)
    lo = (((d - depth) - d) * scale).astype(np.float32)
    hi = (((d + depth) - d) * scale).astype(np.float32)
    pad = np.minimum(np.maximum(((np.float32(padding_val) - d) * scale).astype(np.float32), lo), hi)
    x = _fma32((w - np.float32(1.0)).astype(np.float32), D[beam], (V[beam] + bias).astype(np.float32))
    out = np.minimum(np.maximum(x, lo), hi)
    return np.where(inside, out, pad).astype(np.float32)


@pytest.mark.parametrize("shape,kind,seed", [("drow", "adversarial", 1), ("jrdb", "adversarial", 2), ("jrdb", "structured", 3),
                                             ("drow", "structured", 4)])
def test_fast_arithmetic_stays_within_the_parity_bar(shape, kind, seed):
    phi = synth.phi_for(shape)
    n = len(phi)
    scans = synth.adversarial_scans(1, n, seed=seed) if kind == "adversarial" else synth.structured_sequence(1, n, seed=seed, phi=phi)
    ha = ocut.window_half_angle(scans, 1, True, CFG["window_width"])
    want = ocut.scans_to_cutout(scans, phi, half_alpha=ha, **CFG)[:, 0, :]          # [N, P], float64 arithmetic of the reference
    got = fast_model(scans[0], phi, half_alpha=ha[0], **CFG)
    err = np.abs(got.astype(np.float64) - want)
    # a sample whose index sits within 2^-24 of a beam or of the scan's end may take the other side of a floor / bound test
    diag = ocut.cutout_diagnostics(scans, phi, half_alpha=ha, **CFG)
    near_edge = diag["edge_margin"].transpose(1, 0, 2)[:, 0, :] < 1e-6
    assert err[~near_edge].max() <= 5e-6, err[~near_edge].max()      # 2^-24 of the fraction x the largest neighbour difference
    print("max error %.2e, bit-equal %.1f %%" % (err[~near_edge].max(), 100 * (got == want).mean()))
