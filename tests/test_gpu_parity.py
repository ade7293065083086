"""Parity of the CUDA path (through the C ABI) against the CPU oracle and the golden
reference fixtures.  Needs a B200: run with `-m gpu`."""
import ast
import os

import numpy as np
import pytest
import torch
import torch.nn.functional as F

from oracle import cutout as ocut
from oracle import model as omodel
from oracle import nms as onms
from planar_optical_flow_b200 import ops, synth, utils
from planar_optical_flow_b200.model import SpatialDROW, _SpatialAttention
from tests.helpers import REL_TOL, assert_parity, assert_rel, f64_state_dict, rel_err

pytestmark = pytest.mark.gpu

CFG = dict(fixed=True, centered=True, window_width=1.0, window_depth=0.5,
           num_cutout_pts=56, padding_val=29.99, area_mode=True)


@pytest.fixture(scope="module", autouse=True)
def strict_fp32():
    """Parity mode: no TF32 anywhere (the 1e-5 bar is an fp32 bar)."""
    old = (torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32)
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False
    yield
    torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32 = old


def test_device_is_blackwell_and_extension_loaded():
    from planar_optical_flow_b200 import _lib

    sm, major, minor = _lib.device_info()
    assert major == 10, "libpof.so is built for sm_100a only"
    assert sm >= 100


# ------------------------------------------------------------------ cutout
def _ulp_distance(a, b):
    ia = np.ascontiguousarray(a, np.float32).view(np.int32).astype(np.int64)
    ib = np.ascontiguousarray(b, np.float32).view(np.int32).astype(np.int64)
    return np.abs(ia - ib)


def _gpu_cutout(scans, phi, kw, stride=1, half_alpha=None, fast=False, exact_pieces=False):
    s = torch.from_numpy(np.ascontiguousarray(scans, np.float32)).cuda().unsqueeze(0)
    ha = None if half_alpha is None else torch.from_numpy(np.ascontiguousarray(half_alpha, np.float32)).cuda().unsqueeze(0)
    out, ha_used = ops.cutout(s, torch.from_numpy(np.ascontiguousarray(phi)).cuda(), stride=stride,
                              half_alpha=ha, return_half_alpha=True, fast=fast, exact_pieces=exact_pieces, **kw)
    return out[0].cpu().numpy(), ha_used[0].cpu().numpy()


def _check_cutout(scans, phi, kw, stride=1, want=None, ha_ref=None):
    """Three-way cutout parity.

    The reference's arithmetic has exactly one step no other machine reproduces bit for
    bit: NumPy's SIMD float32 arctan (1-2 ulp from correctly rounded, platform specific;
    SURVEY.md section 7).  So parity is split into
      (1) reference half-angles fed INTO the kernel -> the output must match the reference
          to 1e-5 at EVERY sample (in practice bit for bit): every other operation is exact;
      (2) the kernel's own half-angles are within 2 ulp of NumPy's, and the oracle evaluated
          with the kernel's half-angles matches the kernel at every sample;
      (3) the plain default-path output is compared raw; samples may differ only where a
          <=2-ulp half-angle difference explains it (bounded count, reported).
    """
    if want is None:
        want = ocut.scans_to_cutout(scans, phi, stride=stride, **kw)
    if ha_ref is None:
        ha_ref = ocut.window_half_angle(scans, stride, kw["fixed"], kw["window_width"])
    scale = max(float(np.abs(want).max()), 1e-30)

    got_fixed, _ = _gpu_cutout(scans, phi, kw, stride, half_alpha=ha_ref)                     # (1)
    assert got_fixed.dtype == np.float32 and got_fixed.shape == want.shape
    err1 = np.abs(got_fixed.astype(np.float64) - want)
    assert err1.max() <= REL_TOL * scale, "with reference half-angles: max err %.3g" % err1.max()
    exact_frac = float((got_fixed == want).mean())
    assert exact_frac >= 0.9999, exact_frac

    got, ha_gpu = _gpu_cutout(scans, phi, kw, stride)                                         # (2)
    ulps = _ulp_distance(ha_gpu, ha_ref)
    assert ulps.max() <= 2, "device arctan differs from NumPy's by %d ulp" % ulps.max()
    want_gpu_ha = ocut.scans_to_cutout(scans, phi, stride=stride, half_alpha=ha_gpu, **kw)
    err2 = np.abs(got.astype(np.float64) - want_gpu_ha)
    assert err2.max() <= REL_TOL * scale, "with device half-angles: max err %.3g" % err2.max()

    bad = np.abs(got.astype(np.float64) - want) > REL_TOL * scale                             # (3)
    rows_differ = (ulps > 0).transpose(1, 0)[..., None]                  # [M, S, 1]
    if not kw["area_mode"] or ocut.cutout_diagnostics(scans, phi, stride=stride, **kw)["s_area"] == 0:
        assert not (bad & ~rows_differ).any(), "mismatch on a row whose half-angle is bit-equal"
    assert bad.mean() <= 2e-3, "default path: %.4f%% of samples differ by more than 1e-5" % (100 * bad.mean())

    # (4) FAST arithmetic (what the streaming engine uses): same algorithm, fixed-point index line
    # and float32 blend.  With the reference's half-angles every sample is within 1e-5 of the
    # reference, except nearest-tap flips where an area index sits within 1e-6 of a .5 boundary.
    got_fast, _ = _gpu_cutout(scans, phi, kw, stride, half_alpha=ha_ref, fast=True)
    err4 = np.abs(got_fast.astype(np.float64) - want)
    bad4 = err4 > REL_TOL * scale
    if bad4.any():
        diag = ocut.cutout_diagnostics(scans, phi, stride=stride, half_alpha=ha_ref, **kw)
        near = ((diag["rint_margin"] < 1e-6) | (diag["edge_margin"] < 1e-6)).transpose(1, 0, 2)
        assert not (bad4 & ~near).any(), "FAST arithmetic: max err %.3g" % err4[bad4 & ~near].max()
        assert bad4.sum() <= 2

    # (5) FAST on its own: the same device arctangent as EXACT (<= 2 ulp from NumPy's), and the oracle evaluated with the
    # half-angles the kernel reports matches the kernel at every sample but proven flips
    got_fast_own, ha_fast = _gpu_cutout(scans, phi, kw, stride, fast=True)
    assert _ulp_distance(ha_fast, ha_ref).max() <= 2
    want_fast = ocut.scans_to_cutout(scans, phi, stride=stride, half_alpha=ha_fast, **kw)
    bad5 = np.abs(got_fast_own.astype(np.float64) - want_fast) > REL_TOL * scale
    if bad5.any():
        diag = ocut.cutout_diagnostics(scans, phi, stride=stride, half_alpha=ha_fast, **kw)
        near = ((diag["rint_margin"] < 1e-6) | (diag["edge_margin"] < 1e-6)).transpose(1, 0, 2)
        assert not (bad5 & ~near).any(), "FAST with device half-angles: %d samples off" % int((bad5 & ~near).sum())
        assert bad5.sum() <= 2
    return int(bad.sum()), exact_frac


@pytest.mark.parametrize("shape", ["drow", "jrdb"])
@pytest.mark.parametrize("kind", ["adversarial", "structured", "edge"])
@pytest.mark.parametrize("flags", [dict(), dict(fixed=False), dict(centered=False), dict(area_mode=False),
                                   dict(window_width=1.66, window_depth=1.0, num_cutout_pts=48)])
def test_cutout_matches_oracle(shape, kind, flags):
    phi = synth.phi_for(shape)
    n = len(phi)
    scans = {"adversarial": lambda: synth.adversarial_scans(3, n, seed=11),
             "structured": lambda: synth.structured_sequence(3, n, seed=12, phi=phi),
             "edge": lambda: synth.edge_scans(n, seed=13)}[kind]()
    _check_cutout(scans, phi, dict(CFG, **flags))


@pytest.mark.parametrize("name", ["cutout_drow_adversarial", "cutout_drow_structured_lastref", "cutout_drow_edge",
                                  "cutout_jrdb_adversarial", "cutout_jrdb_structured_raw", "cutout_drow_linear48"])
def test_cutout_matches_reference_golden(golden_dir, name):
    """Against outputs of the unmodified reference, with the half-angles it used."""
    g = np.load(os.path.join(golden_dir, name + ".npz"))
    kw = dict(ast.literal_eval(str(g["kwargs"])))
    _check_cutout(g["scans"], g["phi"], kw, want=g["out"], ha_ref=g["half_alpha"])


def test_cutout_stride_ragged_and_batched():
    phi = synth.drow_phi()
    scans = synth.adversarial_scans(2, 450, seed=3)
    for stride in (2, 3, 7):
        _check_cutout(scans, phi, CFG, stride=stride)
    # odd point counts, tiny scans
    for n in (2, 5, 33, 129):
        p = synth.drow_phi(n)
        _check_cutout(synth.adversarial_scans(2, n, seed=n), p, CFG)
    # batched entry point == per-sample calls, each with its own s_area
    batch = np.stack([synth.adversarial_scans(3, 450, seed=50 + b, lo=0.3 + 0.4 * b) for b in range(5)])
    out, s_area = ops.cutout(torch.from_numpy(batch).cuda(), torch.from_numpy(phi).cuda(), return_s_area=True, **CFG)
    for b in range(5):
        assert np.array_equal(out[b].cpu().numpy(), utils.scans_to_cutout(batch[b], phi, **CFG))
        assert abs(int(s_area[b]) - ocut.cutout_diagnostics(batch[b], phi, **CFG)["s_area"]) <= \
            (ocut.cutout_diagnostics(batch[b], phi, **CFG)["s_area_margin"] < 1e-3)
    assert len(set(s_area.tolist())) > 1          # the per-sample reduction really is per sample
    empty = ops.cutout(torch.zeros(0, 1, 450, device="cuda"), torch.from_numpy(phi).cuda(), **CFG)
    assert tuple(empty.shape) == (0, 450, 1, 56)


@pytest.mark.parametrize("flags", [dict(), dict(fixed=False), dict(centered=False), dict(area_mode=False),
                                   dict(window_width=1.66, window_depth=1.0, num_cutout_pts=48), dict(num_cutout_pts=4),
                                   dict(window_depth=0.3, num_cutout_pts=64)])
@pytest.mark.parametrize("shape,kind", [("drow", "adversarial"), ("jrdb", "structured"), ("drow", "edge"), ("jrdb", "edge")])
def test_cutout_single_scan_kernel(shape, kind, flags):
    """S = 1 with FAST numerics runs cutout_scan_kernel (one CTA per scan, what the engine and the configs[1] sweep use):
    every flag combination, through the same five-way check."""
    phi = synth.phi_for(shape)
    n = len(phi)
    scans = {"adversarial": lambda: synth.adversarial_scans(1, n, seed=21),
             "structured": lambda: synth.structured_sequence(1, n, seed=22, phi=phi),
             "edge": lambda: synth.edge_scans(n, seed=23)[:1]}[kind]()
    _check_cutout(scans, phi, dict(CFG, **flags))


def test_cutout_single_scan_kernel_ragged_strided_and_close_range():
    for n in (2, 3, 5, 31, 32, 33, 129, 450):
        p = synth.drow_phi(n)
        _check_cutout(synth.adversarial_scans(1, n, seed=100 + n), p, CFG)
    phi = synth.drow_phi()
    scans = synth.adversarial_scans(1, 450, seed=5)
    for stride in (2, 3, 7):
        _check_cutout(scans, phi, CFG, stride=stride)
    # ranges below half the window width (ratio > 1: the reciprocal branch of the table arctangent), below the 1e-2
    # clamp, and a scan where EVERY row is area-resampled
    close = np.linspace(0.004, 0.6, 450, dtype=np.float32)[None]
    _check_cutout(close, phi, CFG)
    _check_cutout(np.full((1, 450), 0.2, np.float32), phi, CFG)
    # batched: per-scan s_area from the near-sensor rows == the oracle's, FAST == per-scan FAST calls
    batch = np.stack([synth.adversarial_scans(1, 450, seed=60 + b, lo=0.3 + 0.5 * b) for b in range(6)])
    out, s_area = ops.cutout(torch.from_numpy(batch).cuda(), torch.from_numpy(phi).cuda(), return_s_area=True, fast=True, **CFG)
    for b in range(6):
        diag = ocut.cutout_diagnostics(batch[b], phi, **CFG)
        assert abs(int(s_area[b]) - diag["s_area"]) <= (diag.get("s_area_margin", 1.0) < 1e-3)
        one = ops.cutout(torch.from_numpy(batch[b:b + 1]).cuda(), torch.from_numpy(phi).cuda(), fast=True, **CFG)
        assert torch.equal(one[0], out[b])
    assert len(set(s_area.tolist())) > 1


def test_cutout_torch_signature_and_full_size_properties():
    """BASELINE config 2 size (B=4096 JRDB rows): properties that need no oracle pass."""
    phi = synth.jrdb_phi()
    B = 4096
    g = torch.Generator(device="cuda").manual_seed(0)
    scans = torch.rand(B, 1, 1091, device="cuda", generator=g) * 24.7 + 0.3
    out = ops.cutout(scans, torch.from_numpy(phi).cuda(), **CFG)
    assert tuple(out.shape) == (B, 1091, 1, 56)
    assert bool(torch.isfinite(out).all()) and float(out.abs().max()) <= 1.0 + 1e-5     # centred + depth-normalised (clip bounds are float32-rounded)
    # batch-order independence and determinism
    perm = torch.randperm(B, device="cuda", generator=g)
    out_p = ops.cutout(scans[perm].contiguous(), torch.from_numpy(phi).cuda(), **CFG)
    assert torch.equal(out_p, out[perm])
    # spot samples against the oracle
    for b in (0, 1777, 4095):
        _check_cutout(scans[b].cpu().numpy(), phi, CFG)
    t = utils.scans_to_cutout_torch(scans[0], torch.from_numpy(phi).cuda(), **CFG)
    assert torch.equal(t, out[0])


# ------------------------------------------------------------------ gate
def _gate_inputs(b, n, seed, e=128, cl=(256, 14)):
    x = torch.from_numpy(synth.feature_like((b, n) + cl, seed))
    t = torch.from_numpy(synth.feature_like((b, n) + cl, seed + 1))
    ex = torch.from_numpy(synth.feature_like((b, n, e), seed + 2)) * 0.7
    et = torch.from_numpy(synth.feature_like((b, n, e), seed + 3)) * 0.7
    return x, t, ex, et


@pytest.mark.parametrize("b,n,window", [(1, 450, 11), (2, 1091, 11), (3, 37, 7), (1, 5, 11), (2, 1, 3), (1, 200, 1),
                                        (1, 161, 15), (2, 64, 5)])
def test_gate_kernel_matches_windowed_oracle(b, n, window):
    x, t, ex, et = _gate_inputs(b, n, seed=7 * n + window)
    want_out, want_ff, want_w = omodel.gate_windowed(x, t, ex, et, 0.5, window)
    out, ff, w = ops.gate_forward(x.cuda(), t.cuda(), ex.cuda(), et.cuda(), 0.5, window, want_weights=True)
    assert_rel(ff.cpu(), want_ff, what="feat_fused")
    assert_rel(out.cpu(), want_out, what="out_temp")
    assert_rel(w.cpu(), want_w, tol=1e-4, what="attn weights")
    assert float(w.sum(-1).sub(1).abs().max()) < 1e-5


def test_gate_other_shapes_alpha_and_aliasing_guard():
    # non-DR-SPAAM feature width (not a multiple of the 512-channel slice), E = 64
    x, t, ex, et = _gate_inputs(2, 77, seed=5, e=64, cl=(50, 6))
    for alpha in (0.0, 0.3, 1.0):
        want_out, want_ff, _ = omodel.gate_windowed(x, t, ex, et, alpha, 11)
        out, ff, _ = ops.gate_forward(x.cuda(), t.cuda(), ex.cuda(), et.cuda(), alpha, 11)
        assert_rel(out.cpu(), want_out)
        assert_rel(ff.cpu(), want_ff)
    xc, tc = x.cuda(), t.cuda()
    with pytest.raises(RuntimeError, match="alias"):
        ops.gate_forward(xc, tc, ex.cuda(), et.cuda(), 0.5, 11, out=tc)


def test_gate_module_matches_reference_golden(golden_dir):
    g = np.load(os.path.join(golden_dir, "gate_n24.npz"))
    seed = int(g["seed"])
    sd = omodel.randomize_bn_stats(omodel.init_state_dict(56, True, seed=seed), seed=seed + 1)
    gate = _SpatialAttention(n_pts=14, n_channel=256, alpha=0.5, window_size=11)
    gate.load_state_dict({k[5:]: v for k, v in sd.items() if k.startswith("gate.")}, strict=True)
    gate.cuda().eval()
    x = torch.from_numpy(synth.feature_like((1, 24, 256, 14), seed + 2)).cuda()
    t = torch.from_numpy(synth.feature_like((1, 24, 256, 14), seed + 3)).cuda()
    with torch.no_grad():
        out, ff = gate(x, t)
    assert_rel(ff.cpu().numpy(), g["feat_fused"], what="feat_fused vs reference")
    assert_rel(out.cpu().numpy(), g["out_temp"], what="out_temp vs reference")


def test_gate_ablation_uniform_weights():
    """Reference ablation (dr_spaam.py:166-169): with identical embeddings everywhere the
    weights are uniform over the in-range window, so out = alpha*x + (1-alpha)*window mean."""
    b, n, window = 1, 40, 11
    x, t, _, _ = _gate_inputs(b, n, seed=9)
    ex = torch.ones(b, n, 128)
    out, ff, w = ops.gate_forward(x.cuda(), t.cuda(), ex.cuda(), ex.cuda(), 0.5, window, want_weights=True)
    flat = t.reshape(b, n, -1)
    for i in (0, 3, 20, 39):
        lo, hi = max(0, i - 5), min(n - 1, i + 5)
        want = 0.5 * x.reshape(b, n, -1)[0, i] + 0.5 * flat[0, lo:hi + 1].mean(0)
        assert_rel(out.cpu().reshape(b, n, -1)[0, i], want)
    assert float((ff.cpu() - 128.0).abs().max()) == 0.0


def test_gate_backward_matches_autograd_of_oracle():
    b, n, window = 2, 53, 11
    x, t, ex, et = _gate_inputs(b, n, seed=21, cl=(32, 6))
    ins_cpu = [v.clone().double().requires_grad_(True) for v in (x, t, ex, et)]
    out_c, ff_c, _ = omodel.gate_windowed(*ins_cpu, 0.5, window)
    g_out = torch.from_numpy(synth.feature_like(tuple(out_c.shape), 31))
    g_ff = torch.from_numpy(synth.feature_like(tuple(ff_c.shape), 32)) * 0.1
    (out_c * g_out.double()).sum().add((ff_c * g_ff.double()).sum()).backward()

    ins = [v.clone().cuda().requires_grad_(True) for v in (x, t, ex, et)]
    out, ff = ops.gate(*ins, 0.5, window)
    ((out * g_out.cuda()).sum() + (ff * g_ff.cuda()).sum()).backward()
    for got, want, name in zip(ins, ins_cpu, ("g_x", "g_tmpl", "g_emb_x", "g_emb_t")):
        assert_rel(got.grad.cpu(), want.grad, tol=2e-5, what=name)


# ------------------------------------------------------------------ NMS
def _nms_case(shape, seed, scan_dtype):
    phi = synth.phi_for(shape)
    n = len(phi)
    scan = synth.structured_sequence(1, n, seed=100 + seed, phi=phi)[0].astype(scan_dtype)
    return scan, phi, synth.distinct_scores(n, seed), synth.clustered_votes(scan.astype(np.float64), phi, seed)


@pytest.mark.parametrize("shape", ["drow", "jrdb"])
@pytest.mark.parametrize("scan_dtype", [np.float32, np.float64])
def test_nms_indices_bit_exact(shape, scan_dtype):
    flips = 0
    for seed in range(8):
        scan, phi, cls, reg = _nms_case(shape, seed, scan_dtype)
        want_xy, want_cls, want_mask = onms.nms_predicted_center(scan, phi, cls, reg)
        spec = onms.nms_sweep_spec(scan, phi, cls, reg)
        xy, c, mask = utils.nms_predicted_center(scan, phi, cls, reg)
        assert mask.dtype == np.int32 and xy.dtype == want_xy.dtype and c.dtype == want_cls.dtype
        if spec["margin"] < 1e-6 and not np.array_equal(mask, want_mask):
            flips += 1            # a pair sits within float rounding of the 0.5 m threshold
            continue
        assert np.array_equal(mask, want_mask), "instance_mask differs (margin %.3g)" % spec["margin"]
        assert np.array_equal(c, want_cls)
        assert xy.shape == want_xy.shape
        # float32 scans: atan2/cos run in float32 (NumPy promotion) and differ by an ulp or two
        # between NumPy's SIMD kernels and the device; float64 scans: float64 throughout
        assert_rel(xy, want_xy, tol=1e-5 if scan_dtype == np.float32 else 1e-12)
    assert flips <= 1


@pytest.mark.parametrize("name", ["nms_drow_f32scan", "nms_drow_f64scan", "nms_jrdb_f32scan", "nms_jrdb_f64scan"])
def test_nms_matches_reference_golden(golden_dir, name):
    g = np.load(os.path.join(golden_dir, name + ".npz"))
    xy, c, mask = utils.nms_predicted_center(g["scan"], g["phi"], g["cls"], g["reg"])
    assert np.array_equal(mask, g["instance_mask"])
    assert np.array_equal(c, g["det_cls"])
    assert_rel(xy, g["det_xys"], tol=1e-5 if "f32scan" in name else 1e-12)


def test_nms_batched_order_keep_and_edges():
    phi = synth.jrdb_phi()
    cases = [_nms_case("jrdb", s, np.float32) for s in range(6)]
    scan = torch.from_numpy(np.stack([c[0] for c in cases])).cuda()
    cls = torch.from_numpy(np.stack([c[2][:, 0] for c in cases])).cuda()
    reg = torch.from_numpy(np.stack([c[3] for c in cases])).cuda()
    res = ops.nms_centers(scan, torch.from_numpy(phi).cuda(), cls, reg)
    for b, (s, p, c, r) in enumerate(cases):
        spec = onms.nms_sweep_spec(s, p, c, r)
        k = int(res["n_keep"][b])
        assert np.array_equal(res["order"][b].cpu().numpy(), spec["order"])
        assert np.array_equal(res["keep_idx"][b, :k].cpu().numpy(), spec["keep_idx"])
        assert np.array_equal(res["instance_mask"][b].cpu().numpy(), spec["instance_mask"])
    # single point, all-coincident votes, min_dist = 0 (nothing suppresses anything, not even itself)
    one = utils.nms_predicted_center(np.array([2.0], np.float32), np.array([0.1]), np.array([[0.7]], np.float32),
                                     np.zeros((1, 2), np.float32))
    assert one[0].shape == (1, 2) and one[2].tolist() == [1]
    n = 70
    phi70 = synth.drow_phi(n)
    scan70 = np.full(n, 3.0, np.float32)
    cls70 = synth.distinct_scores(n, 5)
    reg70 = synth.clustered_votes(scan70.astype(np.float64), phi70, 5, n_people=1, spread=0.0)
    for md in (0.5, 0.0, 100.0):
        want = onms.nms_predicted_center(scan70, phi70, cls70, reg70, min_dist=md)
        got = utils.nms_predicted_center(scan70, phi70, cls70, reg70, min_dist=md)
        assert np.array_equal(got[2], want[2]) and got[0].shape == want[0].shape
    # tied scores: documented rule = stable argsort reversed
    tied = np.repeat(np.array([[0.9], [0.2], [0.5]], np.float32), 10, axis=0)
    scan30, phi30 = np.full(30, 4.0, np.float32), synth.drow_phi(30)
    reg30 = np.zeros((30, 2), np.float32)
    want = onms.nms_predicted_center(scan30, phi30, tied, reg30)
    got = utils.nms_predicted_center(scan30, phi30, tied, reg30)
    assert np.array_equal(got[2], want[2]) and np.array_equal(got[1], want[1])


# ------------------------------------------------------------------ whole model
def _product_model(sd, window=11):
    m = SpatialDROW(num_scans=10, num_pts=56, alpha=0.5, window_size=window, pedestrian_only=True)
    m.load_state_dict(sd, strict=True)
    return m.cuda()


def test_streaming_model_matches_reference_golden(golden_dir):
    g = np.load(os.path.join(golden_dir, "model_stream_drow40.npz"))
    seed = int(g["seed"])
    sd = omodel.randomize_bn_stats(omodel.init_state_dict(56, True, seed=seed), seed=seed + 1)
    m = _product_model(sd).eval()
    scans, phi = g["scans"], g["phi"]
    phi_d = torch.from_numpy(phi).cuda()
    tmpl = tmpl64 = None
    sd64 = f64_state_dict(sd)
    with torch.no_grad():
        for t in range(scans.shape[1]):
            ct = ops.cutout(torch.from_numpy(scans[:, t:t + 1]).cuda(), phi_d, **CFG)
            cls, reg, tmpl, ff = m(ct, testing=True, fea_template=tmpl)
            # the float64 evaluation of the same network on the same cutouts arbitrates where two float32 pipelines differ
            ct_o = np.stack([ocut.scans_to_cutout(scans[b, t:t + 1], phi, **CFG) for b in range(scans.shape[0])])
            c64, r64, tmpl64, f64 = omodel.spatial_drow_stream(torch.from_numpy(ct_o).double(), sd64, 0.5, 11, tmpl64)
            assert_parity(cls.cpu().numpy(), g["cls_%d" % t], c64.numpy(), what="pred_cls step %d" % t)
            assert_parity(reg.cpu().numpy(), g["reg_%d" % t], r64.numpy(), what="pred_reg step %d" % t)
            assert_parity(ff.cpu().numpy(), g["feat_fused_%d" % t], f64.numpy(), what="feat_fused step %d" % t)
    assert_parity(tmpl.cpu().numpy()[:, ::8, ::16], g["template_last_sample"], tmpl64.numpy()[:, ::8, ::16], what="memory")


def test_one_instance_serves_any_point_count():
    """The reference locks a module to its first N; this one must not."""
    sd = omodel.init_state_dict(56, True, seed=2)
    m = _product_model(sd).eval()
    with torch.no_grad():
        for n in (450, 1091, 64):
            x = torch.randn(1, n, 1, 56, device="cuda")
            cls, reg, tmpl, ff = m(x, testing=True)
            assert tuple(ff.shape) == (1, n, 11) and tuple(tmpl.shape) == (1, n, 256, 14)


def test_training_branch_forward_backward_matches_oracle():
    torch.manual_seed(0)
    sd = omodel.randomize_bn_stats(omodel.init_state_dict(56, True, seed=8))
    x = torch.randn(2, 19, 4, 56)
    # oracle, train-mode BN, autograd through the dense gate
    sd_o = {k: v.clone().requires_grad_(v.is_floating_point() and "running" not in k) for k, v in sd.items()}
    cls_o, reg_o, ff_o = omodel.spatial_drow_sequence(x, sd_o, 0.5, 11, training=True)
    loss_o = cls_o.square().mean() + reg_o.square().mean() + 1e-3 * ff_o.square().mean()
    loss_o.backward()
    m = _product_model(sd).train()
    cls, reg, ff = m(x.cuda())
    loss = cls.square().mean() + reg.square().mean() + 1e-3 * ff.square().mean()
    loss.backward()
    assert_rel(cls.detach().cpu(), cls_o.detach(), tol=1e-4)
    assert_rel(ff.detach().cpu(), ff_o.detach(), tol=1e-4)
    assert abs(loss.item() - loss_o.item()) <= 1e-4 * abs(loss_o.item())
    for name, p in m.named_parameters():
        want = sd_o[name].grad
        if name.endswith(".0.bias") and name != "gate.conv.0.bias" or name == "gate.conv.0.bias":
            # a conv bias in front of a train-mode BatchNorm has zero gradient in exact arithmetic
            assert float(p.grad.abs().max()) < 1e-5 and float(want.abs().max()) < 1e-5, name
            continue
        # fp32 conv/BN backward on two devices through 3 gate steps with tiny-batch BN statistics:
        # a handful of LeakyReLU sign flips give isolated O(1e-2) differences, the direction agrees
        g, w = p.grad.cpu().flatten().double(), want.flatten().double()
        cos = float((g @ w) / (g.norm() * w.norm() + 1e-300))
        assert cos > 0.9999 and rel_err(g, w) < 5e-2, (name, cos, rel_err(g, w))
    for name, buf in m.named_buffers():
        if "running" in name:
            assert_rel(buf.cpu(), sd_o[name].detach(), tol=1e-4, what=name)


# ------------------------------------------------------------------ backbone glue kernels
@pytest.mark.parametrize("rows,C,pool", [(56 * 37, 64, 1), (56 * 37, 128, 2), (28 * 5, 256, 2), (14 * 3, 512, 1), (7, 4, 1)])
def test_act_kernel_matches_torch(rows, C, pool):
    from tests.test_backbone_layout import fake_act

    torch.manual_seed(rows + C)
    y = torch.randn(rows, C) * 3
    bias = torch.randn(C)
    for b in (bias, None):
        for slope in (0.1, 1.0):
            p, s = ops.act(y.cuda(), None if b is None else b.cuda(), pool=pool, slope=slope, want_plain=True, want_split=True)
            wp, ws = fake_act(y, b, pool=pool, slope=slope, want_plain=True, want_split=True)
            assert torch.equal(p.cpu(), wp) and torch.equal(s.cpu(), ws)
    with pytest.raises(RuntimeError):
        ops.act(y.cuda()[:, :3].contiguous(), None)


def test_conv_first_and_head_kernels_match_torch():
    from tests.test_backbone_layout import fake_conv_first, fake_head

    torch.manual_seed(3)
    cut = torch.randn(41, 56).clamp(-1, 1)
    w, b = torch.randn(64, 3), torch.randn(64)
    p, s = ops.conv_first(cut.cuda(), w.cuda(), b.cuda(), want_plain=True, want_split=True)
    wp, _ = fake_conv_first(cut, w, b, want_plain=True, want_split=True)
    assert_rel(p.cpu(), wp, tol=1e-6, what="first conv layer")
    hi = s[:, :64].cpu()
    assert torch.equal(hi, s[:, 128:].cpu()) and torch.equal(hi + s[:, 64:128].cpu(), p.cpu())
    assert int((hi.view(torch.int32) & 0x1fff).abs().max()) == 0
    # float16 parts (the engine's operand format)
    _, s16 = ops.conv_first(cut.cuda(), w.cuda(), b.cuda(), want_plain=False, want_split=True, parts=ops.SPLIT_F16)
    p16, s16b = ops.conv_first(cut.cuda(), w.cuda(), b.cuda(), want_plain=True, want_split=True, parts=ops.SPLIT_F16)
    assert s16.dtype == torch.float16 and torch.equal(s16, s16b) and torch.equal(p16, p)
    assert torch.equal(s16[:, :64], p.half())
    assert_rel((s16[:, :64].double() + s16[:, 64:].double()).cpu(), p.double().cpu(), tol=2.0 ** -21, what="float16 split")
    for L, C, H, nsig in ((7, 128, 3, 1), (3, 256, 6, 4), (1, 4, 1, 0)):
        y, bias = torch.randn(29 * L, C), torch.randn(C)
        wh, bh = torch.randn(H, C) * 0.2, torch.randn(H)
        got = ops.head(y.cuda(), bias.cuda(), 29, L, wh.cuda(), bh.cuda(), n_sigmoid=nsig)
        assert_rel(got.cpu(), fake_head(y, bias, 29, L, wh, bh, nsig), tol=2e-6, what="heads")


# ------------------------------------------------------------------ tcgen05 convolution (row N1)
def _tc_weights(w, f16):
    """[Cout, Cin, taps] -> ([taps, 2, Cout, Cin] parts, out_scale), as engine._ChannelsLastBackbone does."""
    from planar_optical_flow_b200.engine import _ChannelsLastBackbone

    holder = _ChannelsLastBackbone.__new__(_ChannelsLastBackbone)
    holder.f16 = f16
    return holder._tc_weight(w)


@pytest.mark.parametrize("f16", [True, False])
@pytest.mark.parametrize("M,LA,Cin,Cout,taps,pad,pool", [(5, 56, 64, 64, 3, 1, 1), (9, 28, 128, 256, 3, 1, 2), (40, 14, 256, 512, 3, 1, 2),
                                                         (61, 7, 512, 256, 3, 1, 1), (130, 14, 256, 128, 14, 0, 1)])
def test_conv_tc_matches_fp64(f16, M, LA, Cin, Cout, taps, pad, pool):
    """pof_conv_tc_fwd / pof_conv_tc_f16_fwd against an fp64 convolution: fp32-level accuracy from split tensor-core products."""
    g = torch.Generator().manual_seed(M + Cin)
    x = torch.randn(M, LA, Cin, generator=g).abs() * torch.rand(M, LA, Cin, generator=g)
    x = torch.where(torch.rand(M, LA, Cin, generator=g) < 0.3, -0.1 * x, x).cuda()
    w = (torch.randn(Cout, Cin, taps, generator=g) * (2.0 / (Cin * taps)) ** 0.5).cuda()
    b = (torch.randn(Cout, generator=g) * 0.1).cuda()
    Lout = LA if pad else LA - taps + 1
    parts = ops.SPLIT_F16 if f16 else 2
    _, a = ops.act(x.view(M * LA, Cin), None, pool=1, slope=1.0, want_plain=False, want_split=True, parts=parts)
    assert a.dtype == (torch.float16 if f16 else torch.float32)
    assert_rel((a[:, :Cin].double() + a[:, Cin:].double()).cpu(), x.view(M * LA, Cin).double().cpu(), tol=2.0 ** -21, what="operand split")
    ws, out_scale = _tc_weights(w, f16)
    plain, split = ops.conv_tc(a, ws, b, M, LA, Lout, taps, pad, pool=pool, slope=0.1, want_plain=True, want_split=True,
                               out_scale=out_scale)
    torch.cuda.synchronize()
    assert ops.conv_tc_status(x.device) == 0
    y = F.conv1d(x.permute(0, 2, 1).double(), w.double(), b.double(), padding=pad)
    if pool == 2:
        y = F.max_pool1d(y, 2)
    want = torch.where(y > 0, y, y * 0.1).permute(0, 2, 1).reshape(-1, Cout)
    assert_rel(plain.double().cpu(), want.cpu(), tol=1.5e-6, what="convolution")
    assert_rel((split[:, :Cout].double() + split[:, Cout:].double()).cpu(), plain.double().cpu(), tol=2.0 ** -21, what="output split")


def test_conv_tc_f16_reports_activations_beyond_float16():
    """An output the float16 split cannot hold must raise the device status, not pass silently."""
    M, LA, C = 4, 14, 64
    x = torch.full((M * LA, C), 40.0).cuda()
    w = torch.full((C, C, 3), 30.0).cuda()                      # sums of 64 * 3 * 1200 = 230,400 > 65504
    _, a = ops.act(x, None, pool=1, slope=1.0, want_plain=False, want_split=True, parts=ops.SPLIT_F16)
    ws, out_scale = _tc_weights(w, True)
    try:
        ops.conv_tc(a, ws, None, M, LA, LA, 3, 1, want_plain=False, want_split=True, out_scale=out_scale)
        torch.cuda.synchronize()
        assert ops.conv_tc_status(x.device) == 16
        plain, _ = ops.conv_tc(a, ws, None, M, LA, LA, 3, 1, want_plain=True, want_split=False, out_scale=out_scale,
                               chain_channels=0x40000)               # POF_CONV_TC_NO_DEBIAS: the bare sum
        assert float(plain.max()) == 40.0 * 30.0 * 64 * 3          # the fp32 output itself is exact
    finally:
        ops._conv_tc_status[x.device].zero_()


def test_gate_emits_the_operand_split_in_the_same_pass():
    """pof_spaam_gate_fwd's optional second output == pof_act_fwd's float16 split of the new memory, bit for bit."""
    torch.manual_seed(4)
    b, n, L, C, E = 2, 37, 14, 256, 128
    x, t = torch.randn(b, n, L, C).cuda(), torch.randn(b, n, L, C).cuda()
    ex, et = torch.randn(b, n, E).cuda() * 0.2, torch.randn(b, n, E).cuda() * 0.2
    plain_only, ff0, _ = ops.gate_forward(x, t, ex, et, 0.5, 11)
    split = torch.empty((b * n * L, 2 * C), dtype=torch.float16, device="cuda")
    status = ops.new_status(x.device)
    out, ff, _ = ops.gate_forward(x, t, ex, et, 0.5, 11, split_out=split, split_channels=C, status=status)
    assert torch.equal(out, plain_only) and torch.equal(ff, ff0)
    _, want = ops.act(out.view(b * n * L, C), None, pool=1, slope=1.0, want_plain=False, want_split=True, parts=ops.SPLIT_F16)
    assert torch.equal(split, want)
    assert ops.read_status(status) == 0
    # a memory value beyond binary16 is reported, not silently turned into inf
    x[1, 3, 2, 5] = 3.0e5
    ops.gate_forward(x, t, ex, et, 0.5, 11, split_out=split, split_channels=C, status=status)
    assert ops.read_status(status) == 16 and ops.read_status(status) == 0          # reading clears it


def test_head_writes_scores_and_votes_to_separate_tensors():
    torch.manual_seed(6)
    M, L, C = 53, 7, 128
    y, wh, bh = torch.randn(M * L, C).cuda(), (torch.randn(3, C) * 0.2).cuda(), torch.randn(3).cuda()
    both = ops.head(y, None, M, L, wh, bh, n_sigmoid=1, slope=1.0)
    cls, reg = torch.empty(M, 1, device="cuda"), torch.empty(M, 2, device="cuda")
    ops.head(y, None, M, L, wh, bh, n_sigmoid=1, slope=1.0, out=cls, out_rest=reg)
    assert torch.equal(cls, both[:, :1]) and torch.equal(reg, both[:, 1:])


def test_detector_status_is_its_own_and_clears():
    """A float16-range overflow in one detector raises in ITS step() and leaves other detectors (and its own next step) clean."""
    from planar_optical_flow_b200.engine import StreamingDetector

    n = 64
    phi = synth.drow_phi(n)
    sd = omodel.randomize_bn_stats(omodel.init_state_dict(56, True, seed=3))
    hot = {k: v.clone() for k, v in sd.items()}
    hot["conv_block_2.2.1.weight"] = hot["conv_block_2.2.1.weight"] * 3.0e5        # block-2 features far beyond 65504
    scans = synth.structured_sequence(2, n, seed=1, phi=phi)
    bad = StreamingDetector(_product_model(hot), phi, CFG, 1)
    good = StreamingDetector(_product_model(sd), phi, CFG, 1)
    with pytest.raises(RuntimeError, match="float16 range"):
        bad.step(scans[0:1])
    good.step(scans[0:1])
    good.check()
    bad.check()                                                                   # the failed step was reported once
    ok32 = StreamingDetector(_product_model(hot), phi, CFG, 1, precision="fp32-tf32")    # what the message recommends
    ok32.step(scans[0:1])
    ok32.check()


# ------------------------------------------------------------------ legacy preprocessing (row N4)
@pytest.mark.parametrize("shape", ["drow", "jrdb"])
def test_cutout_original_and_polar_grid_match_reference_golden(golden_dir, shape):
    from oracle import cutout_legacy as ol

    g = np.load(os.path.join(golden_dir, "cutout_original_%s.npz" % shape))
    incre = g["incre"][()]
    kw = dict(fixed=True, centered=True, window_width=1.66, window_depth=1.0, num_cutout_pts=48, padding_val=29.99)
    cases = ((g["scans"], g["out"], kw), (g["adv"], g["out_adv"], kw),
             (g["scans"], g["out_lastref"], dict(kw, fixed=False, centered=False, num_cutout_pts=56)))
    for scans, want, k in cases:
        got = utils.scans_to_cutout_original(scans, incre, **k)
        assert got.dtype == np.float32 and got.shape == want.shape
        bad = np.abs(got.astype(np.float64) - want) > REL_TOL * np.abs(want).max()          # [N, S, P]
        # a window end that sits within 1e-5 beams of a rounding boundary may move with the device arctangent
        near = (ol.window_margins(scans, incre, k["fixed"], k["window_width"]) < 1e-5).T[..., None]
        assert not (bad & ~near).any(), "legacy cutout: %d samples off" % int((bad & ~near).sum())
        assert bad.any(axis=2).sum() <= 2
    assert np.array_equal(utils.scans_to_polar_grid(g["scans"][:, ::16]), g["polar"])
    assert np.array_equal(utils.scans_to_polar_grid(g["adv"][:, ::16], 0.5, 20.0, 0.5, 0.0, False), g["polar_raw"])


def test_cutout_original_batched_and_f32_pitch():
    from oracle import cutout_legacy as ol

    phi32 = synth.phi_for("jrdb")                                              # float32 angles: float32 beam pitch
    assert phi32.dtype == np.float32
    scans = np.stack([synth.structured_sequence(2, len(phi32), seed=k, phi=phi32) for k in range(3)])      # [B, S, N]
    incre = phi32[1] - phi32[0]
    got = ops.cutout_original(torch.from_numpy(scans).cuda(), float(incre), angle_incre_is_f32=True, num_cutout_pts=32).cpu().numpy()
    for b in range(3):
        want = ol.scans_to_cutout_original(scans[b], incre, num_cutout_pts=32)
        bad = np.abs(got[b] - want) > REL_TOL * np.abs(want).max()
        near = (ol.window_margins(scans[b], incre, True, 1.66) < 1e-4).T[..., None]
        assert not (bad & ~near).any()


# ------------------------------------------------------------------ scan-pair flow prototype (row N3)
@pytest.mark.parametrize("b,c,n,k,d", [(2, 256, 57, 3, 5), (1, 8, 7, 3, 5), (3, 40, 137, 5, 2), (1, 3, 1, 3, 5), (2, 300, 64, 3, 15)])
def test_patch_corr_forward_backward(b, c, n, k, d):
    from oracle import prototype as oproto

    torch.manual_seed(b + c + n)
    f1 = torch.randn(b, c, n, dtype=torch.float64, requires_grad=True)
    f2 = torch.randn(b, c, n, dtype=torch.float64, requires_grad=True)
    want = oproto.fusion_dense(f1, f2, k, d)
    w = torch.randn_like(want)
    (want * w).sum().backward()
    g1 = f1.detach().float().cuda().requires_grad_(True)
    g2 = f2.detach().float().cuda().requires_grad_(True)
    got = ops.patch_corr(g1, g2, kernel_size=k, max_displacement=d)
    (got * w.float().cuda()).sum().backward()
    assert_rel(got.detach().cpu(), want.detach(), tol=2e-6, what="patch correlation")
    assert_rel(g1.grad.cpu(), f1.grad, tol=2e-6, what="grad feat1")
    assert_rel(g2.grad.cpu(), f2.grad, tol=2e-6, what="grad feat2")
    # deterministic backward: bit-equal on a second run
    h1 = g1.detach().clone().requires_grad_(True)
    h2 = g2.detach().clone().requires_grad_(True)
    (ops.patch_corr(h1, h2, kernel_size=k, max_displacement=d) * w.float().cuda()).sum().backward()
    assert torch.equal(h1.grad, g1.grad) and torch.equal(h2.grad, g2.grad)


def test_prototype_matches_reference_golden_and_trains(golden_dir):
    from oracle import prototype as oproto
    from planar_optical_flow_b200.model.prototype import Prototype

    g = np.load(os.path.join(golden_dir, "prototype_drow450.npz"))
    sd = oproto.init_state_dict(2, 5, seed=int(g["seed"]))
    m = Prototype(in_channel=2, max_displacement=5)
    m.load_state_dict(sd, strict=True)                                    # the reference's checkpoint layout
    m = m.cuda().eval()
    with torch.no_grad():
        flow = m(torch.from_numpy(g["scan1"]).cuda(), torch.from_numpy(g["scan2"]).cuda())
        fused = m._fusion(torch.from_numpy(g["f1"]).cuda(), torch.from_numpy(g["f2"]).cuda())
    assert_rel(fused.cpu().numpy(), g["fused"], tol=2e-6, what="fusion vs reference")
    assert_rel(flow.cpu().numpy(), g["flow"], tol=2e-5, what="flow vs reference")
    # one training step against the autograd of the oracle (train-mode BN)
    m.train()
    s1, s2 = torch.from_numpy(g["scan1"][:2]), torch.from_numpy(g["scan2"][:2])
    tgt = torch.from_numpy(g["flow"][:2]) * 0.5
    loss, _ = m.loss_fn(m(s1.cuda(), s2.cuda()), tgt.cuda())
    loss.backward()
    sd_o = {k: v.clone().requires_grad_(v.is_floating_point() and "running" not in k) for k, v in sd.items()}
    loss_o = torch.mean(torch.mean(torch.norm(oproto.prototype_forward(s1, s2, sd_o, training=True) - tgt, dim=-1), dim=1))
    loss_o.backward()
    assert abs(loss.item() - loss_o.item()) <= 1e-4 * abs(loss_o.item())
    for name, p in m.named_parameters():
        if name.endswith(".0.bias"):
            continue                        # a conv bias in front of train-mode BN has zero gradient in exact arithmetic
        gq, wq = p.grad.cpu().flatten().double(), sd_o[name].grad.flatten().double()
        cos = float((gq @ wq) / (gq.norm() * wq.norm() + 1e-300))
        assert cos > 0.999, (name, cos)


# ------------------------------------------------------------------ streaming engine
# tf32x3 on cuDNN: the split products are exact, but the tensor cores ACCUMULATE with truncation, a bias of
# ~5e-8 per 8-wide k-step that adds up coherently over the 4608-deep reductions (measured 8e-5 on the
# votes, profiles/r1_precision_modes.txt) - 10x closer than plain TF32, not the 1e-5 parity bar.
@pytest.mark.parametrize("precision,tol", [("fp32", REL_TOL), ("fp32-tf32", REL_TOL), ("fp32-simt", REL_TOL), ("tf32x3", 4e-4), ("tf32", 2e-2)])
def test_streaming_engine_matches_oracle_stream(precision, tol):
    """StreamingDetector (BN folded, memory resident, NMS on device) against the oracle's
    cutout -> SpatialDROW(testing=True) -> sigmoid -> NMS loop, 3 steps, 3 sequences.

    The network is compared on IDENTICAL cutouts: the oracle is fed the cutout kernel's output for the same ranges (the
    kernel's parity with the reference's NumPy cutout - bit-equal samples given the half-angles, half-angles within 2 ulp
    of NumPy's platform-specific float32 arctan - is the subject of the cutout tests above; a 1-ulp half-angle moves a
    sample next to a range discontinuity by up to ~1e-5 of the window depth, which would otherwise be charged to the network).
    The whole chain from the oracle's own NumPy cutouts is checked as well for every sequence whose device cutouts have
    matched NumPy's to 1e-5 at every sample so far (a nearest-beam flip of an area-resampled sample changes that sample
    by up to the clip range, and everything downstream of it)."""
    from planar_optical_flow_b200.engine import StreamingDetector

    n, b, steps = 90, 3, 3
    phi = synth.drow_phi(n)
    scans = np.stack([synth.structured_sequence(steps, n, seed=70 + k, phi=phi) for k in range(b)], axis=1)  # [T,B,N]
    sd = omodel.randomize_bn_stats(omodel.init_state_dict(56, True, seed=9))
    det = StreamingDetector(_product_model(sd), phi, CFG, b, precision=precision, seq_chunk=2)
    tmpl, tmpl64, tmpl_np = [None] * b, [None] * b, [None] * b
    same_input, n_e2e = [True] * b, 0
    sd64 = f64_state_dict(sd)
    phi_d = torch.from_numpy(phi).cuda()
    for t in range(steps):
        host = det.step(scans[t])
        ct_dev = ops.cutout(torch.from_numpy(scans[t]).cuda().unsqueeze(1), phi_d, **CFG).cpu()       # [B, N, 1, P]: what the engine saw
        for k in range(b):
            ct = ct_dev[k].numpy()
            with torch.no_grad():
                cls, reg, tmpl[k], ff = omodel.spatial_drow_stream(torch.from_numpy(ct)[None], sd, 0.5, 11, tmpl[k])
                c64, r64, tmpl64[k], f64 = omodel.spatial_drow_stream(torch.from_numpy(ct)[None].double(), sd64, 0.5, 11, tmpl64[k])
                ct_np = ocut.scans_to_cutout(scans[t, k][None], phi, **CFG)                            # the reference's own cutout
                c_np, r_np, tmpl_np[k], _ = omodel.spatial_drow_stream(torch.from_numpy(ct_np)[None], sd, 0.5, 11, tmpl_np[k])
            same_input[k] = same_input[k] and float(np.abs(ct - ct_np).max()) <= REL_TOL
            if tol <= REL_TOL and same_input[k]:          # no nearest-beam flip from a 1-ulp half-angle so far: the whole chain agrees too
                n_e2e += 1
                assert_parity(c_full(det, k), torch.sigmoid(c_np[0]).numpy(), torch.sigmoid(c64[0]).numpy(), tol=2 * tol, what="end-to-end scores step %d" % t)
                assert_parity(r_full(det, k), r_np[0].numpy(), r64[0].numpy(), tol=2 * tol, what="end-to-end votes step %d" % t)
                assert_parity(det.template[k].cpu(), tmpl_np[k][0], tmpl64[k][0], tol=2 * tol, what="end-to-end memory step %d" % t)
            conf = torch.sigmoid(cls[0]).numpy()
            assert_parity(det.template[k].cpu(), tmpl[k][0], tmpl64[k][0], tol=tol, what="memory step %d" % t)
            assert_parity(c_full(det, k), conf, torch.sigmoid(c64[0]).numpy(), tol=tol, what="scores step %d" % t)
            assert_parity(r_full(det, k), reg[0].numpy(), r64[0].numpy(), tol=tol, what="votes step %d" % t)
            assert_parity(det._last["feat_fused"][k].cpu(), ff[0], f64[0], tol=tol, what="similarities step %d" % t)
            want = onms.nms_sweep_spec(scans[t, k], phi, conf, reg[0].numpy())
            xy, c, mask = det.detections(host, k)
            # the engine's NMS consumes ITS OWN scores; indices agree whenever no score pair or
            # distance sits within the fp32 noise between the two pipelines
            mine = onms.nms_sweep_spec(scans[t, k], phi, c_full(det, k), r_full(det, k))
            assert np.array_equal(mask, mine["instance_mask"])
            assert np.array_equal(host["keep_idx"][k, :len(xy)], mine["keep_idx"])
            if np.array_equal(mine["order"], want["order"]) and want["margin"] > 1e-4:
                assert np.array_equal(mask, want["instance_mask"])
    det.check()
    assert det.steps_done == steps


def c_full(det, k):
    return det._last["pred_cls"][k].cpu().numpy().reshape(-1, 1)


def r_full(det, k):
    return det._last["pred_reg"][k].cpu().numpy()


def test_cuda_graph_replay_equals_eager_steps():
    """StreamingDetector(cuda_graph=True): from the third step on every step is one graph replay; results bit-equal to eager."""
    from planar_optical_flow_b200.engine import StreamingDetector

    n, b, steps = 90, 2, 7
    phi = synth.drow_phi(n)
    scans = np.stack([synth.structured_sequence(steps, n, seed=20 + k, phi=phi) for k in range(b)], axis=1)
    sd = omodel.randomize_bn_stats(omodel.init_state_dict(56, True, seed=4))
    eager = StreamingDetector(_product_model(sd), phi, CFG, b)
    graph = StreamingDetector(_product_model(sd), phi, CFG, b, cuda_graph=True)
    for t in range(steps):
        he = {k: v.copy() for k, v in eager.step(scans[t]).items()}
        hg = graph.step(scans[t])
        assert np.array_equal(he["n_keep"], hg["n_keep"]) and np.array_equal(he["instance_mask"], hg["instance_mask"]), t
        for q in range(b):                                  # rows beyond n_keep are never written
            kq = int(he["n_keep"][q])
            for k in ("keep_idx", "det_xy", "det_cls"):
                assert np.array_equal(he[k][q, :kq], hg[k][q, :kq]), (t, k)
        assert torch.equal(eager.template, graph.template)
        assert torch.equal(eager._last["pred_reg"], graph._last["pred_reg"])
    assert len(graph._graphs) == 2 and graph.kernel_launches == eager.kernel_launches
    graph.check()


@pytest.mark.parametrize("fast", [False, True])
@pytest.mark.parametrize("fixed", [True, False])
def test_cutout_span_reduction_paths_agree(fast, fixed):
    """The per-scan span reduction (nearest rows only, atomic max per sample; what a plain call uses) and the per-sample
    one that also reports `s_area` give bit-equal cutouts: structured, adversarial and edge ranges, S = 1 and S = 4."""
    phi = synth.phi_for("jrdb")
    n = len(phi)
    kw = dict(CFG, fixed=fixed)
    phi_d = torch.from_numpy(phi).cuda()
    for S in (1, 4):
        rows = [synth.structured_sequence(S, n, seed=3, phi=phi), synth.adversarial_scans(S, n, seed=4),
                np.full((S, n), 0.004, np.float32), np.full((S, n), 29.99, np.float32)]
        rows[2][:, ::7] = 3.0
        scans = torch.from_numpy(np.stack(rows)).cuda()                     # [B = 4, S, N]
        plain = ops.cutout(scans, phi_d, fast=fast, **kw)
        full, s_area = ops.cutout(scans, phi_d, fast=fast, return_s_area=True, **kw)
        assert torch.equal(plain, full), (S, int((plain != full).sum()))
        assert int(s_area.max()) >= 4 and int(s_area.min()) >= 0


@pytest.mark.parametrize("flags", [dict(), dict(fixed=False), dict(centered=False), dict(area_mode=False), dict(window_depth=0.3),
                                   dict(window_width=1.66, window_depth=1.0, num_cutout_pts=48), dict(num_cutout_pts=4),
                                   dict(padding_val=float("inf"))])
@pytest.mark.parametrize("phi_dtype", [np.float32, np.float64])
def test_cutout_exact_kernels_are_bit_equal(flags, phi_dtype):
    """The EXACT arithmetic has two kernels: one CTA per scan (what a call runs) and the first piece-per-thread one
    (`exact_pieces`, the fallback for scans that do not fit shared memory).  Same operations in the same order, so the same
    bits - NaN patterns included - on structured, adversarial, edge, all-area, non-finite and out-of-scan inputs, S = 1 and
    S = 3, strided and ragged."""
    kw = dict(CFG, **flags)
    for shape in ("jrdb", "drow"):
        phi = synth.phi_for(shape).astype(phi_dtype)
        n = len(phi)
        phi_d = torch.from_numpy(phi).cuda()
        for S in (1, 3):
            rows = [synth.structured_sequence(S, n, seed=3, phi=phi), synth.adversarial_scans(S, n, seed=4),
                    synth.edge_scans(n, seed=5)[:S], np.full((S, n), 0.2, np.float32), np.full((S, n), 29.99, np.float32),
                    np.linspace(0.004, 0.6, n, dtype=np.float32)[None].repeat(S, 0)]
            rows[3][:, ::7] = 3.0
            bad = synth.adversarial_scans(S, n, seed=6)
            bad[:, 5] = np.inf
            bad[:, 17:19] = np.inf
            bad[:, 40] = np.nan
            bad[:, n - 1] = np.inf
            rows.append(bad)
            scans = torch.from_numpy(np.stack(rows)).cuda()
            for stride in (1, 3):
                a, sa = ops.cutout(scans, phi_d, stride=stride, return_s_area=True, **kw)
                b, sb = ops.cutout(scans, phi_d, stride=stride, return_s_area=True, exact_pieces=True, **kw)
                assert torch.equal(sa, sb)
                assert torch.equal(a.view(torch.int32), b.view(torch.int32)), (shape, S, stride, int((a.view(torch.int32) != b.view(torch.int32)).sum()))
                a2 = ops.cutout(scans, phi_d, stride=stride, **kw)               # the plain call (per-scan span reduction)
                assert torch.equal(a2.view(torch.int32), a.view(torch.int32))
    for n in (2, 3, 5, 33, 129):
        phi = synth.drow_phi(n).astype(phi_dtype)
        scans = torch.from_numpy(synth.adversarial_scans(2, n, seed=n)[None]).cuda()
        a = ops.cutout(scans, torch.from_numpy(phi).cuda(), **kw)
        b = ops.cutout(scans, torch.from_numpy(phi).cuda(), exact_pieces=True, **kw)
        assert torch.equal(a.view(torch.int32), b.view(torch.int32)), n


@pytest.mark.parametrize("flags", [dict(), dict(fixed=False), dict(centered=False), dict(window_depth=0.3), dict(area_mode=False)])
@pytest.mark.parametrize("S", [1, 3])
def test_cutout_infinite_ranges_match_the_reference(S, flags):
    """`inf` (no return) is inside the reference's domain: a blend across an infinite beam is inf or NaN (inf - inf), a row whose
    own range is infinite has NaN clip bounds after centring, and np.clip propagates every NaN (utils.py:300, 327-330).  Both EXACT
    kernels reproduce the NaN and inf masks of the oracle (itself equal to the unmodified reference on these inputs) and the
    finite samples to 1e-5."""
    kw = dict(CFG, **flags)
    for shape in ("jrdb", "drow"):
        phi = synth.phi_for(shape)
        n = len(phi)
        scans = synth.adversarial_scans(S, n, seed=6)
        scans[:, 5] = np.inf
        scans[:, 17:19] = np.inf
        scans[:, n // 2] = np.inf
        scans[:, n - 1] = np.inf
        with np.errstate(all="ignore"):
            want = ocut.scans_to_cutout(scans, phi, **kw)
            ha_ref = ocut.window_half_angle(scans, 1, kw["fixed"], kw["window_width"])
        assert np.isnan(want).sum() > 50
        finite = np.isfinite(want)
        scale = float(np.abs(want[finite]).max())
        for pieces in (False, True):
            got, _ = _gpu_cutout(scans, phi, kw, half_alpha=ha_ref, exact_pieces=pieces)
            assert np.array_equal(np.isnan(got), np.isnan(want)), (shape, pieces, int((np.isnan(got) != np.isnan(want)).sum()))
            assert np.array_equal(np.isinf(got), np.isinf(want)) and np.array_equal(got[np.isinf(want)], want[np.isinf(want)])
            assert np.abs(got[finite].astype(np.float64) - want[finite]).max() <= REL_TOL * scale
            assert (got[finite] == want[finite]).mean() >= 0.9999


@pytest.mark.parametrize("n", [6000, 9000])
def test_cutout_long_scans_take_the_fallback_kernels(n):
    """Scans too long for the scan kernels' shared-memory staging (16 B per beam for EXACT, 8192 beams for any staging) fall
    back to the piece kernels: same five-way check against the oracle, S = 1 and S = 2."""
    phi = np.linspace(-np.pi, np.pi, n).astype(np.float32)
    for S in (1, 2):
        scans = synth.adversarial_scans(S, n, seed=n + S, lo=2.0)
        scans[:, ::97] = 1.5                                              # a few close returns: area rows with a large s_area
        _check_cutout(scans, phi, CFG)


def test_cutout_exact_kernel_at_full_size_equals_pieces_kernel():
    """BASELINE configs[1] size (4096 JRDB scans): both EXACT kernels, every sample."""
    phi = torch.from_numpy(synth.jrdb_phi()).cuda()
    g = torch.Generator(device="cuda").manual_seed(1)
    scans = torch.rand(4096, 1, 1091, device="cuda", generator=g) * 24.7 + 0.3
    scans[::3] = scans[::3] * 0.1 + 0.2                                       # a third of the scans close to the sensor: area rows
    a = ops.cutout(scans, phi, **CFG)
    b = ops.cutout(scans, phi, exact_pieces=True, **CFG)
    assert torch.equal(a, b), int((a != b).sum())


@pytest.mark.parametrize("M,L,C,pool", [(37, 56, 64, 1), (37, 56, 128, 2), (11, 28, 256, 2), (9, 14, 512, 2), (300, 1, 128, 1)])
def test_bn_act_pool_forward_backward_matches_float64_autograd(M, L, C, pool):
    """libpof's training-mode BatchNorm + LeakyReLU (+ pool) operator against torch autograd in float64."""
    g = torch.Generator().manual_seed(M + C)
    y = (torch.randn(M, C, 1, L, generator=g) * 1.5 + 0.3)
    gamma, beta = torch.rand(C, generator=g) + 0.5, torch.randn(C, generator=g) * 0.2
    rm, rv = torch.randn(C, generator=g) * 0.1, torch.rand(C, generator=g) + 0.5
    w = torch.randn(M, C, 1, L // pool, generator=g)
    # float64 reference
    y64 = y.double().requires_grad_(True)
    g64, b64 = gamma.double().requires_grad_(True), beta.double().requires_grad_(True)
    rm64, rv64 = rm.double().clone(), rv.double().clone()
    z64 = F.leaky_relu(F.batch_norm(y64, rm64, rv64, g64, b64, True, 0.1, 1e-5), 0.1)
    if pool == 2:
        z64 = F.max_pool2d(z64, kernel_size=(1, 2))
    (z64 * w.double()).sum().backward()
    # the operator
    yc = y.cuda().contiguous(memory_format=torch.channels_last).requires_grad_(True)
    gc, bc = gamma.cuda().requires_grad_(True), beta.cuda().requires_grad_(True)
    rmc, rvc = rm.cuda().clone(), rv.cuda().clone()
    z = ops.bn_act_pool(yc, gc, bc, rmc, rvc, momentum=0.1, eps=1e-5, slope=0.1, pool=pool)
    assert z.is_contiguous(memory_format=torch.channels_last) and tuple(z.shape) == (M, C, 1, L // pool)
    (z * w.cuda()).sum().backward()
    assert_rel(z.detach().cpu(), z64.detach(), tol=2e-6, what="forward")
    assert_rel(rmc.cpu(), rm64, tol=2e-6, what="running mean")
    assert_rel(rvc.cpu(), rv64, tol=2e-6, what="running variance")
    assert_rel(yc.grad.cpu(), y64.grad, tol=5e-6, what="grad input")
    assert_rel(gc.grad.cpu(), g64.grad, tol=5e-6, what="grad gamma")
    assert_rel(bc.grad.cpu(), b64.grad, tol=5e-6, what="grad beta")


def test_fused_training_layers_equal_the_cudnn_path():
    """SpatialDROW's training branch with the fused operator vs the same module on cuDNN batch-norm + PyTorch's LeakyReLU
    and max-pool kernels: outputs, running statistics and every gradient."""
    from planar_optical_flow_b200.model.dr_spaam import _ConvBnAct

    torch.manual_seed(3)
    sd = omodel.randomize_bn_stats(omodel.init_state_dict(56, True, seed=12))
    x = torch.randn(2, 23, 3, 56).cuda()
    res = {}
    try:
        for fused in (True, False):
            _ConvBnAct.fused = fused
            m = _product_model(sd).train()
            cls, reg, ff = m(x)
            (cls.square().mean() + reg.square().mean() + 1e-3 * ff.square().mean()).backward()
            res[fused] = (cls.detach(), reg.detach(), ff.detach(), {k: p.grad.clone() for k, p in m.named_parameters()},
                          {k: b.clone() for k, b in m.named_buffers() if "running" in k})
    finally:
        _ConvBnAct.fused = True
    a, b = res[True], res[False]
    for i, name in enumerate(("cls", "reg", "feat_fused")):
        assert_rel(a[i].cpu(), b[i].cpu(), tol=1e-4, what=name)          # two TF32-free fp32 paths through 11 train-mode BN layers
    for k in b[4]:
        assert_rel(a[4][k].cpu(), b[4][k].cpu(), tol=1e-4, what=k)
    for k in b[3]:
        if k.endswith(".0.bias") or k == "gate.conv.0.bias":
            continue                                                      # zero in exact arithmetic (a bias in front of train-mode BN)
        # where the two values of a pooled pair agree to the last bits, the two paths may route the gradient to different
        # rows (and a pre-activation next to zero may take the other slope): isolated elements differ, the direction agrees
        ga, gb = a[3][k].flatten().double().cpu(), b[3][k].flatten().double().cpu()
        cos = float((ga @ gb) / (ga.norm() * gb.norm() + 1e-300))
        assert cos > 0.9999 and rel_err(ga, gb) < 0.2, (k, cos, rel_err(ga, gb))


def test_first_layer_training_conv_and_weight_gradient():
    """ops.conv_first_train (the 1 -> 64 layer of the training branch: forward kernel + pof_conv_first_wgrad) vs autograd in float64."""
    g = torch.Generator().manual_seed(8)
    M, P, C = 211, 56, 64
    x = torch.randn(M, P, generator=g).clamp(-1, 1)
    w = torch.randn(C, 1, 3, generator=g) * 0.5
    gy = torch.randn(M, C, 1, P, generator=g)
    w64 = w.double().requires_grad_(True)
    y64 = F.conv1d(x.double().view(M, 1, P), w64, None, padding=1)                     # [M, C, P]
    (y64 * gy.double().view(M, C, P)).sum().backward()
    wc = w.cuda().requires_grad_(True)
    y = ops.conv_first_train(x.cuda(), wc)
    assert tuple(y.shape) == (M, C, 1, P) and y.is_contiguous(memory_format=torch.channels_last)
    (y * gy.cuda()).sum().backward()
    assert_rel(y.detach().cpu().view(M, C, P), y64.detach(), tol=1e-6, what="first layer forward")
    assert_rel(wc.grad.cpu(), w64.grad, tol=2e-6, what="first layer weight gradient")


def test_bn_act_pool_groups_equal_separate_calls():
    """groups = G on G stacked blocks == G separate calls in order: outputs, gradients and the running statistics after the
    G sequential updates (the reference runs a layer once per scan, dr_spaam.py:264-273)."""
    g = torch.Generator().manual_seed(21)
    G, Mg, L, C, pool = 5, 23, 28, 128, 2
    y = (torch.randn(G * Mg, C, 1, L, generator=g) * torch.linspace(0.5, 2.0, G).repeat_interleave(Mg)[:, None, None, None]).cuda()
    y = y.contiguous(memory_format=torch.channels_last)
    gamma, beta = (torch.rand(C, generator=g) + 0.5).cuda(), (torch.randn(C, generator=g) * 0.2).cuda()
    w = torch.randn(G * Mg, C, 1, L // pool, generator=g).cuda()
    rm0, rv0 = (torch.randn(C, generator=g) * 0.1).cuda(), (torch.rand(C, generator=g) + 0.5).cuda()
    # separate calls
    ys = y.clone().requires_grad_(True)
    gs, bs = gamma.clone().requires_grad_(True), beta.clone().requires_grad_(True)
    rm_s, rv_s = rm0.clone(), rv0.clone()
    z_s = torch.cat([ops.bn_act_pool(ys[k * Mg:(k + 1) * Mg], gs, bs, rm_s, rv_s, momentum=0.1, pool=pool) for k in range(G)])
    (z_s * w).sum().backward()
    # one grouped call
    yg = y.clone().requires_grad_(True)
    gg, bg = gamma.clone().requires_grad_(True), beta.clone().requires_grad_(True)
    rm_g, rv_g = rm0.clone(), rv0.clone()
    z_g = ops.bn_act_pool(yg, gg, bg, rm_g, rv_g, momentum=0.1, pool=pool, groups=G)
    (z_g * w).sum().backward()
    assert torch.equal(z_g, z_s) and torch.equal(yg.grad, ys.grad)
    assert torch.equal(rm_g, rm_s) and torch.equal(rv_g, rv_s)
    assert_rel(gg.grad.cpu(), gs.grad.cpu(), tol=1e-6, what="grad gamma")
    assert_rel(bg.grad.cpu(), bs.grad.cpu(), tol=1e-6, what="grad beta")


def test_kernels_do_not_write_outside_their_outputs():
    """compute-sanitizer is not available on the GPU pool: outputs are placed inside larger buffers filled with a canary value
    and the bytes before and after them must come back untouched (multi-scan cutout kernel, attention kernel with the operand
    split, tcgen05 convolution writing into a caller's view)."""
    canary = 1234.5
    pad = 4096

    def guarded(shape, dtype=torch.float32):
        n = int(np.prod(shape))
        buf = torch.full((n + 2 * pad,), canary, dtype=dtype, device="cuda")
        return buf, buf[pad:pad + n].view(shape)

    def intact(buf, n):
        return bool((buf[:pad] == canary).all()) and bool((buf[pad + n:] == canary).all())

    phi = synth.phi_for("jrdb")
    n = len(phi)
    scans = torch.from_numpy(np.stack([synth.adversarial_scans(3, n, seed=k) for k in range(2)])).cuda()
    for fast in (True, False):
        buf, out = guarded((2, n, 3, 56))
        ops.cutout(scans, torch.from_numpy(phi).cuda(), fast=fast, out=out, **CFG)
        assert intact(buf, out.numel()) and bool(torch.isfinite(out).all()) and not bool((out == canary).any())
    b, npts, L, C, E = 2, 23, 14, 256, 128
    x, t = torch.randn(b, npts, L, C).cuda(), torch.randn(b, npts, L, C).cuda()
    ex, et = torch.randn(b, npts, E).cuda() * 0.2, torch.randn(b, npts, E).cuda() * 0.2
    b_out, o = guarded((b, npts, L, C))
    b_ff, ff = guarded((b, npts, 11))
    b_sp, sp = guarded((b * npts * L, 2 * C), torch.float16)
    ops.gate_forward(x, t, ex, et, 0.5, 11, out=o, feat_out=ff, split_out=sp, split_channels=C)
    assert intact(b_out, o.numel()) and intact(b_ff, ff.numel()) and intact(b_sp, sp.numel())
    assert not bool((o == canary).any()) and not bool((sp == canary).any())
    M, LA, Cin, Cout = 77, 14, 256, 256
    a = torch.randn(M * LA, Cin).cuda()
    _, a_split = ops.act(a, None, pool=1, slope=1.0, want_plain=False, want_split=True, parts=ops.SPLIT_F16)
    ws, out_scale = _tc_weights((torch.randn(Cout, Cin, 3) * 0.05).cuda(), True)
    b_pl, pl = guarded((M * LA // 2, Cout))
    ops.conv_tc(a_split, ws, None, M, LA, LA, 3, 1, pool=2, want_plain=True, want_split=True, out_scale=out_scale, plain_out=pl)
    torch.cuda.synchronize()
    assert intact(b_pl, pl.numel()) and not bool((pl == canary).any())
