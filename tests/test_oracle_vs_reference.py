"""Pin the oracle restatements to the UNMODIFIED reference (build container only).

Bit-equality is required: the oracle is the checker for the CUDA path, so it
must be the reference's arithmetic, not an approximation of it.
"""
import numpy as np
import pytest
import torch

from oracle import cutout as ocut
from oracle import model as omodel
from oracle import nms as onms
from oracle import ref_shim
from planar_optical_flow_b200 import synth

pytestmark = pytest.mark.reference

CFG = dict(fixed=True, centered=True, window_width=1.0, window_depth=0.5,
           num_cutout_pts=56, padding_val=29.99, area_mode=True)   # config/dr_spaam.yaml


@pytest.fixture(scope="module")
def ref():
    return ref_shim.load()


@pytest.mark.parametrize("shape", ["drow", "jrdb"])
@pytest.mark.parametrize("kind", ["adversarial", "structured", "edge"])
@pytest.mark.parametrize("flags", [
    dict(), dict(fixed=False), dict(centered=False), dict(area_mode=False),
    dict(window_width=1.66, window_depth=1.0, num_cutout_pts=48),
])
def test_cutout_bit_equal(ref, shape, kind, flags):
    ru, _ = ref
    phi = synth.phi_for(shape)
    n = len(phi)
    if kind == "adversarial":
        scans = synth.adversarial_scans(3, n, seed=11)
    elif kind == "structured":
        scans = synth.structured_sequence(3, n, seed=12, phi=phi)
    else:
        scans = synth.edge_scans(n, seed=13)
    kw = dict(CFG, **flags)
    want = ru.scans_to_cutout(scans, phi, stride=1, **kw)
    got = ocut.scans_to_cutout(scans, phi, stride=1, **kw)
    assert got.dtype == want.dtype == np.float32 and got.shape == want.shape
    assert np.array_equal(got, want)


def test_cutout_stride_and_f64_scans(ref):
    ru, _ = ref
    phi = synth.drow_phi()
    scans = synth.adversarial_scans(2, 450, seed=3)
    for stride in (2, 3):
        assert np.array_equal(ocut.scans_to_cutout(scans, phi, stride=stride, **CFG),
                              ru.scans_to_cutout(scans, phi, stride=stride, **CFG))
    s64 = scans.astype(np.float64)
    assert np.array_equal(ocut.scans_to_cutout(s64, phi, **CFG), ru.scans_to_cutout(s64, phi, **CFG))


@pytest.mark.parametrize("shape,scan_dtype", [("drow", np.float32), ("drow", np.float64),
                                              ("jrdb", np.float32), ("jrdb", np.float64)])
def test_nms_bit_equal(ref, shape, scan_dtype):
    ru, _ = ref
    phi = synth.phi_for(shape)
    n = len(phi)
    for seed in range(6):
        scan = synth.structured_sequence(1, n, seed=100 + seed, phi=phi)[0].astype(scan_dtype)
        cls = synth.distinct_scores(n, seed)
        reg = synth.clustered_votes(scan.astype(np.float64), phi, seed)
        want = ru.nms_predicted_center(scan, phi, cls, reg)
        got = onms.nms_predicted_center(scan, phi, cls, reg)
        spec = onms.nms_sweep_spec(scan, phi, cls, reg)
        for w, g in zip(want, got):
            assert w.dtype == g.dtype and np.array_equal(w, g)
        assert np.array_equal(spec["det_xys"], want[0])
        assert np.array_equal(spec["det_cls"], want[1])
        assert np.array_equal(spec["instance_mask"], want[2])
        assert 1 < len(want[0]) < n          # both suppression and survivors happen


def _ref_model(rm, sd, alpha=0.5, window=11, num_pts=56):
    m = rm.SpatialDROW(num_scans=10, num_pts=num_pts, focal_loss_gamma=0.0, alpha=alpha,
                       window_size=window, pedestrian_only=True)
    m.load_state_dict(sd, strict=True)
    return m


def test_state_dict_keys_match_reference(ref):
    _, rm = ref
    m = rm.SpatialDROW(num_scans=10, num_pts=56, alpha=0.5, window_size=11, pedestrian_only=True)
    want = {k: tuple(v.shape) for k, v in m.state_dict().items()}
    got = {k: tuple(v.shape) for k, v in omodel.init_state_dict(56, True).items()}
    assert want == got


@pytest.mark.parametrize("n", [37, 120])
def test_stream_forward_matches_reference(ref, n):
    _, rm = ref
    sd = omodel.randomize_bn_stats(omodel.init_state_dict(56, True, seed=4))
    m = _ref_model(rm, sd).eval()
    torch.manual_seed(0)
    tmpl_ref = tmpl_or = None
    with torch.no_grad():
        for step in range(3):
            x = torch.randn(2, n, 1, 56)
            c_r, r_r, tmpl_ref, f_r = m(x, testing=True, fea_template=tmpl_ref)
            c_o, r_o, tmpl_or, f_o = omodel.spatial_drow_stream(x, sd, 0.5, 11, tmpl_or)
            for a, b in ((c_r, c_o), (r_r, r_o), (tmpl_ref, tmpl_or), (f_r, f_o)):
                assert torch.equal(a, b)


def test_sequence_forward_matches_reference_eval_and_train(ref):
    _, rm = ref
    torch.manual_seed(1)
    x = torch.randn(2, 23, 4, 56)
    for training in (False, True):
        sd = omodel.randomize_bn_stats(omodel.init_state_dict(56, True, seed=5))
        m = _ref_model(rm, sd)
        m.train(training)
        sd = {k: v.clone() for k, v in sd.items()}
        with torch.no_grad():
            want = m(x)
            got = omodel.spatial_drow_sequence(x, sd, 0.5, 11, training=training)
        for a, b in zip(want, got):
            assert torch.equal(a, b)
        if training:   # running statistics advanced identically (gate BN twice per step)
            for k, v in m.state_dict().items():
                assert torch.equal(v, sd[k]), k


def test_windowed_gate_equals_dense(ref):
    sd = omodel.randomize_bn_stats(omodel.init_state_dict(56, True, seed=6))
    torch.manual_seed(2)
    x = torch.from_numpy(synth.feature_like((2, 64, 256, 14), 1))
    t = torch.from_numpy(synth.feature_like((2, 64, 256, 14), 2))
    with torch.no_grad():
        dense_t, dense_f, _ = omodel.gate_dense(x, t, sd, 0.5, 11)
        win_t, win_f, _ = omodel.gate_windowed(x, t, omodel.gate_embed(x, sd), omodel.gate_embed(t, sd), 0.5, 11)
    # 1e-5 relative (BASELINE.json), measured against the tensor's magnitude: the
    # softmax turns an absolute score error ds into a RELATIVE weight error ds, so
    # element-wise relative error on near-zero outputs is not meaningful.
    assert (dense_f - win_f).abs().max() <= 1e-5 * dense_f.abs().max()
    assert (dense_t - win_t).abs().max() <= 1e-5 * dense_t.abs().max()


def test_drow_forward_matches_reference(ref):
    _, rm = ref
    sd = omodel.randomize_bn_stats(omodel.init_state_dict(56, True, seed=7))
    sd_drow = {k: v for k, v in sd.items() if not k.startswith("gate.")}
    m = rm.DROW(num_scans=3, num_pts=56, pedestrian_only=True)
    m.load_state_dict(sd_drow, strict=True)
    m.eval()
    torch.manual_seed(3)
    x = torch.randn(2, 19, 3, 56)
    with torch.no_grad():
        want = m(x)
        got = omodel.drow_forward(x, sd_drow)
    for a, b in zip(want, got):
        assert torch.allclose(a, b, rtol=1e-5, atol=1e-6)


# ------------------------------------------------------------------ scan-pair flow prototype (row N3)
@pytest.mark.reference
@pytest.mark.parametrize("b,c,n,k,d", [(2, 24, 57, 3, 5), (1, 8, 7, 3, 5), (3, 4, 20, 5, 2), (1, 3, 1, 3, 5)])
def test_prototype_fusion_bit_equal(b, c, n, k, d):
    from oracle import prototype as oproto
    from oracle import ref_shim

    rp = ref_shim.load_prototype()
    torch.manual_seed(b * 100 + n)
    f1, f2 = torch.randn(b, c, n), torch.randn(b, c, n)
    m = rp.Prototype(in_channel=2, max_displacement=d)
    with ref_shim.cpu_cuda_noop():
        want = m._fusion(f1, f2, kernel_size=k, max_displacement=d)
    assert torch.equal(oproto.fusion_dense(f1, f2, k, d), want)
    got = oproto.fusion_windowed(f1, f2, k, d)
    assert float((got - want).abs().max()) <= 1e-5 * float(want.abs().max())


@pytest.mark.reference
def test_prototype_forward_bit_equal():
    from oracle import prototype as oproto
    from oracle import ref_shim

    rp = ref_shim.load_prototype()
    sd = oproto.init_state_dict(2, 5, seed=3)
    m = rp.Prototype(in_channel=2, max_displacement=5)
    m.load_state_dict(sd, strict=True)
    m.eval()
    torch.manual_seed(4)
    s1 = torch.randn(2, 450, 2)
    s2 = s1 + 0.1 * torch.randn(2, 450, 2)
    with torch.no_grad(), ref_shim.cpu_cuda_noop():
        want = m(s1, s2)
        want_self = m(s1)
    with torch.no_grad():
        assert torch.equal(oproto.prototype_forward(s1, s2, sd), want)
        assert torch.equal(oproto.prototype_forward(s1, s1, sd), want_self)


# ------------------------------------------------------------------ legacy preprocessing (row N4)
@pytest.mark.reference
def test_resize_column_matches_cv2():
    import cv2

    from oracle import cutout_legacy as ol

    rs = np.random.RandomState(0)
    for _ in range(300):
        n, P = int(rs.randint(1, 600)), int(rs.choice([48, 56, 32, 7]))
        col = (rs.rand(n) * 30).astype(np.float32)
        area = P < n
        want = cv2.resize(col, (1, P), interpolation=cv2.INTER_AREA if area else cv2.INTER_LINEAR).reshape(-1)
        got = ol.resize_column(col, P, area)
        assert np.abs(got - want).max() <= 1e-6 * np.abs(want).max(), (n, P)


@pytest.mark.reference
@pytest.mark.parametrize("shape", ["drow", "jrdb"])
@pytest.mark.parametrize("kw", [dict(), dict(fixed=False, centered=False, window_width=1.0, window_depth=0.5, num_cutout_pts=56)])
def test_cutout_original_and_polar_grid_match_reference(shape, kw):
    from oracle import cutout_legacy as ol
    from oracle import ref_shim

    ru, _ = ref_shim.load()
    phi = synth.phi_for(shape)
    scans = np.concatenate([synth.structured_sequence(2, len(phi), seed=5, phi=phi), synth.adversarial_scans(1, len(phi), seed=6)])
    incre = phi[1] - phi[0]
    want = ru.scans_to_cutout_original(scans, incre, **kw)
    got = ol.scans_to_cutout_original(scans, incre, **kw)
    assert got.dtype == want.dtype == np.float32 and got.shape == want.shape
    assert np.abs(got - want).max() <= 1e-6 * np.abs(want).max()
    assert (got == want).mean() > 0.999
    assert np.array_equal(ol.scans_to_polar_grid(scans), ru.scans_to_polar_grid(scans))
    assert np.array_equal(ol.scans_to_polar_grid(scans, 0.5, 20.0, 0.5, 0.0, False), ru.scans_to_polar_grid(scans, 0.5, 20.0, 0.5, 0.0, False))
