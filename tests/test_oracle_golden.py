"""The oracle restatements against the committed golden fixtures (reference outputs
frozen by oracle/make_golden.py).  Runs anywhere: no GPU, no reference mount."""
import ast
import glob
import os

import numpy as np
import pytest
import torch

from oracle import cutout as ocut
from oracle import model as omodel
from oracle import nms as onms
from oracle.make_golden import weights_digest
from planar_optical_flow_b200 import synth


def _load(golden_dir, name):
    return np.load(os.path.join(golden_dir, name + ".npz"))


def test_fixtures_present(golden_dir):
    names = {os.path.basename(p) for p in glob.glob(os.path.join(golden_dir, "*.npz"))}
    assert len(names) >= 12, names


@pytest.mark.parametrize("name", ["cutout_drow_adversarial", "cutout_drow_structured_lastref", "cutout_drow_edge",
                                  "cutout_jrdb_adversarial", "cutout_jrdb_structured_raw", "cutout_drow_linear48"])
def test_cutout_golden_bit_equal(golden_dir, name):
    g = _load(golden_dir, name)
    kw = dict(ast.literal_eval(str(g["kwargs"])))
    got = ocut.scans_to_cutout(g["scans"], g["phi"], stride=1, **kw)
    assert np.array_equal(got, g["out"])


@pytest.mark.parametrize("name", ["nms_drow_f32scan", "nms_drow_f64scan", "nms_jrdb_f32scan", "nms_jrdb_f64scan"])
def test_nms_golden_bit_equal(golden_dir, name):
    g = _load(golden_dir, name)
    xy, c, mask = onms.nms_predicted_center(g["scan"], g["phi"], g["cls"], g["reg"])
    assert np.array_equal(xy, g["det_xys"]) and xy.dtype == g["det_xys"].dtype
    assert np.array_equal(c, g["det_cls"])
    assert np.array_equal(mask, g["instance_mask"]) and mask.dtype == np.int32
    spec = onms.nms_sweep_spec(g["scan"], g["phi"], g["cls"], g["reg"])
    assert np.array_equal(spec["instance_mask"], g["instance_mask"])
    assert np.array_equal(spec["det_xys"], g["det_xys"])


def test_nms_tie_rule_is_stable_reversed():
    cls = np.array([[0.5], [0.9], [0.5], [0.1], [0.9]], dtype=np.float32)
    assert onms.descending_order(cls[:, 0]).tolist() == [4, 1, 2, 0, 3]


def test_gate_golden(golden_dir):
    g = _load(golden_dir, "gate_n24")
    seed = int(g["seed"])
    sd = omodel.randomize_bn_stats(omodel.init_state_dict(56, True, seed=seed), seed=seed + 1)
    assert weights_digest(sd) == str(g["weights_sha256"]), "seeded weight generator drifted"
    x = torch.from_numpy(synth.feature_like((1, 24, 256, 14), seed + 2))
    t = torch.from_numpy(synth.feature_like((1, 24, 256, 14), seed + 3))
    with torch.no_grad():
        out, ff, _ = omodel.gate_dense(x, t, sd, 0.5, 11)
    assert np.array_equal(out.numpy(), g["out_temp"])
    assert np.array_equal(ff.numpy(), g["feat_fused"])


def test_model_stream_golden(golden_dir):
    g = _load(golden_dir, "model_stream_drow40")
    seed = int(g["seed"])
    sd = omodel.randomize_bn_stats(omodel.init_state_dict(56, True, seed=seed), seed=seed + 1)
    assert weights_digest(sd) == str(g["weights_sha256"])
    cfg = dict(fixed=True, centered=True, window_width=1.0, window_depth=0.5, num_cutout_pts=56,
               padding_val=29.99, area_mode=True)
    scans, phi = g["scans"], g["phi"]
    tmpl = None
    with torch.no_grad():
        for t in range(scans.shape[1]):
            ct = np.stack([ocut.scans_to_cutout(scans[k, t:t + 1], phi, **cfg) for k in range(scans.shape[0])])
            cls, reg, tmpl, ff = omodel.spatial_drow_stream(torch.from_numpy(ct), sd, 0.5, 11, tmpl)
            assert np.array_equal(cls.numpy(), g["cls_%d" % t])
            assert np.array_equal(reg.numpy(), g["reg_%d" % t])
            assert np.array_equal(ff.numpy(), g["feat_fused_%d" % t])
    assert np.array_equal(tmpl.numpy()[:, ::8, ::16], g["template_last_sample"])


def test_prototype_oracle_matches_reference_golden(golden_dir):
    import os

    import torch

    from oracle import prototype as oproto

    g = np.load(os.path.join(golden_dir, "prototype_drow450.npz"))
    sd = oproto.init_state_dict(2, 5, seed=int(g["seed"]))
    with torch.no_grad():
        flow = oproto.prototype_forward(torch.from_numpy(g["scan1"]), torch.from_numpy(g["scan2"]), sd)
        fused = oproto.fusion_dense(torch.from_numpy(g["f1"]), torch.from_numpy(g["f2"]), 3, 5)
    assert np.array_equal(fused.numpy(), g["fused"])
    assert np.abs(flow.numpy() - g["flow"]).max() <= 1e-6 * np.abs(g["flow"]).max()


@pytest.mark.parametrize("shape", ["drow", "jrdb"])
def test_legacy_oracle_matches_reference_golden(golden_dir, shape):
    import os

    from oracle import cutout_legacy as ol

    g = np.load(os.path.join(golden_dir, "cutout_original_%s.npz" % shape))
    kw = dict(fixed=True, centered=True, window_width=1.66, window_depth=1.0, num_cutout_pts=48, padding_val=29.99)
    incre = g["incre"][()]
    for scans, want in ((g["scans"], g["out"]), (g["adv"], g["out_adv"])):
        got = ol.scans_to_cutout_original(scans, incre, **kw)
        assert np.abs(got - want).max() <= 1e-6 * np.abs(want).max()
    assert np.array_equal(ol.scans_to_polar_grid(g["scans"][:, ::16]), g["polar"])
    assert np.array_equal(ol.scans_to_polar_grid(g["adv"][:, ::16], 0.5, 20.0, 0.5, 0.0, False), g["polar_raw"])
