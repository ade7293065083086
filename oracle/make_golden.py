"""Freeze outputs of the UNMODIFIED reference into tests/golden/*.npz.

TEST INFRASTRUCTURE.  Run in the build container (needs /root/reference):

    python -m oracle.make_golden

The reference itself carries no golden vectors, KATs or fixtures for this path
(SURVEY.md §4), so these files are the pin: inputs come from the seeded
generators in planar_optical_flow_b200/synth.py, outputs from the reference's
own `scans_to_cutout`, `nms_predicted_center` and `SpatialDROW`
(/root/reference/src/utils/utils.py:259-334,535-571;
 /root/reference/src/depracted/model/dr_spaam.py:124-277) executed on this
container's CPU (NumPy 2.3.5, torch 2.11.0).  The fixtures travel to the GPU
box; /root/reference does not.
"""
import hashlib
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from oracle import model as omodel  # noqa: E402
from oracle import ref_shim  # noqa: E402
from planar_optical_flow_b200 import synth  # noqa: E402

OUT = os.path.join(ROOT, "tests", "golden")
CFG = dict(fixed=True, centered=True, window_width=1.0, window_depth=0.5,
           num_cutout_pts=56, padding_val=29.99, area_mode=True)


def weights_digest(sd):
    h = hashlib.sha256()
    for k in sorted(sd):
        h.update(k.encode())
        h.update(sd[k].detach().cpu().numpy().tobytes())
    return h.hexdigest()


def save(name, **arrays):
    path = os.path.join(OUT, name + ".npz")
    np.savez_compressed(path, **arrays)
    print("%-40s %8.1f KB" % (name, os.path.getsize(path) / 1024))


def cutout_cases(ru):
    cases = [
        ("cutout_drow_adversarial", "drow", synth.adversarial_scans(2, 450, seed=21), dict()),
        ("cutout_drow_structured_lastref", "drow",
         synth.structured_sequence(3, 450, seed=22, phi=synth.drow_phi()), dict(fixed=False)),
        ("cutout_drow_edge", "drow", synth.edge_scans(450, seed=23), dict()),
        ("cutout_jrdb_adversarial", "jrdb", synth.adversarial_scans(1, 1091, seed=24), dict()),
        ("cutout_jrdb_structured_raw", "jrdb",
         synth.structured_sequence(2, 1091, seed=25, phi=synth.jrdb_phi()), dict(centered=False)),
        ("cutout_drow_linear48", "drow", synth.adversarial_scans(2, 450, seed=26),
         dict(area_mode=False, window_width=1.66, window_depth=1.0, num_cutout_pts=48)),
    ]
    for name, shape, scans, flags in cases:
        kw = dict(CFG, **flags)
        phi = synth.phi_for(shape)
        out = ru.scans_to_cutout(scans, phi, stride=1, **kw)
        # the half-angles NumPy's float32 arctan produced on THIS machine (utils.py:279):
        # the only platform-dependent step, frozen so the fixture is self-contained
        ref = scans if kw["fixed"] else np.broadcast_to(scans[-1], scans.shape)
        half_alpha = np.arctan(0.5 * kw["window_width"] / np.maximum(ref, 1e-2))
        save(name, scans=scans, phi=phi, out=out, half_alpha=half_alpha,
             kwargs=np.array(repr(sorted(kw.items()))))


def nms_cases(ru):
    for name, shape, dt, seed in [("nms_drow_f32scan", "drow", np.float32, 31),
                                  ("nms_drow_f64scan", "drow", np.float64, 32),
                                  ("nms_jrdb_f32scan", "jrdb", np.float32, 33),
                                  ("nms_jrdb_f64scan", "jrdb", np.float64, 34)]:
        phi = synth.phi_for(shape)
        n = len(phi)
        scan = synth.structured_sequence(1, n, seed=seed, phi=phi)[0].astype(dt)
        cls = synth.distinct_scores(n, seed)
        reg = synth.clustered_votes(scan.astype(np.float64), phi, seed)
        xy, c, mask = ru.nms_predicted_center(scan, phi, cls, reg)
        save(name, scan=scan, phi=phi, cls=cls, reg=reg, det_xys=xy, det_cls=c, instance_mask=mask)


def model_cases(rm):
    # streaming branch, 3 steps, memory carried (dr_spaam.py:239-250)
    seed, n, b = 41, 40, 2
    sd = omodel.randomize_bn_stats(omodel.init_state_dict(56, True, seed=seed), seed=seed + 1)
    m = rm.SpatialDROW(num_scans=10, num_pts=56, alpha=0.5, window_size=11, pedestrian_only=True)
    m.load_state_dict(sd, strict=True)
    m.eval()
    phi = synth.drow_phi(n)
    scans = np.stack([synth.structured_sequence(3, n, seed=seed + 10 + k, phi=phi) for k in range(b)])  # [B,T,N]
    ru, _ = ref_shim.load()
    arrays = dict(scans=scans, phi=phi, weights_sha256=np.array(weights_digest(sd)),
                  seed=np.array(seed))
    tmpl = None
    with torch.no_grad():
        for t in range(scans.shape[1]):
            ct = np.stack([ru.scans_to_cutout(scans[k, t:t + 1], phi, stride=1, **CFG) for k in range(b)])
            cls, reg, tmpl, ff = m(torch.from_numpy(ct), testing=True, fea_template=tmpl)
            arrays["cls_%d" % t] = cls.numpy()
            arrays["reg_%d" % t] = reg.numpy()
            arrays["feat_fused_%d" % t] = ff.numpy()
        arrays["template_last_sample"] = tmpl.numpy()[:, ::8, ::16]      # spot rows, keeps the file small
        arrays["template_last_sum"] = tmpl.double().sum(dim=(2, 3)).numpy()
    save("model_stream_drow40", **arrays)

    # the gate alone on seeded feature-like inputs (dr_spaam.py:163-217)
    seed, n, b = 51, 24, 1
    sd = omodel.randomize_bn_stats(omodel.init_state_dict(56, True, seed=seed), seed=seed + 1)
    gate = rm._SpatialAttention(n_pts=14, n_channel=256, alpha=0.5, window_size=11)
    gate.load_state_dict({k[len("gate."):]: v for k, v in sd.items() if k.startswith("gate.")}, strict=True)
    gate.eval()
    x = synth.feature_like((b, n, 256, 14), seed + 2)
    t = synth.feature_like((b, n, 256, 14), seed + 3)
    with torch.no_grad():
        out, ff = gate(torch.from_numpy(x), torch.from_numpy(t))
    save("gate_n24", out_temp=out.numpy(), feat_fused=ff.numpy(), seed=np.array(seed),
         weights_sha256=np.array(weights_digest(sd)))


def prototype_cases():
    """Scan-pair flow prototype (prototype.py:34-156): the fusion alone and the whole forward, eval mode."""
    from oracle import prototype as oproto

    rp = ref_shim.load_prototype()
    seed, b, n = 61, 3, 450
    sd = oproto.init_state_dict(2, 5, seed=seed)
    m = rp.Prototype(in_channel=2, max_displacement=5)
    m.load_state_dict(sd, strict=True)
    m.eval()
    g = torch.Generator().manual_seed(seed + 1)
    scan1 = torch.randn(b, n, 2, generator=g) * 3
    scan2 = scan1 + 0.05 * torch.randn(b, n, 2, generator=g)
    f1, f2 = torch.randn(2, 24, 19, generator=g), torch.randn(2, 24, 19, generator=g)
    with torch.no_grad(), ref_shim.cpu_cuda_noop():
        flow = m(scan1, scan2)
        fused = m._fusion(f1, f2, kernel_size=3, max_displacement=5)
    save("prototype_drow450", scan1=scan1.numpy(), scan2=scan2.numpy(), flow=flow.numpy(), f1=f1.numpy(), f2=f2.numpy(),
         fused=fused.numpy(), seed=np.array(seed), weights_sha256=np.array(weights_digest(sd)))


def legacy_cases(ru):
    """scans_to_cutout_original / scans_to_polar_grid (utils.py:423-531) on DROW- and JRDB-shaped scans."""
    for shape, n in (("drow", 450), ("jrdb", 1091)):
        phi = synth.phi_for(shape)
        scans = synth.structured_sequence(2, n, seed=81, phi=phi)
        adv = synth.adversarial_scans(2, n, seed=82)
        adv[0, :7] = [0.004, 0.0099, 0.01, 0.02, 0.05, 29.99, 0.3]          # tiny ranges: windows longer than the scan
        incre = phi[1] - phi[0]
        kw = dict(fixed=True, centered=True, window_width=1.66, window_depth=1.0, num_cutout_pts=48, padding_val=29.99)
        save("cutout_original_%s" % shape, scans=scans, adv=adv, incre=np.array(incre),
             out=ru.scans_to_cutout_original(scans, incre, **kw), out_adv=ru.scans_to_cutout_original(adv, incre, **kw),
             out_lastref=ru.scans_to_cutout_original(scans, incre, **dict(kw, fixed=False, centered=False, num_cutout_pts=56)),
             polar=ru.scans_to_polar_grid(scans[:, ::16]), polar_raw=ru.scans_to_polar_grid(adv[:, ::16], 0.5, 20.0, 0.5, 0.0, False))


def main():
    os.makedirs(OUT, exist_ok=True)
    ru, rm = ref_shim.load()
    only = [a for a in sys.argv[1:] if a.startswith("--only-")]
    if not only:
        cutout_cases(ru)
        nms_cases(ru)
        model_cases(rm)
    if not only or "--only-prototype" in only:
        prototype_cases()
    if not only or "--only-legacy" in only:
        legacy_cases(ru)


if __name__ == "__main__":
    main()
