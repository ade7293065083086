"""Relative error (magnitude-relative) of every engine precision mode, per output and step, against BOTH the CPU
oracle's float32 reference loop and a float64 evaluation of the same network on the same cutouts (the arbiter: it
tells how far the float32 oracle itself is from the exact result).

    python tests/precision_report.py [steps] [sequences] [points] [modes] [chains]     (needs a B200; reads nothing outside the repo)

`modes`: comma list of engine precisions; a mode written `fp32@64` runs with POF_CONV_TC_CHAIN=64 (tensor-core chain length).
"""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import cutout as ocut            # noqa: E402
from oracle import model as omodel           # noqa: E402
from planar_optical_flow_b200 import synth   # noqa: E402
from planar_optical_flow_b200.engine import StreamingDetector   # noqa: E402
from planar_optical_flow_b200.model import SpatialDROW          # noqa: E402

CFG = dict(fixed=True, centered=True, window_width=1.0, window_depth=0.5, num_cutout_pts=56, padding_val=29.99, area_mode=True)


def rel(a, b):
    a, b = np.asarray(a, np.float64), np.asarray(b, np.float64)
    return float(np.abs(a - b).max() / max(np.abs(b).max(), 1e-30))


def main():
    steps = int(sys.argv[1]) if len(sys.argv) > 1 else 4
    b = int(sys.argv[2]) if len(sys.argv) > 2 else 2
    n = int(sys.argv[3]) if len(sys.argv) > 3 else 200
    modes = sys.argv[4].split(",") if len(sys.argv) > 4 else ["fp32", "tf32x3", "tf32"]
    phi = synth.drow_phi(n)
    scans = np.stack([synth.structured_sequence(steps, n, seed=40 + k, phi=phi) for k in range(b)], axis=1)
    sd = omodel.randomize_bn_stats(omodel.init_state_dict(56, True, seed=9))
    sd64 = {k: (v.double() if v.is_floating_point() else v) for k, v in sd.items()}
    want, want64 = [], []
    tmpl, tmpl64 = [None] * b, [None] * b
    with torch.no_grad():
        for t in range(steps):
            row, row64 = [], []
            for k in range(b):
                ct = ocut.scans_to_cutout(scans[t, k][None], phi, **CFG)
                cls, reg, tmpl[k], ff = omodel.spatial_drow_stream(torch.from_numpy(ct)[None], sd, 0.5, 11, tmpl[k])
                row.append((torch.sigmoid(cls[0, :, 0]).numpy(), reg[0].numpy(), ff[0].numpy(), tmpl[k][0].numpy()))
                cls, reg, tmpl64[k], ff = omodel.spatial_drow_stream(torch.from_numpy(ct)[None].double(), sd64, 0.5, 11, tmpl64[k])
                row64.append((torch.sigmoid(cls[0, :, 0]).numpy(), reg[0].numpy(), ff[0].numpy(), tmpl64[k][0].numpy()))
            want.append(row)
            want64.append(row64)
    names = ("scores", "votes", "similarities", "memory")
    print("max |a - b| / max |b|, %d sequences x %d points, per step; 'o32' = vs the float32 oracle, 'f64' = vs the float64 evaluation" % (b, n))
    for t in range(steps):
        e = [max(rel(want[t][k][i], want64[t][k][i]) for k in range(b)) for i in range(4)]
        print("  %-10s step %d  " % ("oracle32", t) + "  ".join("%s f64 %.2e" % (nm, v) for nm, v in zip(names, e)))
    for spec in modes:
        prec, _, chain = spec.partition("@")
        os.environ["POF_CONV_TC_CHAIN"] = chain or "0"
        m = SpatialDROW(num_scans=10, num_pts=56, alpha=0.5, window_size=11, pedestrian_only=True)
        m.load_state_dict(sd, strict=True)
        det = StreamingDetector(m.cuda(), phi, CFG, b, precision=prec)
        for t in range(steps):
            det.step(scans[t])
            got = lambda k: (det._last["pred_cls"][k].cpu().numpy(), det._last["pred_reg"][k].cpu().numpy(),      # noqa: E731
                             det._last["feat_fused"][k].cpu().numpy(), det.template[k].cpu().numpy())
            g = [got(k) for k in range(b)]
            e32 = [max(rel(g[k][i], want[t][k][i]) for k in range(b)) for i in range(4)]
            e64 = [max(rel(g[k][i], want64[t][k][i]) for k in range(b)) for i in range(4)]
            print("  %-10s step %d  " % (spec, t) + "  ".join("%s o32 %.2e f64 %.2e" % (nm, a, c) for nm, a, c in zip(names, e32, e64)))
    os.environ.pop("POF_CONV_TC_CHAIN", None)


main()
