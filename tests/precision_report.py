"""Relative error (vs the CPU oracle, magnitude-relative) of every engine precision mode, per output and step.

    python tests/precision_report.py [steps] [sequences] [points]        (needs a B200; reads nothing outside the repo)
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
    want = []
    tmpl = [None] * b
    with torch.no_grad():
        for t in range(steps):
            row = []
            for k in range(b):
                ct = ocut.scans_to_cutout(scans[t, k][None], phi, **CFG)
                cls, reg, tmpl[k], ff = omodel.spatial_drow_stream(torch.from_numpy(ct)[None], sd, 0.5, 11, tmpl[k])
                row.append((torch.sigmoid(cls[0, :, 0]).numpy(), reg[0].numpy(), ff[0].numpy(), tmpl[k][0].numpy()))
            want.append(row)
    print("max |got - oracle| / max |oracle|, %d sequences x %d points, per step" % (b, n))
    for prec in modes:
        m = SpatialDROW(num_scans=10, num_pts=56, alpha=0.5, window_size=11, pedestrian_only=True)
        m.load_state_dict(sd, strict=True)
        det = StreamingDetector(m.cuda(), phi, CFG, b, precision=prec)
        for t in range(steps):
            det.step(scans[t])
            e = {"scores": 0.0, "votes": 0.0, "similarities": 0.0, "memory": 0.0}
            for k in range(b):
                c, r, f, mem = want[t][k]
                e["scores"] = max(e["scores"], rel(det._last["pred_cls"][k].cpu().numpy(), c))
                e["votes"] = max(e["votes"], rel(det._last["pred_reg"][k].cpu().numpy(), r))
                e["similarities"] = max(e["similarities"], rel(det._last["feat_fused"][k].cpu().numpy(), f))
                e["memory"] = max(e["memory"], rel(det.template[k].cpu().numpy(), mem))
            print("  %-7s step %d  " % (prec, t) + "  ".join("%s %.2e" % kv for kv in e.items()))


main()
