"""Time the cutout kernel on bench-shaped inputs (structured and adversarial ranges)."""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from planar_optical_flow_b200 import ops, synth  # noqa: E402

CFG = dict(fixed=True, centered=True, window_width=1.0, window_depth=0.5, num_cutout_pts=56, padding_val=29.99,
           area_mode=True)
dev = torch.device("cuda")
ONLY = os.environ.get("TUNE_ONLY")
for shape in ("jrdb", "drow") if not ONLY else ("jrdb",):
    phi = synth.phi_for(shape)
    n = len(phi)
    phi_d = torch.from_numpy(phi).to(dev)
    for kind in ("structured", "adversarial") if not ONLY else ("structured",):
        for B, S in ((256, 1), (4096, 1), (64, 11), (512, 11)) if not ONLY else ((4096, 1),):
            if kind == "structured":
                base = np.stack([synth.structured_sequence(S, n, seed=k, phi=phi) for k in range(16)])
                scans = np.tile(base, (B // 16 + 1, 1, 1))[:B]
            else:
                scans = np.stack([synth.adversarial_scans(S, n, seed=k) for k in range(min(B, 64))])
                scans = np.tile(scans, (B // len(scans) + 1, 1, 1))[:B]
            s = torch.from_numpy(np.ascontiguousarray(scans)).to(dev)
            out = torch.empty((B, n, S, 56), device=dev)
            for mode in ("EXACT", "EXACT-pieces", "FAST"):
                fast = dict(fast=mode == "FAST", exact_pieces=mode == "EXACT-pieces")
                for _ in range(3):
                    _, sa = ops.cutout(s, phi_d, out=out, return_s_area=True, **fast, **CFG)
                torch.cuda.synchronize()
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record()
                for _ in range(10):
                    ops.cutout(s, phi_d, out=out, **fast, **CFG)
                e1.record()
                torch.cuda.synchronize()
                ms = e0.elapsed_time(e1) / 10
                gb = B * S * n * 228 / 1e9
                print("%-5s %-11s B=%4d S=%2d %-12s %8.3f ms  %7.1f GB/s (%.1f%% of 6547)  s_area max %d" %
                      (shape, kind, B, S, mode, ms, gb / ms * 1e3, 100 * gb / ms * 1e3 / 6547,
                       int(sa.max())))
