"""Small invocations of every libpof kernel added or changed in round 2, for `compute-sanitizer --tool memcheck` (and racecheck).

    compute-sanitizer --tool memcheck python tools/sanitize_run.py
"""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from planar_optical_flow_b200 import ops, synth                                   # noqa: E402
from planar_optical_flow_b200.engine import StreamingDetector                     # noqa: E402
from planar_optical_flow_b200.model import SpatialDROW                            # noqa: E402

CFG = dict(fixed=True, centered=True, window_width=1.0, window_depth=0.5, num_cutout_pts=56, padding_val=29.99, area_mode=True)
dev = torch.device("cuda")
torch.manual_seed(0)

# cutout: both span reductions, FAST scan kernel for S = 1 and S > 1, fixed and not
phi = synth.phi_for("jrdb")
phi_d = torch.from_numpy(phi).to(dev)
for S in (1, 3):
    scans = torch.from_numpy(np.stack([synth.adversarial_scans(S, len(phi), seed=k) for k in range(3)])).to(dev)
    for fast in (False, True):
        for fixed in (True, False):
            ops.cutout(scans, phi_d, fast=fast, **dict(CFG, fixed=fixed))
            ops.cutout(scans, phi_d, fast=fast, return_s_area=True, **dict(CFG, fixed=fixed))

# attention kernel: forward with the operand split, backward (MODE 2 + embed + MODE 1)
b, n, L, C, E = 2, 37, 14, 256, 128
x = torch.randn(b, n, L, C, device=dev, requires_grad=True)
t = torch.randn(b, n, L, C, device=dev, requires_grad=True)
ex = (torch.randn(b, n, E, device=dev) * 0.2).requires_grad_(True)
et = (torch.randn(b, n, E, device=dev) * 0.2).requires_grad_(True)
split = torch.empty((b * n * L, 2 * C), dtype=torch.float16, device=dev)
st = ops.new_status(dev)
ops.gate_forward(x.detach(), t.detach(), ex.detach(), et.detach(), 0.5, 11, split_out=split, split_channels=C, status=st)
out, ff = ops.gate(x, t, ex, et, 0.5, 11)
(out.square().mean() + ff.square().mean()).backward()

# training operator: batch-norm + LeakyReLU + pool, grouped and not; first-layer kernels
for M, Lr, Cc, pool, groups in ((36, 56, 64, 1, 1), (36, 56, 128, 2, 3), (10, 14, 512, 2, 2), (30, 1, 128, 1, 1)):
    y = torch.randn(M, Cc, 1, Lr, device=dev).contiguous(memory_format=torch.channels_last).requires_grad_(True)
    g, be = torch.rand(Cc, device=dev).requires_grad_(True), torch.randn(Cc, device=dev).requires_grad_(True)
    rm, rv = torch.zeros(Cc, device=dev), torch.ones(Cc, device=dev)
    z = ops.bn_act_pool(y, g, be, rm, rv, pool=pool, groups=groups)
    z.square().mean().backward()
cut = torch.randn(41, 56, device=dev).clamp(-1, 1)
w = torch.randn(64, 1, 3, device=dev, requires_grad=True)
ops.conv_first_train(cut, w).square().mean().backward()

# the engine (tcgen05 convolutions with the warp-uniform roles, fused split, in-place outputs) and a CUDA-graph replay
model = SpatialDROW(num_scans=10, num_pts=56, alpha=0.5, window_size=11, pedestrian_only=True)
phi_s = synth.drow_phi(70)
seq = np.stack([synth.structured_sequence(5, 70, seed=k, phi=phi_s) for k in range(2)], axis=1)
for graph in (False, True):
    det = StreamingDetector(model, phi_s, CFG, 2, cuda_graph=graph)
    for tt in range(5):
        det.step(seq[tt])
    det.check()

# the training branch: 3 scans, 2 per call
m = SpatialDROW(num_scans=3, num_pts=56, alpha=0.5, window_size=11, pedestrian_only=True).to(dev).train()
m.scans_per_call = 2
cls, reg, sim = m(torch.randn(2, 19, 3, 56, device=dev))
(cls.square().mean() + reg.square().mean() + sim.square().mean()).backward()
torch.cuda.synchronize()
print("sanitize_run: all launches completed")
