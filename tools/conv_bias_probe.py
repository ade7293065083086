"""Is the tensor core's accumulation error a BIAS?  One backbone layer on activation-like inputs, chain lengths 64 / 128:
max and MEAN SIGNED error of the outputs against fp64, in units of the output's own magnitude (sign-aligned: positive =
the magnitude came out too large).  Needs a B200.

    python tools/conv_bias_probe.py
"""
import os
import sys

import torch
import torch.nn.functional as F

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from planar_optical_flow_b200 import ops                                  # noqa: E402
from planar_optical_flow_b200.engine import _ChannelsLastBackbone        # noqa: E402

dev = torch.device("cuda")
holder = _ChannelsLastBackbone.__new__(_ChannelsLastBackbone)
F16 = os.environ.get("POF_PROBE_TF32", "0") != "1"          # POF_PROBE_TF32=1: TF32 parts (kind::tf32, K = 8 per MMA)
holder.f16 = F16
print("operand parts:", "binary16 (kind::f16)" if F16 else "TF32 (kind::tf32)")
for LA, Cin, Cout in ((56, 64, 64), (28, 128, 128), (14, 256, 256), (14, 256, 512), (7, 512, 256)):
    g = torch.Generator(device="cuda").manual_seed(1)
    M = 2048
    x = torch.randn(M, LA, Cin, generator=g, device=dev)
    x = torch.where(x > 0, x, 0.1 * x)                                   # post-LeakyReLU statistics
    w = torch.randn(Cout, Cin, 3, generator=g, device=dev) * (2.0 / (Cin * 3)) ** 0.5
    ws, out_scale = holder._tc_weight(w)
    _, a = ops.act(x.view(M * LA, Cin), None, pool=1, slope=1.0, want_plain=False, want_split=True, parts=ops.SPLIT_F16 if F16 else 2)
    want = F.conv1d(x.permute(0, 2, 1).double(), w.double(), None, padding=1).permute(0, 2, 1).reshape(-1, Cout)
    for chain, flag in ((64, 0x40000), (128, 0x40000), (256, 0x40000), (64, 0), (128, 0)):      # 0x40000 = POF_CONV_TC_NO_DEBIAS
        if chain > Cin:
            continue
        plain, _ = ops.conv_tc(a, ws, None, M, LA, LA, 3, 1, pool=1, slope=1.0, want_plain=True, want_split=False,
                               out_scale=out_scale, chain_channels=chain | flag)
        err = plain.double() - want
        scale = want.abs().max()
        big = want.abs() > 0.25 * scale                                   # outputs with a meaningful magnitude
        signed = (err * want.sign())[big] / want.abs()[big]
        print("LA=%-2d %3d->%3d chain=%-3d %s max|err|/max|out| %.2e   mean signed rel err of large outputs %+.2e  (std %.2e)  rms rel %.2e" % (
            LA, Cin, Cout, chain, "raw     " if flag else "debiased", float(err.abs().max() / scale), float(signed.mean()), float(signed.std()), float((err[big] / want[big]).pow(2).mean().sqrt())))
