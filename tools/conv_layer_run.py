"""Run ONE backbone layer shape of pof_conv_tc_f16_fwd a few times (the target of an `ncu --set full` capture).

    python tools/conv_layer_run.py LA Cin Cout taps pad pool [M=69824] [reps=3] [flags=0]
"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from planar_optical_flow_b200 import ops                      # noqa: E402
from planar_optical_flow_b200.engine import _ChannelsLastBackbone   # noqa: E402

LA, Cin, Cout, taps, pad, pool = [int(v) for v in sys.argv[1:7]]
M = int(sys.argv[7]) if len(sys.argv) > 7 else 69824
reps = int(sys.argv[8]) if len(sys.argv) > 8 else 3
flags = int(sys.argv[9], 0) if len(sys.argv) > 9 else 0
dev = torch.device("cuda")
g = torch.Generator(device="cuda").manual_seed(0)
x = torch.randn(M * LA, Cin, generator=g, device=dev).abs()
w = torch.randn(Cout, Cin, taps, generator=g, device=dev) * (2.0 / (Cin * taps)) ** 0.5
b = torch.randn(Cout, generator=g, device=dev) * 0.1
holder = _ChannelsLastBackbone.__new__(_ChannelsLastBackbone)
holder.f16 = True
ws, out_scale = holder._tc_weight(w)
_, a = ops.act(x, None, pool=1, slope=1.0, want_plain=False, want_split=True, parts=ops.SPLIT_F16)
del x
Lout = LA if pad else LA - taps + 1
status = ops.new_status(dev)
rows = M * Lout // pool
split = None
e = [torch.cuda.Event(enable_timing=True) for _ in range(reps + 1)]
e[0].record()
for r in range(reps):
    _, split = ops.conv_tc(a, ws, b, M, LA, Lout, taps, pad, pool=pool, slope=0.1, want_plain=False, want_split=True,
                           out_scale=out_scale, chain_channels=flags, status=status)
    e[r + 1].record()
torch.cuda.synchronize()
ms = [e[r].elapsed_time(e[r + 1]) for r in range(reps)]
fl = 2.0 * M * Lout * Cin * Cout * taps
print("LA=%d %d->%d taps=%d pool=%d M=%d flags=%#x status=%d  ms=%s  best %.3f ms = %.1f TFLOP/s algorithmic, %.2f TB/s in+out" % (
    LA, Cin, Cout, taps, pool, M, flags, ops.read_status(status), ["%.3f" % v for v in ms], min(ms), fl / min(ms) / 1e9,
    (a.numel() * 2 + split.numel() * 2) / min(ms) / 1e9))
