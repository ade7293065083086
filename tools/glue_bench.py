"""Time the backbone glue kernels (first layer, operand split, heads) at the engine's chunk size.  Needs a B200."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from planar_optical_flow_b200 import ops  # noqa: E402

dev = torch.device("cuda")
M = int(sys.argv[1]) if len(sys.argv) > 1 else 64 * 1091


def timed(fn, n=10):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n


cut = torch.rand(M, 56, device=dev) * 2 - 1
w, b = torch.randn(64, 3, device=dev), torch.randn(64, device=dev)
ms = timed(lambda: ops.conv_first(cut, w, b, want_plain=False, want_split=True, parts=ops.SPLIT_F16))
print("conv_first f16 split  M=%d: %.3f ms  %.0f GB/s written" % (M, ms, M * 56 * 64 * 4 / ms / 1e6))
mem = torch.randn(M * 14, 256, device=dev)
ms = timed(lambda: ops.act(mem, None, pool=1, slope=1.0, want_plain=False, want_split=True, parts=ops.SPLIT_F16))
print("act operand split     rows=%d: %.3f ms  %.0f GB/s moved" % (M * 14, ms, M * 14 * 256 * 8 / ms / 1e6))
y = torch.randn(M * 7, 128, device=dev)
wh, bh = torch.randn(3, 128, device=dev), torch.randn(3, device=dev)
ms = timed(lambda: ops.head(y, None, M, 7, wh, bh, n_sigmoid=1, slope=1.0))
print("head                  M=%d: %.3f ms  %.0f GB/s read" % (M, ms, M * 7 * 128 * 4 / ms / 1e6))
