"""Run the attention-memory kernel at the bench's in-step shape (one chunk of 64 JRDB-shaped sequences, memory rows of
14 x 256 channels, float16 operand split emitted in the same pass) a few times: the target of an `ncu --set full` capture.

    python tools/gate_run.py [sequences=64] [reps=5] [split=1]
"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from planar_optical_flow_b200 import ops       # noqa: E402

B = int(sys.argv[1]) if len(sys.argv) > 1 else 64
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 5
split = (int(sys.argv[3]) if len(sys.argv) > 3 else 1) != 0
N, L, C, E, W = 1091, 14, 256, 128, 11
dev = torch.device("cuda")
g = torch.Generator(device="cuda").manual_seed(0)
x = torch.randn(B, N, L, C, device=dev, generator=g)
t = torch.randn(B, N, L, C, device=dev, generator=g)
ex = torch.randn(B, N, E, device=dev, generator=g) * 0.2
et = torch.randn(B, N, E, device=dev, generator=g) * 0.2
out = torch.empty_like(x)
ff = torch.empty(B, N, W, device=dev)
sp = torch.empty(B * N * L, 2 * C, dtype=torch.float16, device=dev) if split else None
status = ops.new_status(dev)
ev = [torch.cuda.Event(enable_timing=True) for _ in range(reps + 1)]
ev[0].record()
for r in range(reps):
    ops.gate_forward(x, t, ex, et, 0.5, W, out=out, feat_out=ff, split_out=sp, split_channels=C, status=status)
    ev[r + 1].record()
torch.cuda.synchronize()
ms = [ev[r].elapsed_time(ev[r + 1]) for r in range(reps)]
alg = B * N * 44076
tot = alg + (B * N * L * C * 2 * 2 if split else 0)
print("gate B=%d split=%d status=%d ms=%s  best %.3f ms: %.0f GB/s by the SURVEY 8d bytes (%.1f%% of 6547), %.0f GB/s incl. the operand split (%.1f%%)" % (
    B, split, ops.read_status(status), ["%.3f" % v for v in ms], min(ms), alg / min(ms) / 1e6, alg / min(ms) / 1e6 / 65.47,
    tot / min(ms) / 1e6, tot / min(ms) / 1e6 / 65.47))
