"""Accuracy (vs fp64) and speed of pof_conv_tc_fwd on the DR-SPAAM layer shapes.  Needs a B200.

    python tools/conv_tc_check.py [small|layers|all]
"""
import os
import sys

import torch
import torch.nn.functional as F

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from planar_optical_flow_b200 import ops                      # noqa: E402
from planar_optical_flow_b200.engine import split_tf32        # noqa: E402

dev = torch.device("cuda")
F16 = os.environ.get("POF_CHECK_F16", "1") == "1"        # float16 parts (kind::f16, the engine's default) or TF32 parts
PARTS = ops.SPLIT_F16 if F16 else 2
OUT_SCALE = [1.0]


def w_split(w):
    """[Cout, Cin, taps] -> [taps, 2, Cout, Cin] (hi, lo): TF32 parts, or float16 parts of w * 2^s (as the engine does)."""
    if F16:
        import numpy as np
        s = 13 - int(np.ceil(np.log2(float(w.abs().max()))))
        OUT_SCALE[0] = 2.0 ** -s
        ws = w * 2.0 ** s
        hi = ws.half()
        lo = (ws - hi.float()).half()
        return torch.stack([hi, lo], dim=0).permute(3, 0, 1, 2).contiguous()
    OUT_SCALE[0] = 1.0
    hi, lo = split_tf32(w)
    lo, _ = split_tf32(lo)
    return torch.stack([hi, lo], dim=0).permute(3, 0, 1, 2).contiguous()


def reference(x, w, b, pad, pool, slope):
    y = F.conv1d(x.permute(0, 2, 1).double(), w.double(), b.double(), padding=pad)      # [M, Cout, Lout]
    if pool == 2:
        y = F.max_pool1d(y, 2)
    y = torch.where(y > 0, y, y * slope)
    return y.permute(0, 2, 1).reshape(-1, w.shape[0])


def run(M, LA, Cin, Cout, taps, pad, pool, seed=0, time_it=False, flags=0):
    g = torch.Generator(device="cpu").manual_seed(seed)
    x = (torch.randn(M, LA, Cin, generator=g).abs() * torch.rand(M, LA, Cin, generator=g)).to(dev)
    x = torch.where(torch.rand(M, LA, Cin, generator=g).to(dev) < 0.3, -0.1 * x, x)        # LeakyReLU-like
    w = (torch.randn(Cout, Cin, taps, generator=g) * (2.0 / (Cin * taps)) ** 0.5).to(dev)
    b = (torch.randn(Cout, generator=g) * 0.1).to(dev)
    Lout = LA if pad else LA - taps + 1
    _, a = ops.act(x.view(M * LA, Cin), None, pool=1, slope=1.0, want_plain=False, want_split=True, parts=PARTS)
    ws = w_split(w)
    plain, split = ops.conv_tc(a, ws, b, M, LA, Lout, taps, pad, pool=pool, slope=0.1, want_plain=True, want_split=True, chain_channels=flags, out_scale=OUT_SCALE[0])
    torch.cuda.synchronize()
    st = ops.conv_tc_status(dev)
    want = reference(x, w, b, pad, pool, 0.1)
    scale = float(want.abs().max())
    err = float((plain.double() - want).abs().max()) / scale
    err_split = float(((split[:, :Cout].double() + split[:, Cout:].double()) - want).abs().max()) / scale
    old = torch.backends.cudnn.allow_tf32
    torch.backends.cudnn.allow_tf32 = False
    y32 = F.conv1d(x.permute(0, 2, 1).contiguous(), w, b, padding=pad)
    if pool == 2:
        y32 = F.max_pool1d(y32, 2)
    y32 = F.leaky_relu(y32, 0.1).permute(0, 2, 1).reshape(-1, Cout)
    torch.backends.cudnn.allow_tf32 = old
    err32 = float((y32.double() - want).abs().max()) / scale
    msg = "%s chain=%-3d %s " % ("f16 " if F16 else "tf32", flags & 0xffff, "1cta" if flags & 0x10000 else "pair") + "M=%-6d LA=%-2d %3d->%3d taps=%-2d pool=%d  status=%d  err=%.2e (split %.2e)  cudnn-fp32 err=%.2e" % (
        M, LA, Cin, Cout, taps, pool, st, err, err_split, err32)
    if time_it:
        for _ in range(2):
            ops.conv_tc(a, ws, b, M, LA, Lout, taps, pad, pool=pool, slope=0.1, want_plain=False, want_split=True, chain_channels=flags, out_scale=OUT_SCALE[0])
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(5):
            ops.conv_tc(a, ws, b, M, LA, Lout, taps, pad, pool=pool, slope=0.1, want_plain=False, want_split=True, chain_channels=flags, out_scale=OUT_SCALE[0])
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / 5
        flops = 2.0 * M * Lout * Cin * Cout * taps
        msg += "  %.2f ms  %.1f TF/s fp32-equivalent (x3 split products)" % (ms, flops / ms / 1e9)
    print(msg, flush=True)
    return st


SMALL = [  # (M, LA, Cin, Cout, taps, pad, pool)
    (2, 56, 64, 64, 3, 1, 1), (5, 56, 64, 128, 3, 1, 2), (37, 28, 128, 128, 3, 1, 1), (9, 28, 128, 256, 3, 1, 2),
    (23, 14, 256, 256, 3, 1, 1), (40, 14, 256, 512, 3, 1, 2), (100, 7, 512, 256, 3, 1, 1), (61, 7, 256, 128, 3, 1, 1),
    (300, 14, 256, 128, 14, 0, 1), (3000, 14, 256, 512, 3, 1, 2)]
LAYERS = [(64, 64, 56, 1), (64, 128, 56, 2), (128, 128, 28, 1), (128, 128, 28, 1), (128, 256, 28, 2), (256, 256, 14, 1),
          (256, 256, 14, 1), (256, 512, 14, 2), (512, 256, 7, 1), (256, 128, 7, 1)]

def clocks_under_load(cfg, flags, seconds=3.0):
    """Loop one layer for a few seconds and sample nvidia-smi: the SM clock the tensor pipe really runs at."""
    import subprocess
    import threading
    import time

    M, LA, Cin, Cout, taps, pad, pool = cfg
    x = torch.randn(M * LA, Cin, device=dev)
    _, a = ops.act(x, None, pool=1, slope=1.0, want_plain=False, want_split=True, parts=PARTS)
    ws = w_split(torch.randn(Cout, Cin, taps, device=dev) * 0.05)
    b = torch.zeros(Cout, device=dev)
    samples = []
    stop = threading.Event()

    def sample():
        while not stop.is_set():
            out = subprocess.run(["nvidia-smi", "--query-gpu=clocks.sm,power.draw,clocks_event_reasons.sw_power_cap",
                                  "--format=csv,noheader,nounits", "-i", "0"], capture_output=True, text=True).stdout.strip()
            samples.append(out)
            time.sleep(0.1)

    th = threading.Thread(target=sample)
    th.start()
    t0 = time.time()
    n = 0
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    while time.time() - t0 < seconds:
        for _ in range(20):
            ops.conv_tc(a, ws, b, M, LA, LA, taps, pad, pool=pool, slope=0.1, want_plain=False, want_split=True, chain_channels=flags, out_scale=OUT_SCALE[0])
        n += 20
        torch.cuda.synchronize()
    e1.record()
    torch.cuda.synchronize()
    stop.set()
    th.join()
    print("sustained: %.3f ms per launch over %d launches; nvidia-smi (MHz, W, power cap): %s" % (
        e0.elapsed_time(e1) / n, n, " | ".join(samples[3::4])), flush=True)


if __name__ == "__main__":
    what = sys.argv[1] if len(sys.argv) > 1 else "small"
    if what == "clocks":
        clocks_under_load((64 * 1091, 14, 256, 512, 3, 1, 2), int(sys.argv[2], 0) if len(sys.argv) > 2 else 32)
    if what == "chains":
        for cfg in ((64 * 1091, 14, 256, 512, 3, 1, 2), (64 * 1091, 7, 512, 256, 3, 1, 1), (64 * 1091, 28, 128, 256, 3, 1, 2)):
            for flags in (64, 128, 256, 1024, 8192, 8192 | 0x10000):
                run(*cfg, time_it=True, flags=flags)
    if what == "layer":          # layer <cin> <cout> <L> <pool> [flags]: one launch per call, for ncu
        cin, cout, L, pool = (int(v) for v in sys.argv[2:6])
        flags = int(sys.argv[6], 0) if len(sys.argv) > 6 else 64
        for _ in range(3):
            run(64 * 1091, L, cin, cout, 3, 1, pool, time_it=False, flags=flags)
    if what == "one":
        flags = int(sys.argv[2], 0) if len(sys.argv) > 2 else 32
        for _ in range(3):
            run(64 * 1091, 14, 256, 512, 3, 1, 2, time_it=False, flags=flags)
    if what in ("small", "all"):
        for k, c in enumerate(SMALL):
            for flags in (64, 64 | 0x10000, 128):
                if run(*c, seed=k, flags=flags):
                    sys.exit("pipeline wait timed out")
    if what in ("layers", "all"):
        M = int(sys.argv[2]) if len(sys.argv) > 2 else 64 * 1091
        for cin, cout, L, pool in LAYERS:
            for flags in (64, 128, 128 | 0x10000):
                run(M, L, cin, cout, 3, 1, pool, time_it=True, flags=flags)
        for flags in (64, 64 | 0x10000):
            run(M, 14, 256, 128, 14, 0, 1, time_it=True, flags=flags)
