"""How fast can cuDNN run the DR-SPAAM conv stack in IEEE fp32 (and TF32)?  Decides engine settings."""
import sys
import time

import torch
import torch.nn.functional as F

dev = torch.device("cuda")
M = int(sys.argv[1]) if len(sys.argv) > 1 else 128 * 1091          # cutouts per chunk

LAYERS = [  # (cin, cout, L)
    (1, 64, 56), (64, 64, 56), (64, 128, 56),
    (128, 128, 28), (128, 128, 28), (128, 256, 28),
    (256, 256, 14), (256, 256, 14), (256, 512, 14),
    (512, 256, 7), (256, 128, 7)]


def timeit(fn, n=5):
    fn(); fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n


def main():
    for tf32 in (False, True):
        for bench in (False, True):
            torch.backends.cudnn.allow_tf32 = tf32
            torch.backends.cuda.matmul.allow_tf32 = tf32
            torch.backends.cudnn.benchmark = bench
            total = {"ncl": 0.0, "nlc(channels_last 2d)": 0.0, "gemm3": 0.0}
            print("== tf32=%s cudnn.benchmark=%s, M=%d cutouts" % (tf32, bench, M))
            for cin, cout, L in LAYERS:
                flops = 2.0 * M * L * cin * cout * 3
                x = torch.randn(M, cin, L, device=dev)
                w = torch.randn(cout, cin, 3, device=dev) * 0.05
                b = torch.randn(cout, device=dev)
                with torch.no_grad():
                    t1 = timeit(lambda: F.conv1d(x, w, b, padding=1))
                    x4 = x.unsqueeze(2).contiguous(memory_format=torch.channels_last)
                    w4 = w.unsqueeze(2).contiguous(memory_format=torch.channels_last)
                    t2 = timeit(lambda: F.conv2d(x4, w4, b, padding=(0, 1)))
                    # conv as 3 shifted GEMMs over a zero-padded channels-last buffer [M, L+2, cin]
                    xp = torch.zeros(M, L + 2, cin, device=dev)
                    xp[:, 1:-1] = x.transpose(1, 2)
                    wk = [w[:, :, k].t().contiguous() for k in range(3)]          # [cin, cout]
                    flat = xp.view(M * (L + 2), cin)
                    out = torch.empty(M * (L + 2) - 2, cout, device=dev)

                    def gemm3():
                        torch.addmm(b, flat[:-2], wk[0], out=out)
                        out.addmm_(flat[1:-1], wk[1])
                        out.addmm_(flat[2:], wk[2])
                    t3 = timeit(gemm3)
                total["ncl"] += t1
                total["nlc(channels_last 2d)"] += t2
                total["gemm3"] += t3
                print("  %3d->%3d L=%2d  %6.1f GFLOP | conv1d NCL %7.2f ms %6.1f TF/s | conv2d NHWC %7.2f ms %6.1f TF/s | 3xGEMM %7.2f ms %6.1f TF/s"
                      % (cin, cout, L, flops / 1e9, t1, flops / t1 / 1e9, t2, flops / t2 / 1e9, t3, flops / t3 / 1e9))
                del x, w, x4, w4, xp, flat, out
                torch.cuda.empty_cache()
            print("  totals (ms):", {k: round(v, 1) for k, v in total.items()})


main()
