"""Time gate kernel build variants (threads / min blocks / prefetch depth) on the bench shape."""
import ctypes
import glob
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from planar_optical_flow_b200 import _lib  # noqa: E402

B, N, CL, E, W = int(os.environ.get("TB", 128)), 1091, 3584, 128, 11
dev = torch.device("cuda")
x = torch.randn(B, N, CL, device=dev)
t = torch.randn(B, N, CL, device=dev)
ex = torch.randn(B, N, E, device=dev) * 0.3
et = torch.randn(B, N, E, device=dev) * 0.3
out = torch.empty_like(x)
ff = torch.empty(B, N, W, device=dev)
bytes_alg = B * N * 44076
ref = None
variants = [] if os.environ.get("TUNE_ONLY_DEFAULT") else sorted(
    glob.glob(os.path.join(os.path.dirname(_lib.LIB_PATH), "_variants", "*.so")))
for path in variants + [_lib.LIB_PATH]:
    h = ctypes.CDLL(path)
    fn = h.pof_spaam_gate_fwd
    fn.restype = ctypes.c_int
    fn.argtypes = _lib.SIGNATURES["pof_spaam_gate_fwd"][1]
    args = [ctypes.c_void_p(v.data_ptr()) for v in (x, t, ex, et)] + [B, N, CL, E, W, 0.5,
            ctypes.c_void_p(out.data_ptr()), ctypes.c_void_p(ff.data_ptr()), None, None, 0, None, None]
    for _ in range(3):
        rc = fn(*args)
    torch.cuda.synchronize()
    assert rc == 0, rc
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(10):
        fn(*args)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 10
    if ref is None:
        ref = out.clone()
    same = bool(torch.equal(ref, out))
    print("%-40s %7.3f ms  %7.1f GB/s  (%.1f%% of 6547)  same=%s" % (os.path.basename(path), ms, bytes_alg / ms / 1e6,
                                                                    100 * bytes_alg / ms / 1e6 / 6547, same))
