"""Device-level operators: torch CUDA tensors in, torch CUDA tensors out.

Thin, allocation-only wrappers over the C ABI (include/pof.h).  torch is used
for device memory, streams and autograd plumbing; all arithmetic of the three
hot-path stages happens in libpof.so.  There is no CPU implementation: every
function raises if handed a CPU tensor.
"""
import ctypes
import os

import torch

from . import _lib
from ._lib import check, current_stream_ptr, require_cuda_tensor


def _ptr(t):
    return ctypes.c_void_p(t.data_ptr()) if t is not None else ctypes.c_void_p(0)


# --------------------------------------------------------------------------- cutout
def cutout(scans, phi, stride=1, centered=True, fixed=False, window_width=1.66, window_depth=1.0,
           num_cutout_pts=48, padding_val=29.99, area_mode=False, out=None, return_s_area=False,
           half_alpha=None, return_half_alpha=False, fast=False, exact_pieces=False):
    """Batched `scans_to_cutout` (reference: src/utils/utils.py:259-334).

    scans [B, S, N] float32 CUDA, phi [N] float32|float64 CUDA  ->  [B, M, S, P] float32,
    M = ceil(N / stride).  Each b is one independent reference call (its own `s_area`).
    `fast` selects POF_CUTOUT_FAST; `exact_pieces` runs the EXACT arithmetic on the first (piece-per-thread) kernel,
    which the tests hold against the default EXACT kernel bit for bit.
    """
    require_cuda_tensor(scans, "scans", torch.float32)
    require_cuda_tensor(phi, "phi")
    if scans.dim() != 3:
        raise ValueError("scans must be [B, S, N] (got %s)" % (tuple(scans.shape),))
    if phi.dtype not in (torch.float32, torch.float64):
        raise TypeError("phi must be float32 or float64 (got %s)" % phi.dtype)
    B, S, N = scans.shape
    if phi.numel() != N:
        raise ValueError("phi has %d angles for %d-point scans" % (phi.numel(), N))
    P = int(num_cutout_pts)
    M = (N + stride - 1) // stride
    dev = scans.device
    with torch.cuda.device(dev):
        if out is None:
            out = torch.empty((B, M, S, P), dtype=torch.float32, device=dev)
        else:
            require_cuda_tensor(out, "out", torch.float32)
            if tuple(out.shape) != (B, M, S, P):
                raise ValueError("out must be %s" % ((B, M, S, P),))
        L = _lib.lib()
        ws_bytes = L.pof_cutout_ws_bytes(B)
        ws = torch.empty(max(ws_bytes, 8), dtype=torch.uint8, device=dev)
        s_area = torch.empty(max(B, 1), dtype=torch.int32, device=dev) if return_s_area else None
        if half_alpha is not None:
            require_cuda_tensor(half_alpha, "half_alpha", torch.float32)
            if tuple(half_alpha.shape) != (B, S, M):
                raise ValueError("half_alpha must be [B, S, M] = %s" % ((B, S, M),))
        ha_out = torch.empty((B, S, M), dtype=torch.float32, device=dev) if return_half_alpha else None
        check(L.pof_cutout_fwd(_ptr(scans), _ptr(phi), int(phi.dtype == torch.float64), B, S, N, int(stride), P,
                               float(window_width), float(window_depth), float(padding_val),
                               int(bool(fixed)), int(bool(centered)), int(bool(area_mode)), 1 if fast else (2 if exact_pieces else 0),
                               _ptr(out), _ptr(s_area), _ptr(half_alpha), _ptr(ha_out),
                               _ptr(ws), ws_bytes, current_stream_ptr(dev)),
              "pof_cutout_fwd")
    extra = ([s_area[:B]] if return_s_area else []) + ([ha_out] if return_half_alpha else [])
    return (out, *extra) if extra else out


def cutout_original(scans, angle_incre, angle_incre_is_f32=False, fixed=True, centered=True, window_width=1.66, window_depth=1.0,
                    num_cutout_pts=48, padding_val=29.99):
    """Batched `scans_to_cutout_original` (utils.py:423-489): scans [B, S, N] float32 CUDA -> [B, N, S, P]."""
    require_cuda_tensor(scans, "scans", torch.float32)
    if scans.dim() != 3:
        raise ValueError("scans must be [B, S, N] (got %s)" % (tuple(scans.shape),))
    B, S, N = scans.shape
    dev = scans.device
    with torch.cuda.device(dev):
        out = torch.empty((B, N, S, int(num_cutout_pts)), dtype=torch.float32, device=dev)
        check(_lib.lib().pof_cutout_original_fwd(_ptr(scans), B, S, N, float(angle_incre), int(bool(angle_incre_is_f32)),
                                                 int(num_cutout_pts), float(window_width), float(window_depth), float(padding_val),
                                                 int(bool(fixed)), int(bool(centered)), _ptr(out), current_stream_ptr(dev)),
              "pof_cutout_original_fwd")
    return out


def polar_grid(scans, min_range=0.0, max_range=30.0, range_bin_size=1.0, tsdf_clip=1.0, normalize=True):
    """`scans_to_polar_grid` (utils.py:492-531): scans [S, N] float32 CUDA -> [S, R, N]."""
    require_cuda_tensor(scans, "scans", torch.float32)
    S, N = scans.shape
    R = int((max_range - min_range) / range_bin_size) + 1
    dev = scans.device
    with torch.cuda.device(dev):
        out = torch.empty((S, R, N), dtype=torch.float32, device=dev)
        check(_lib.lib().pof_polar_grid_fwd(_ptr(scans), S, N, float(min_range), float(max_range), float(range_bin_size),
                                            float(tsdf_clip), int(bool(normalize)), _ptr(out), current_stream_ptr(dev)),
              "pof_polar_grid_fwd")
    return out


# --------------------------------------------------------------------------- gate
def gate_forward(x, tmpl, emb_x, emb_t, alpha, window, want_weights=False, out=None, feat_out=None, split_out=None,
                 split_channels=0, status=None):
    """Windowed attention memory update (reference: dr_spaam.py:183-215).

    x, tmpl [B, N, C, L] (or [B, N, CL]) float32; emb_* [B, N, E] float32; window = 2*hw+1.
    Returns (out_tmpl like x, feat_fused [B, N, W], attn_w [B, N, W] or None).
    `split_out` (float16 [B * N * CL / split_channels, 2 * split_channels]) additionally receives the new memory as the
    [hi | lo] operand rows of `conv_tc`; `status` is the caller's device status word (see `new_status`).
    """
    for name, t in (("x", x), ("tmpl", tmpl), ("emb_x", emb_x), ("emb_t", emb_t)):
        require_cuda_tensor(t, name, torch.float32)
    if x.shape != tmpl.shape:
        raise ValueError("x %s and template %s differ" % (tuple(x.shape), tuple(tmpl.shape)))
    B, N = x.shape[0], x.shape[1]
    CL = x[0, 0].numel()
    E = emb_x.shape[-1]
    if tuple(emb_x.shape) != (B, N, E) or tuple(emb_t.shape) != (B, N, E):
        raise ValueError("embeddings must be [B, N, E]")
    dev = x.device
    with torch.cuda.device(dev):
        if out is None:
            out = torch.empty_like(x)
        if feat_out is None:
            feat = torch.empty((B, N, window), dtype=torch.float32, device=dev)
        else:
            feat = require_cuda_tensor(feat_out, "feat_out", torch.float32)
            if tuple(feat.shape) != (B, N, window):
                raise ValueError("feat_out must be [B, N, W] = %s" % ((B, N, window),))
        attn = torch.empty((B, N, window), dtype=torch.float32, device=dev) if want_weights else None
        if split_out is not None:
            require_cuda_tensor(split_out, "split_out", torch.float16)
            if split_out.numel() != 2 * B * N * CL:
                raise ValueError("split_out must hold 2 * B * N * CL = %d halves (got %d)" % (2 * B * N * CL, split_out.numel()))
        check(_lib.lib().pof_spaam_gate_fwd(_ptr(x), _ptr(tmpl), _ptr(emb_x), _ptr(emb_t), B, N, CL, E, int(window),
                                            float(alpha), _ptr(out), _ptr(feat), _ptr(attn), _ptr(split_out),
                                            int(split_channels), _ptr(status), current_stream_ptr(dev)),
              "pof_spaam_gate_fwd")
    return out, feat, attn


def gate_backward(tmpl, emb_x, emb_t, attn_w, g_out, g_feat, alpha, window):
    """Gradients of `gate_forward` w.r.t. (x, tmpl, emb_x, emb_t)."""
    B, N = tmpl.shape[0], tmpl.shape[1]
    CL = tmpl[0, 0].numel()
    E = emb_x.shape[-1]
    dev = tmpl.device
    g_out = g_out.contiguous()
    g_feat = g_feat.contiguous() if g_feat is not None else None
    with torch.cuda.device(dev):
        g_x = torch.empty_like(tmpl)
        g_t = torch.empty_like(tmpl)
        g_ex = torch.empty_like(emb_x)
        g_et = torch.empty_like(emb_t)
        L = _lib.lib()
        ws_bytes = L.pof_spaam_gate_bwd_ws_bytes(B, N, int(window))
        ws = torch.empty(max(ws_bytes, 4), dtype=torch.uint8, device=dev)
        check(L.pof_spaam_gate_bwd(_ptr(tmpl), _ptr(emb_x), _ptr(emb_t), _ptr(attn_w), _ptr(g_out), _ptr(g_feat),
                                   B, N, CL, E, int(window), float(alpha),
                                   _ptr(g_x), _ptr(g_t), _ptr(g_ex), _ptr(g_et), _ptr(ws), ws_bytes,
                                   current_stream_ptr(dev)),
              "pof_spaam_gate_bwd")
    return g_x, g_t, g_ex, g_et


class _GateFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, tmpl, emb_x, emb_t, alpha, window):
        x, tmpl = x.contiguous(), tmpl.contiguous()
        emb_x, emb_t = emb_x.contiguous(), emb_t.contiguous()
        need = any(ctx.needs_input_grad[:4])
        out, feat, attn = gate_forward(x, tmpl, emb_x, emb_t, alpha, window, want_weights=need)
        if need:
            ctx.save_for_backward(tmpl, emb_x, emb_t, attn)
        ctx.alpha, ctx.window = alpha, window
        return out, feat

    @staticmethod
    def backward(ctx, g_out, g_feat):
        tmpl, emb_x, emb_t, attn = ctx.saved_tensors
        g_x, g_t, g_ex, g_et = gate_backward(tmpl, emb_x, emb_t, attn, g_out, g_feat, ctx.alpha, ctx.window)
        return g_x, g_t, g_ex, g_et, None, None


def gate(x, tmpl, emb_x, emb_t, alpha, window):
    """Differentiable fused gate: returns (out_tmpl, feat_fused)."""
    return _GateFn.apply(x, tmpl, emb_x, emb_t, float(alpha), int(window))


# --------------------------------------------------------------------------- backbone glue
SPLIT_F16 = 16          # include/pof.h POF_SPLIT_F16


def _split_buffer(rows, C, parts, dev):
    if parts == SPLIT_F16:
        return torch.empty((rows, 2 * C), dtype=torch.float16, device=dev)
    return torch.empty((rows, parts * C), dtype=torch.float32, device=dev)


def new_status(device):
    """A zeroed device status word for `conv_tc` / `act` / `conv_first` / `gate_forward` (see `read_status`)."""
    return torch.zeros(1, dtype=torch.int32, device=device)


def read_status(status):
    """Synchronise, return the status word and clear it: 0 = fine, 16 = an activation left the float16 range of the
    float16 operand split (|x| > 65504 or NaN), 32 = a staged copy of the attention kernel never landed, anything
    else = a pipeline wait of the tcgen05 convolution timed out (results invalid)."""
    code = int(status.item())
    if code:
        status.zero_()
    return code


def act(y, bias=None, pool=1, slope=0.1, want_plain=True, want_split=False, parts=3, status=None):
    """Channels-last activations y [rows, C] -> (+bias) LeakyReLU, max over `pool` consecutive rows.

    Returns (plain [rows/pool, C] or None, split [rows/pool, parts*C] or None); parts = 3: [hi | lo | hi]
    (cuDNN operand), parts = 2: [hi | lo] (operand of `conv_tc`), parts = SPLIT_F16: [hi | lo] in float16
    (operand of `conv_tc` on float16 operands)."""
    require_cuda_tensor(y, "y", torch.float32)
    rows, C = y.shape
    dev = y.device
    with torch.cuda.device(dev):
        plain = torch.empty((rows // pool, C), dtype=torch.float32, device=dev) if want_plain else None
        split = _split_buffer(rows // pool, C, parts, dev) if want_split else None
        check(_lib.lib().pof_act_fwd(_ptr(y), _ptr(bias), rows, C, int(pool), float(slope), _ptr(plain), _ptr(split),
                                     int(parts), _ptr(status), current_stream_ptr(dev)), "pof_act_fwd")
    return plain, split


def conv_first(cutouts, weight, bias, slope=0.1, want_plain=False, want_split=True, parts=3, status=None):
    """cutouts [M, P], weight [C, 3], bias [C] -> first conv layer + LeakyReLU, channels-last [M*P, C] / [M*P, 3C]."""
    require_cuda_tensor(cutouts, "cutouts", torch.float32)
    require_cuda_tensor(weight, "weight", torch.float32)
    require_cuda_tensor(bias, "bias", torch.float32)
    M, P = cutouts.shape
    C = weight.shape[0]
    dev = cutouts.device
    with torch.cuda.device(dev):
        plain = torch.empty((M * P, C), dtype=torch.float32, device=dev) if want_plain else None
        split = _split_buffer(M * P, C, parts, dev) if want_split else None
        check(_lib.lib().pof_conv_first_fwd(_ptr(cutouts), _ptr(weight), _ptr(bias), M, P, C, float(slope), _ptr(plain),
                                            _ptr(split), int(parts), _ptr(status), current_stream_ptr(dev)), "pof_conv_first_fwd")
    return plain, split


_conv_tc_status = {}


def conv_tc_status(device):
    """The status word shared by `conv_tc` calls on `device` that did not bring their own (0 = every pipeline wait
    completed); reading clears it, so one failed launch is reported once."""
    dev = torch.device(device)
    if dev.type == "cuda" and dev.index is None:          # "cuda" means the current device, as everywhere in torch
        dev = torch.device("cuda", torch.cuda.current_device())
    t = _conv_tc_status.get(dev)
    return 0 if t is None else read_status(t)


def conv_tc(a_split, w_split, bias, Mcut, LA, Lout, taps, pad, pool=1, slope=0.1, want_plain=False, want_split=True, chain_channels=0,
            out_scale=1.0, status=None, plain_out=None):
    """fp32-accurate convolution / whole-row GEMM on tcgen05 (csrc/pof_conv_tc.cu).

    a_split [Mcut*LA, 2*Cin] = [hi | lo] rows, w_split [taps, 2, Cout, Cin]  ->
    (plain [Mcut*Lout/pool, Cout] or None, split [.., 2*Cout] or None).
    float32 operands hold TF32 parts (kind::tf32); float16 operands run on kind::f16 at twice the rate:
    `w_split` then holds the parts of weight * 2^s and `out_scale` = 2^-s, and `split` is float16."""
    f16 = a_split.dtype == torch.float16
    dt = torch.float16 if f16 else torch.float32
    require_cuda_tensor(a_split, "a_split", dt)
    require_cuda_tensor(w_split, "w_split", dt)
    Cin = a_split.shape[-1] // 2
    if w_split.dim() != 4 or w_split.shape[0] != taps or w_split.shape[1] != 2 or w_split.shape[3] != Cin:
        raise ValueError("w_split must be [taps, 2, Cout, Cin] (got %s for Cin = %d)" % (tuple(w_split.shape), Cin))
    if a_split.numel() != Mcut * LA * 2 * Cin:
        raise ValueError("a_split has %d elements, expected %d x %d x %d" % (a_split.numel(), Mcut, LA, 2 * Cin))
    if not f16 and out_scale != 1.0:
        raise ValueError("out_scale belongs to the float16 operand form")
    Cout = w_split.shape[2]
    dev = a_split.device
    if not chain_channels:
        chain_channels = int(os.environ.get("POF_CONV_TC_CHAIN", "0"), 0)      # tuning aid; 0 = the library default
    with torch.cuda.device(dev):
        if status is None:
            status = _conv_tc_status.get(dev)
            if status is None:
                status = _conv_tc_status[dev] = new_status(dev)
        rows = Mcut * Lout // pool
        if plain_out is not None:                     # write the fp32 rows straight into the caller's buffer (a view is fine)
            plain = require_cuda_tensor(plain_out, "plain_out", torch.float32)
            if plain.numel() != rows * Cout:
                raise ValueError("plain_out must hold %d x %d values (got %d)" % (rows, Cout, plain.numel()))
        else:
            plain = torch.empty((rows, Cout), dtype=torch.float32, device=dev) if want_plain else None
        split = torch.empty((rows, 2 * Cout), dtype=dt, device=dev) if want_split else None
        if f16:
            check(_lib.lib().pof_conv_tc_f16_fwd(_ptr(a_split), _ptr(w_split), _ptr(bias), Mcut, int(LA), int(Lout), Cin, Cout,
                                                 int(taps), int(pad), int(pool), float(slope), float(out_scale), _ptr(plain),
                                                 _ptr(split), _ptr(status), int(chain_channels), current_stream_ptr(dev)),
                  "pof_conv_tc_f16_fwd")
        else:
            check(_lib.lib().pof_conv_tc_fwd(_ptr(a_split), _ptr(w_split), _ptr(bias), Mcut, int(LA), int(Lout), Cin, Cout,
                                             int(taps), int(pad), int(pool), float(slope), _ptr(plain), _ptr(split),
                                             _ptr(status), int(chain_channels), current_stream_ptr(dev)), "pof_conv_tc_fwd")
    return plain, split


def head(y, bias, M, L, w_head, b_head, n_sigmoid, slope=0.1, out=None, out_rest=None):
    """y [M*L, C] raw last-conv output -> +bias, LeakyReLU, mean over L, heads [H, C] (+ sigmoid on the first n_sigmoid) -> [M, H].

    With `out` [M, n_sigmoid] and `out_rest` [M, H - n_sigmoid] the two groups of heads (classes | regression) go to
    the caller's own tensors instead of one [M, H] matrix."""
    require_cuda_tensor(y, "y", torch.float32)
    require_cuda_tensor(w_head, "w_head", torch.float32)
    require_cuda_tensor(b_head, "b_head", torch.float32)
    C = y.shape[-1]
    H = w_head.shape[0]
    dev = y.device
    with torch.cuda.device(dev):
        if out_rest is not None:
            require_cuda_tensor(out, "out", torch.float32)
            require_cuda_tensor(out_rest, "out_rest", torch.float32)
            if out.numel() != M * n_sigmoid or out_rest.numel() != M * (H - n_sigmoid):
                raise ValueError("out / out_rest must hold M x n_sigmoid and M x (H - n_sigmoid) values")
        else:
            out = torch.empty((M, H), dtype=torch.float32, device=dev)
        check(_lib.lib().pof_head_fwd(_ptr(y), _ptr(bias), M, int(L), C, float(slope), _ptr(w_head), _ptr(b_head), H,
                                      int(n_sigmoid), _ptr(out), _ptr(out_rest), current_stream_ptr(dev)), "pof_head_fwd")
    return out if out_rest is None else (out, out_rest)


# --------------------------------------------------------------------------- training: batch-norm + LeakyReLU (+ pool)
class _BnActPoolFn(torch.autograd.Function):
    """y [M, C, 1, L] (channels-last memory: rows = M * L, C contiguous) -> [M, C, 1, L / pool]."""

    @staticmethod
    def forward(ctx, y, gamma, beta, running_mean, running_var, momentum, eps, slope, pool, groups):
        require_cuda_tensor(gamma, "gamma", torch.float32)
        require_cuda_tensor(beta, "beta", torch.float32)
        if y.dim() != 4 or y.shape[2] != 1 or y.dtype != torch.float32 or not y.is_cuda:
            raise ValueError("bn_act_pool takes float32 CUDA activations [M, C, 1, L]")
        if not y.is_contiguous(memory_format=torch.channels_last):
            y = y.contiguous(memory_format=torch.channels_last)
        M, C, _, Lr = y.shape
        if pool not in (1, 2) or Lr % pool:
            raise ValueError("pool must be 1, or 2 with an even length (got pool=%d, L=%d)" % (pool, Lr))
        rows = M * Lr
        if groups < 1 or M % groups:
            raise ValueError("the %d cutouts do not split into %d groups" % (M, groups))
        dev = y.device
        with torch.cuda.device(dev):
            sums = torch.empty(2 * groups * C, dtype=torch.float64, device=dev)
            z = torch.empty((M, C, 1, Lr // pool), dtype=torch.float32, device=dev, memory_format=torch.channels_last)
            mean = torch.empty(groups * C, dtype=torch.float32, device=dev)
            invstd = torch.empty(groups * C, dtype=torch.float32, device=dev)
            L = _lib.lib()
            stream = current_stream_ptr(dev)
            check(L.pof_bn_act_stats(_ptr(y), rows, C, int(groups), _ptr(sums), stream), "pof_bn_act_stats")
            check(L.pof_bn_act_fwd(_ptr(y), _ptr(sums), _ptr(gamma), _ptr(beta), rows, C, int(groups), int(pool), float(eps), float(slope),
                                   float(momentum), _ptr(z), _ptr(mean), _ptr(invstd), _ptr(running_mean), _ptr(running_var), stream),
                  "pof_bn_act_fwd")
        ctx.save_for_backward(y, gamma, beta, mean, invstd)
        ctx.cfg = (float(slope), int(pool), int(groups))
        return z

    @staticmethod
    def backward(ctx, dz):
        y, gamma, beta, mean, invstd = ctx.saved_tensors
        slope, pool, groups = ctx.cfg
        M, C, _, Lr = y.shape
        dev = y.device
        dz = dz.contiguous(memory_format=torch.channels_last)
        with torch.cuda.device(dev):
            dx = torch.empty_like(y, memory_format=torch.channels_last)
            dgamma = torch.empty(C, dtype=torch.float32, device=dev)
            dbeta = torch.empty(C, dtype=torch.float32, device=dev)
            sums = torch.empty(2 * groups * C, dtype=torch.float64, device=dev)
            check(_lib.lib().pof_bn_act_bwd(_ptr(y), _ptr(dz), _ptr(mean), _ptr(invstd), _ptr(gamma), _ptr(beta), M * Lr, C, groups, pool,
                                            slope, _ptr(sums), _ptr(dx), _ptr(dgamma), _ptr(dbeta), current_stream_ptr(dev)),
                  "pof_bn_act_bwd")
        return dx, dgamma, dbeta, None, None, None, None, None, None, None


class _ConvFirstFn(torch.autograd.Function):
    """Conv1d(1 -> C, k = 3, p = 1) without bias on cutouts [M, P] -> channels-last activations [M, C, 1, P]."""

    @staticmethod
    def forward(ctx, cutouts, weight):
        require_cuda_tensor(cutouts, "cutouts", torch.float32)
        M, P = cutouts.shape
        C = weight.shape[0]
        w = weight.detach().reshape(C, 3).contiguous()
        zero = torch.zeros(C, dtype=torch.float32, device=cutouts.device)
        plain, _ = conv_first(cutouts, w, zero, slope=1.0, want_plain=True, want_split=False)       # [M * P, C]
        ctx.save_for_backward(cutouts)
        ctx.w_shape = tuple(weight.shape)
        return plain.view(M, 1, P, C).permute(0, 3, 1, 2)                                           # [M, C, 1, P], channels-last memory

    @staticmethod
    def backward(ctx, dy):
        (cutouts,) = ctx.saved_tensors
        M, P = cutouts.shape
        C = ctx.w_shape[0]
        dev = cutouts.device
        dy = dy.contiguous(memory_format=torch.channels_last)
        with torch.cuda.device(dev):
            sums = torch.empty(3 * C, dtype=torch.float64, device=dev)
            dw = torch.empty((C, 3), dtype=torch.float32, device=dev)
            check(_lib.lib().pof_conv_first_wgrad(_ptr(dy), _ptr(cutouts), M, P, C, _ptr(sums), _ptr(dw), current_stream_ptr(dev)),
                  "pof_conv_first_wgrad")
        return None, dw.view(ctx.w_shape)


def conv_first_train(cutouts, weight):
    """The first layer's convolution for the training branch (no bias: it cancels under batch statistics), differentiable in
    the weight; the cutouts are data and get no gradient.  cutouts [M, P] float32 CUDA, weight [C, 1, 3]."""
    return _ConvFirstFn.apply(cutouts, weight)


def bn_act_pool(y, gamma, beta, running_mean=None, running_var=None, momentum=0.1, eps=1e-5, slope=0.1, pool=1, groups=1):
    """Training-mode BatchNorm (batch statistics, running statistics updated in place) + LeakyReLU + max-pool over `pool`
    consecutive positions, as one differentiable operator on channels-last activations y [M, C, 1, L] (csrc/pof_bnact.cu).
    `groups` > 1: the M cutouts are `groups` equal consecutive blocks (the scans of a training sample), each normalised with
    its own batch statistics exactly as `groups` separate calls would, the running statistics updated once per block in order."""
    return _BnActPoolFn.apply(y, gamma, beta, running_mean, running_var, momentum, eps, slope, pool, groups)


# --------------------------------------------------------------------------- patch correlation (prototype)
class _PatchCorrFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, f1, f2, kernel_size, max_displacement):
        f1, f2 = f1.contiguous(), f2.contiguous()
        require_cuda_tensor(f1, "feat1", torch.float32)
        require_cuda_tensor(f2, "feat2", torch.float32)
        if f1.shape != f2.shape or f1.dim() != 3:
            raise ValueError("feat1 and feat2 must both be [B, C, N] (got %s, %s)" % (tuple(f1.shape), tuple(f2.shape)))
        B, C, N = f1.shape
        dev = f1.device
        with torch.cuda.device(dev):
            out = torch.empty((B, 2 * max_displacement + 1, N), dtype=torch.float32, device=dev)
            check(_lib.lib().pof_patch_corr_fwd(_ptr(f1), _ptr(f2), B, C, N, int(kernel_size), int(max_displacement), _ptr(out),
                                                current_stream_ptr(dev)), "pof_patch_corr_fwd")
        ctx.save_for_backward(f1, f2)
        ctx.cfg = (int(kernel_size), int(max_displacement))
        return out

    @staticmethod
    def backward(ctx, g):
        f1, f2 = ctx.saved_tensors
        B, C, N = f1.shape
        dev = f1.device
        g = g.contiguous()
        with torch.cuda.device(dev):
            g1, g2 = torch.empty_like(f1), torch.empty_like(f2)
            check(_lib.lib().pof_patch_corr_bwd(_ptr(f1), _ptr(f2), _ptr(g), B, C, N, ctx.cfg[0], ctx.cfg[1], _ptr(g1), _ptr(g2),
                                                current_stream_ptr(dev)), "pof_patch_corr_bwd")
        return g1, g2, None, None


def patch_corr(feat1, feat2, kernel_size=3, max_displacement=5):
    """Windowed patch correlation [B, C, N] x [B, C, N] -> [B, 2*max_displacement+1, N] (prototype.py:118-156), differentiable."""
    return _PatchCorrFn.apply(feat1, feat2, kernel_size, max_displacement)


# --------------------------------------------------------------------------- nms
def nms_centers(scan, phi, cls, reg, min_dist=0.5):
    """Batched `nms_predicted_center` (reference: src/utils/utils.py:535-571).

    scan [B, N] float32|float64, phi [N] float32|float64, cls [B, N] float32 (post-sigmoid),
    reg [B, N, 2] float32; all CUDA.  Returns a dict of device tensors:
      order [B,N] i32, keep_idx [B,N] i32, n_keep [B] i32, instance_mask [B,N] i32,
      det_xy [B,N,2] f64, det_cls [B,N] f32  (first n_keep[b] rows of the last three are valid).
    """
    require_cuda_tensor(scan, "scan")
    require_cuda_tensor(phi, "phi")
    require_cuda_tensor(cls, "cls", torch.float32)
    require_cuda_tensor(reg, "reg", torch.float32)
    for name, t in (("scan", scan), ("phi", phi)):
        if t.dtype not in (torch.float32, torch.float64):
            raise TypeError("%s must be float32 or float64" % name)
    B, N = scan.shape
    if tuple(cls.shape) != (B, N) or tuple(reg.shape) != (B, N, 2) or phi.numel() != N:
        raise ValueError("shape mismatch: scan %s cls %s reg %s phi %s" %
                         (tuple(scan.shape), tuple(cls.shape), tuple(reg.shape), tuple(phi.shape)))
    dev = scan.device
    with torch.cuda.device(dev):
        i32 = dict(dtype=torch.int32, device=dev)
        res = {
            "order": torch.empty((B, N), **i32),
            "keep_idx": torch.empty((B, N), **i32),
            "n_keep": torch.empty((B,), **i32),
            "instance_mask": torch.empty((B, N), **i32),
            "det_xy": torch.empty((B, N, 2), dtype=torch.float64, device=dev),
            "det_cls": torch.empty((B, N), dtype=torch.float32, device=dev),
        }
        L = _lib.lib()
        ws_bytes = L.pof_nms_ws_bytes(B, N)
        ws = torch.empty(max(ws_bytes, 4), dtype=torch.uint8, device=dev)
        check(L.pof_nms_centers(_ptr(scan), int(scan.dtype == torch.float64), _ptr(phi), int(phi.dtype == torch.float64),
                                _ptr(cls), _ptr(reg), B, N, float(min_dist),
                                _ptr(res["order"]), _ptr(res["keep_idx"]), _ptr(res["n_keep"]), _ptr(res["instance_mask"]),
                                _ptr(res["det_xy"]), _ptr(res["det_cls"]), _ptr(ws), ws_bytes, current_stream_ptr(dev)),
              "pof_nms_centers")
    return res
