"""Host logic of the channels-last backbone (engine._ChannelsLastBackbone), on CPU.

The three libpof glue kernels it calls are replaced by torch restatements HERE (test
infrastructure, never shipped) so that the weight plumbing — BN folding, NHWC weight
layout, the [w_hi | w_hi | w_lo] concatenation against [hi | lo | hi] operands, the
permuted embedding GEMM, the stacked heads — can be checked against the oracle
network without a GPU.  The kernels themselves are checked on the B200 in
tests/test_gpu_parity.py.
"""
import numpy as np
import pytest
import torch
import torch.nn.functional as F

from oracle import model as omodel
from planar_optical_flow_b200 import engine, ops
from planar_optical_flow_b200.model import SpatialDROW
from tests.helpers import assert_rel


def _tf32_round(x):
    hi, _ = engine.split_tf32(x)
    return hi


def _emit(v, want_plain, want_split, parts=3):
    split = None
    if want_split and parts == ops.SPLIT_F16:
        hi = v.half()
        split = torch.cat([hi, (v - hi.float()).half()], dim=1)
    elif want_split:
        hi = _tf32_round(v)
        split = torch.cat([hi, v - hi, hi], dim=1) if parts == 3 else torch.cat([hi, _tf32_round(v - hi)], dim=1)
    return (v if want_plain else None), split


def fake_act(y, bias=None, pool=1, slope=0.1, want_plain=True, want_split=False, parts=3, status=None):
    rows, C = y.shape
    v = y.view(rows // pool, pool, C).max(dim=1).values
    if bias is not None:
        v = v + bias
    v = torch.where(v > 0, v, v * slope)
    return _emit(v, want_plain, want_split, parts)


def fake_conv_first(cutouts, weight, bias, slope=0.1, want_plain=False, want_split=True, parts=3, status=None):
    M, P = cutouts.shape
    v = F.conv1d(cutouts.view(M, 1, P), weight.view(-1, 1, 3), bias, padding=1)        # [M, C, P]
    v = F.leaky_relu(v, slope).permute(0, 2, 1).reshape(M * P, -1)
    return _emit(v, want_plain, want_split, parts)


def fake_conv_tc(a_split, w_split, bias, Mcut, LA, Lout, taps, pad, pool=1, slope=0.1, want_plain=False, want_split=True,
                 chain_channels=0, out_scale=1.0, status=None, plain_out=None):
    """The definition pof_conv_tc_fwd / pof_conv_tc_f16_fwd implement, with the three split products in plain fp32."""
    parts = ops.SPLIT_F16 if a_split.dtype == torch.float16 else 2
    a_split, w_split = a_split.float(), w_split.float()
    cin = a_split.shape[1] // 2
    hi, lo = a_split[:, :cin].view(Mcut, LA, cin).permute(0, 2, 1), a_split[:, cin:].view(Mcut, LA, cin).permute(0, 2, 1)
    w_hi, w_lo = w_split[:, 0].permute(1, 2, 0), w_split[:, 1].permute(1, 2, 0)          # [Cout, Cin, taps]
    y = F.conv1d(lo, w_hi, None, padding=pad) + F.conv1d(hi, w_lo, None, padding=pad) + F.conv1d(hi, w_hi, None, padding=pad)
    y = y * out_scale
    if bias is not None:
        y = y + bias[None, :, None]
    if pool == 2:
        y = F.max_pool1d(y, 2)
    y = torch.where(y > 0, y, y * slope).permute(0, 2, 1).reshape(Mcut * Lout // pool, -1)
    plain, split = _emit(y, want_plain or plain_out is not None, want_split, parts)
    if plain_out is not None:
        plain_out.view(plain.shape).copy_(plain)
        plain = plain_out
    return plain, split


def fake_head(y, bias, M, L, w_head, b_head, n_sigmoid, slope=0.1, out=None, out_rest=None):
    if bias is None:
        bias = torch.zeros(y.shape[1])
    dst_cls = out
    v = F.leaky_relu(y + bias, slope).view(M, L, -1).mean(dim=1)
    out = v @ w_head.t() + b_head
    out[:, :n_sigmoid] = torch.sigmoid(out[:, :n_sigmoid])
    if out_rest is None:
        return out
    dst_cls.view(M, n_sigmoid).copy_(out[:, :n_sigmoid])
    out_rest.view(M, -1).copy_(out[:, n_sigmoid:])
    return dst_cls, out_rest


@pytest.fixture
def cpu_glue(monkeypatch):
    monkeypatch.setattr(ops, "act", fake_act)
    monkeypatch.setattr(ops, "conv_first", fake_conv_first)
    monkeypatch.setattr(ops, "head", fake_head)
    monkeypatch.setattr(ops, "conv_tc", fake_conv_tc)


def test_split_tf32_is_exact_and_tf32_representable():
    torch.manual_seed(0)
    w = torch.randn(4096) * torch.logspace(-6, 6, 4096)
    hi, lo = engine.split_tf32(w)
    assert torch.equal(hi + lo, w)
    assert int((hi.view(torch.int32) & 0x1fff).abs().max()) == 0
    assert float((lo.abs() / w.abs()).max()) <= 2.0 ** -11


@pytest.mark.parametrize("split,tc,f16", [(False, False, False), (True, False, False), (True, True, False), (True, True, True)])
def test_channels_last_backbone_matches_oracle(cpu_glue, split, tc, f16):
    sd = omodel.randomize_bn_stats(omodel.init_state_dict(56, True, seed=5))
    m = SpatialDROW(num_scans=10, num_pts=56, alpha=0.5, window_size=11, pedestrian_only=True)
    m.load_state_dict(sd, strict=True)
    m.eval()
    with torch.no_grad():
        net = engine._ChannelsLastBackbone(m, split=split, tc=tc, f16=f16)
        b, n = 2, 9
        torch.manual_seed(1)
        cut = torch.randn(b, n, 56).clamp(-1, 1)
        feat, x_op = net.features(cut.view(b * n, 56))                                  # [M*14, 256]
        want = omodel.backbone_front(cut, sd)                                           # [b, n, 256, 14]
        assert_rel(feat.view(b, n, 14, 256).permute(0, 1, 3, 2), want, tol=2e-6, what="features")
        emb = net.embed(x_op, b * n).view(b, n, -1)
        assert_rel(emb, omodel.gate_embed(want, sd), tol=2e-6, what="embedding")
        got_cls, got_reg = torch.empty(b, n, 1), torch.empty(b, n, 2)
        net.votes(net.operand(feat), b * n, 14, got_cls, got_reg)
        cls, reg = omodel.backbone_back(want, sd)
        assert_rel(got_cls, torch.sigmoid(cls), tol=2e-6, what="scores")
        assert_rel(got_reg, reg, tol=2e-6, what="votes")
