"""DROW3 / DR-SPAAM host modules with the reference's API and checkpoint layout.

Mirrors /root/reference/src/depracted/model/dr_spaam.py:
    DROW               :41-121    constructor args, forward(x) -> (pred_cls, pred_reg)
    _SpatialAttention  :124-217   (n_pts, n_channel, alpha, window_size); forward(x, x_template)
    SpatialDROW        :220-277   forward(x, testing=False, fea_template=None)
`state_dict()` keys and shapes equal the reference's, so `torch.load(f)["model_state"]`
from a reference checkpoint loads with strict=True.

What changed underneath:
  * the gate no longer builds dense [N, N] similarity / weight matrices: after the two
    embedding convolutions (cuDNN) it calls ONE fused sm_100a kernel
    (csrc/pof_gate.cu) through `ops.gate`, differentiable via its own backward kernels;
  * no neighbour table is cached on the module, so one instance serves any N
    (the reference locks an instance to its first N);
  * the convolutions themselves stay on PyTorch/cuDNN (the only dense contraction; the streaming engine has its own tcgen05
    kernel for inference, engine.py); in the training branch what runs between them - batch-norm with batch statistics,
    LeakyReLU, the block's max-pool - is libpof's one fused operator (`ops.bn_act_pool`, csrc/pof_bnact.cu), the first
    layer (one input channel) has its own forward / weight-gradient kernels, and `scans_per_call` scans of a sample share
    one call per layer with per-scan batch statistics.
The gate only runs on CUDA tensors; there is no CPU path for it.
"""
import os
from math import ceil

import torch
import torch.nn as nn
import torch.nn.functional as F

from .. import ops
from .loss_utils import BinaryFocalLoss, FocalLoss, flow_loss

_SLOPE = 0.1


class _ConvBnAct(nn.Sequential):
    """Conv1d + BatchNorm1d + LeakyReLU with the reference's child names ("0", "1", "2": dr_spaam.py:8-12).

    A 3-D input [M, C, L] runs the children as the reference does.  A 4-D channels-last input [M, C, 1, L] (the
    training branch of SpatialDROW) runs the SAME parameters through conv2d / batch_norm: cuDNN then stays in NHWC for
    the TF32 implicit GEMMs and the batch-norm statistics, instead of transposing around every convolution and using
    its per-channel NCHW batch-norm kernels, which took 45 % + 16 % of the training step (profiles/r1_train_step.txt).
    """

    fused = True        # batch-norm + LeakyReLU (+ pool) of the training branch as libpof's one operator (False: cuDNN + PyTorch kernels)

    def forward(self, x, pool=1, groups=1):
        """`pool` = 2 applies the block's max_pool1d(2) (dr_spaam.py:81) here, so it can be fused with the activation.
        `groups` > 1 (4-D training path only): x holds `groups` equal consecutive blocks of cutouts - the scans of the training
        samples - that the reference pushes through this layer in `groups` separate calls (dr_spaam.py:264-273); the
        convolution runs once over all of them and the batch norm keeps per-block statistics (`can_group`)."""
        if x.dim() != 4:
            y = super().forward(x)
            return F.max_pool1d(y, kernel_size=pool) if pool > 1 else y
        return self._forward_4d(x, pool, groups)

    def can_group(self, x):
        """True if this layer can take several scans at once with per-scan batch statistics (the fused CUDA operator)."""
        conv, bn = self[0], self[1]
        c = conv.out_channels
        return bool(self.fused and x.is_cuda and bn.training and bn.track_running_stats and bn.affine and bn.momentum is not None
                    and conv.bias is not None and c % 4 == 0 and 256 % (c // 4) == 0)

    def _forward_4d(self, x, pool, groups=1):
        conv, bn = self[0], self[1]
        w4, pad = conv.weight.unsqueeze(2), (0, conv.padding[0])
        track = bn.track_running_stats

        def pooled(t):
            return F.max_pool2d(t, kernel_size=(1, pool)) if pool > 1 else t

        if not (bn.training or not track):                      # eval: running statistics
            y = F.batch_norm(F.conv2d(x, w4, conv.bias, padding=pad), bn.running_mean, bn.running_var, bn.weight, bn.bias,
                             False, 0.0, bn.eps)
            return pooled(F.leaky_relu_(y, _SLOPE))
        # Batch statistics: the convolution's bias cancels in (y + b) - mean(y + b), so it is left out of the forward
        # pass (no bias-add pass, and no reduction of the output gradient for a bias gradient that is zero in exact
        # arithmetic - 27 % of the step).  It still belongs to the running mean, and it still gets its (zero) gradient.
        if bn.training and track:
            bn.num_batches_tracked.add_(1)
        if groups > 1:
            if not self.can_group(x):
                raise RuntimeError("this layer cannot take several scans at once; call it per scan")
            if bn.training and track:
                bn.num_batches_tracked.add_(groups - 1)           # one more call per extra scan
        factor = bn.momentum if bn.momentum is not None else 1.0 / max(float(bn.num_batches_tracked), 1.0)
        skip_bias = conv.bias is not None and factor < 1.0
        if skip_bias and bn.training and track:
            # rm <- (1 - f) rm + f (mean(y) + b): fold f b in before batch_norm's own update.  Through .data: earlier calls of
            # the same layer in this step (the gate embeds every scan) saved the buffer, and a version bump would make
            # autograd refuse their backward although batch-statistics backward never reads it.
            # For G grouped calls the bias enters every one of the G updates: rm_G = (1-f)^G rm_0 + sum_g f (1-f)^(G-1-g) (m_g + b),
            # and the b terms sum to (1 - (1-f)^G) b, which is what adding b (1 - (1-f)^G) / (1-f)^G up front gives.
            keep = (1.0 - factor) ** groups
            bn.running_mean.data.add_(conv.bias.detach(), alpha=(1.0 - keep) / keep)
        if (self.fused and skip_bias and x.is_cuda and conv.in_channels == 1 and conv.kernel_size == (3,) and pad == (0, 1)
                and not x.requires_grad and x.dtype == torch.float32 and conv.out_channels % 4 == 0 and 256 % (conv.out_channels // 4) == 0):
            y = ops.conv_first_train(x.reshape(x.shape[0], x.shape[3]), conv.weight)      # cuDNN has no fast engine for one input channel
        else:
            y = F.conv2d(x, w4, None if skip_bias else conv.bias, padding=pad)
        c = y.shape[1]
        if self.fused and y.is_cuda and bn.affine and c % 4 == 0 and 256 % (c // 4) == 0 and (pool == 1 or y.shape[3] % 2 == 0):
            y = ops.bn_act_pool(y, bn.weight, bn.bias, bn.running_mean if track else None, bn.running_var if track else None,
                                momentum=factor, eps=bn.eps, slope=_SLOPE, pool=pool, groups=groups)
        else:
            y = F.batch_norm(y, bn.running_mean if track else None, bn.running_var if track else None, bn.weight, bn.bias,
                             True, factor, bn.eps)
            y = pooled(F.leaky_relu_(y, _SLOPE))
        return _ZeroGradOperand.apply(y, conv.bias) if skip_bias else y


class _ZeroGradOperand(torch.autograd.Function):
    """y unchanged; `p` joins the graph with an exactly-zero gradient (DDP wants every parameter to receive one)."""

    @staticmethod
    def forward(ctx, y, p):
        ctx.p_shape, ctx.p_dtype, ctx.p_device = p.shape, p.dtype, p.device
        return y.view_as(y)

    @staticmethod
    def backward(ctx, g):
        return g, torch.zeros(ctx.p_shape, dtype=ctx.p_dtype, device=ctx.p_device)


def _conv(in_channel, out_channel, kernel_size, padding):
    return _ConvBnAct(
        nn.Conv1d(in_channel, out_channel, kernel_size=kernel_size, padding=padding),
        nn.BatchNorm1d(out_channel),
        nn.LeakyReLU(negative_slope=_SLOPE, inplace=True),
    )


def _stack(*channels):
    """Chain of k=3, p=1 conv-bn-lrelu layers through the given channel counts."""
    return nn.Sequential(*[_conv(cin, cout, 3, 1) for cin, cout in zip(channels[:-1], channels[1:])])


def _init_like_reference(module):
    for m in module.modules():
        if isinstance(m, (nn.Conv1d, nn.Conv2d)):
            nn.init.kaiming_normal_(m.weight, a=_SLOPE, nonlinearity="leaky_relu")
        elif isinstance(m, (nn.BatchNorm1d, nn.BatchNorm2d)):
            nn.init.constant_(m.weight, 1)
            nn.init.constant_(m.bias, 0)


class DROW(nn.Module):
    def __init__(self, dropout=0.5, num_scans=5, num_pts=48, focal_loss_gamma=0.0, pedestrian_only=False):
        super().__init__()
        self.dropout = 0.0                       # the reference forces dropout off (:46-47)
        self.conv_block_1 = _stack(1, 64, 64, 128)
        self.conv_block_2 = _stack(128, 128, 128, 256)
        self.conv_block_3 = _stack(256, 256, 256, 512)
        self.conv_block_4 = _stack(512, 256, 128)
        if pedestrian_only:
            self.conv_cls = nn.Conv1d(128, 1, kernel_size=1)
            self.cls_loss = BinaryFocalLoss(gamma=focal_loss_gamma) if focal_loss_gamma > 0.0 \
                else F.binary_cross_entropy
        else:
            self.conv_cls = nn.Conv1d(128, 4, kernel_size=1)
            self.cls_loss = FocalLoss(gamma=focal_loss_gamma) if focal_loss_gamma > 0.0 else F.cross_entropy
        self.conv_reg = nn.Conv1d(128, 2, kernel_size=1)
        _init_like_reference(self)

    # -- per-cutout feature extractor: [B, N, S, P] -> [B, N, S, 256, P/4]   (:87-97)
    def _forward_cutout(self, x):
        b, n, s, p = x.shape
        y = x.reshape(b * n * s, 1, p)
        y = F.max_pool1d(self.conv_block_1(y), kernel_size=2)
        y = F.max_pool1d(self.conv_block_2(y), kernel_size=2)
        return y.view(b, n, s, y.shape[-2], y.shape[-1])

    def _fuse_cutout(self, x):
        return torch.sum(x, dim=2)

    # -- the same two halves on channels-last 4-D activations [M, C, 1, L] (see _ConvBnAct) ------------------
    def _features_cl(self, scan_cutouts, groups=1):
        """[B, N, P] cutouts of one scan -> [B*N, 256, 1, P/4], channels-last memory.  With `groups` = S the input is
        [S * B, N, P] (scan-major): all scans of the samples in one pass, batch statistics per scan."""
        b, n, p = scan_cutouts.shape
        y = scan_cutouts.reshape(b * n, 1, 1, p).contiguous(memory_format=torch.channels_last)
        for blk in (self.conv_block_1, self.conv_block_2):
            for k, layer in enumerate(blk):
                y = layer(y, pool=2 if k == len(blk) - 1 else 1, groups=groups)   # the block's max_pool1d(2), fused into its last layer
        return y

    def _can_group_features(self, x):
        return all(layer.can_group(x) for blk in (self.conv_block_1, self.conv_block_2) for layer in blk)

    def _votes_cl(self, t, b, n):
        """[B*N, 256, 1, L] channels-last fused features -> ([B, N, C], [B, N, 2])."""
        y = t
        for k, layer in enumerate(self.conv_block_3):
            y = layer(y, pool=2 if k == len(self.conv_block_3) - 1 else 1)
        for layer in self.conv_block_4:
            y = layer(y)
        y = y.mean(dim=3)                                      # avg_pool1d over the whole length: [M, 128, 1]
        return self.conv_cls(y).view(b, n, -1), self.conv_reg(y).view(b, n, 2)

    # -- fused feature -> votes: [B, N, 256, L] -> ([B, N, C], [B, N, 2])     (:102-114)
    def _forward_fused_cutout(self, x):
        b, n, c, l = x.shape
        y = x.reshape(b * n, c, l)
        y = F.max_pool1d(self.conv_block_3(y), kernel_size=2)
        y = self.conv_block_4(y)
        y = F.avg_pool1d(y, kernel_size=y.shape[-1])
        return self.conv_cls(y).view(b, n, -1), self.conv_reg(y).view(b, n, 2)

    def forward(self, x):
        return self._forward_fused_cutout(self._fuse_cutout(self._forward_cutout(x)))


class _SpatialAttention(nn.Module):
    def __init__(self, n_pts, n_channel, alpha=0.5, window_size=7):
        super().__init__()
        self._alpha = alpha
        self._window_size = window_size
        self.conv = _conv(n_channel, 128, kernel_size=n_pts, padding=0)
        _init_like_reference(self)

    @property
    def window(self):
        """Effective window 2*hw+1 (the reference uses hw = int(window_size / 2), :148)."""
        return 2 * int(self._window_size / 2) + 1

    def embed(self, feat):
        b, n, c, l = feat.shape
        return self.conv(feat.reshape(b * n, c, l)).view(b, n, -1)

    def forward(self, x, x_template):
        """(x, x_template) [B, N, C, L] -> (out_temp [B, N, C, L], feat_fused [B, N, W])."""
        emb_x = self.embed(x)                    # :176-177
        emb_t = self.embed(x_template)           # :180-181
        return ops.gate(x, x_template, emb_x, emb_t, self._alpha, self.window)   # :183-215 fused


class SpatialDROW(DROW):
    # Training branch: how many scans of a sample go through conv blocks 1-2 in one call (per-scan batch statistics are kept
    # either way, ops.bn_act_pool(groups=...)).  Measured on B200 (bench.py --workload train, 8 x 11 scans x 450 points): 1 scan per
    # call 30.5 ms per step, 2: 26.3, 4: 22.1, all 11: 32.3 - fewer, larger launches and no per-scan gradient accumulation up to
    # a point; beyond it cuDNN answers the batch with a much slower (strided) dgrad engine and the activations of a call no
    # longer stay in L2 between the passes of the batch-norm operator.
    scans_per_call = int(os.environ.get("POF_TRAIN_SCANS_PER_CALL", "4"))

    def __init__(self, dropout=0.5, num_scans=5, num_pts=48, focal_loss_gamma=0.0, alpha=0.5, window_size=7,
                 pedestrian_only=False):
        super().__init__(dropout=dropout, num_scans=num_scans, num_pts=num_pts,
                         focal_loss_gamma=focal_loss_gamma, pedestrian_only=pedestrian_only)
        self.gate = _SpatialAttention(n_pts=int(ceil(num_pts / 4)), n_channel=256, alpha=alpha,
                                      window_size=window_size)
        self.loss_fn = flow_loss

    def _scan_features(self, x, s):
        return self._forward_cutout(x[:, :, s, :].unsqueeze(2)).squeeze(2)

    def forward(self, x, testing=False, fea_template=None):
        if testing:                                                   # streaming branch (:239-250)
            out = self._scan_features(x, 0)
            if fea_template is None:
                out_template = out.clone()                            # first frame: memory := features
                _, feat_fused = self.gate(out, out_template)
            else:
                out_template, feat_fused = self.gate(out, fea_template)
            pred_cls, pred_reg = self._forward_fused_cutout(out_template)
            return pred_cls, pred_reg, out_template, feat_fused

        # training / eval branch (:262-277), on channels-last activations.  The gate kernel works on whole rows of
        # C*L features and only needs x and the memory to agree on their order: here both are [L, C]
        b, n, n_scan = x.shape[0], x.shape[1], x.shape[2]

        def rows(t4):                                                 # [M, C, 1, L] channels-last -> [B, N, L, C] view
            return t4.permute(0, 2, 3, 1).reshape(b, n, t4.shape[3], t4.shape[1])

        def embed(t4):
            return self.gate.conv(t4).view(b, n, -1)

        # The per-scan features do not depend on the recurrence: with the fused operator all S scans go through conv blocks
        # 1-2 in ONE pass (one convolution and one weight gradient per layer instead of S, no S-fold gradient accumulation),
        # the batch norm keeping the per-scan statistics and running-statistics updates of the reference's S separate calls.
        feats = None
        per_call = max(1, min(int(self.scans_per_call), n_scan))
        if per_call > 1 and self.training and self._can_group_features(x):
            feats = []
            for s0 in range(0, n_scan, per_call):                 # consecutive scans, in scan order: the running statistics see the
                g = min(per_call, n_scan - s0)                    # same sequence of updates as the reference's per-scan calls
                block = x[:, :, s0:s0 + g, :].permute(2, 0, 1, 3).reshape(g * b, n, x.shape[3])      # scan-major
                f = self._features_cl(block, groups=g)                                             # [g*B*N, 256, 1, L]
                feats.extend(f.view(g, b * n, f.shape[1], 1, f.shape[3]).unbind(0))

        def scan_features(s):
            if feats is not None:
                return feats[s]
            return self._features_cl(x[:, :, s, :])

        tmpl = scan_features(0)
        feat_fused = None
        for s in range(1, max(n_scan, 2)):          # a single-scan input gates scan 0 with itself (:271-273)
            cur = scan_features(min(s, n_scan - 1))
            out_rows, feat_fused = ops.gate(rows(cur), rows(tmpl), embed(cur), embed(tmpl), self.gate._alpha, self.gate.window)
            tmpl = out_rows.view(b * n, 1, out_rows.shape[2], out_rows.shape[3]).permute(0, 3, 1, 2)
        pred_cls, pred_reg = self._votes_cl(tmpl, b, n)
        return pred_cls, pred_reg, feat_fused
