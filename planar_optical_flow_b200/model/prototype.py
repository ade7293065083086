"""Scan-pair flow prototype with the reference's API and checkpoint layout.

Mirrors /root/reference/src/depracted/model/prototype.py:
    Prototype   :34-156   (in_channel, max_displacement); forward(scan1, scan2=None) -> [B, N, 2]
    flow_loss   :27-32    -> (loss, err_batch)
`state_dict()` keys equal the reference's (encoder_{0,1,2}, decoder_{1,0}, flow_reg: Sequential of
Conv1d, BatchNorm1d, LeakyReLU(0.01)).  The encoder / decoder convolutions stay on cuDNN; what changed
is `_fusion`: the reference builds K-tap patches, a dense [N, N] correlation matrix per sample and then
gathers +-max_displacement entries (:118-156); here that is one windowed-correlation kernel
(csrc/pof_corr.cu) with its own deterministic backward.  CUDA tensors only.
"""
import torch
import torch.nn as nn
import torch.nn.functional as F

from .. import ops


def _encode_decode(in_channel, out_channel, stride=1):
    return nn.Sequential(nn.Conv1d(in_channel, out_channel, kernel_size=3, stride=stride, padding=1),
                         nn.BatchNorm1d(out_channel),
                         nn.LeakyReLU(negative_slope=0.01, inplace=True))


def _pw_conv(in_channel, out_channel):
    return nn.Sequential(nn.Conv1d(in_channel, out_channel, kernel_size=1),
                         nn.BatchNorm1d(out_channel),
                         nn.LeakyReLU(negative_slope=0.01, inplace=True))


def flow_loss(pred, target, mask=None):
    """Mean end-point error (prototype.py:27-32): returns (loss, err_batch)."""
    err_batch = torch.mean(torch.norm(pred - target, dim=-1), dim=1)
    return torch.mean(err_batch), err_batch


class Prototype(nn.Module):
    def __init__(self, in_channel=1, max_displacement=5):
        super().__init__()
        self.max_displacement = max_displacement
        self.encoder_0 = _encode_decode(in_channel, 64, 2)
        self.encoder_1 = _encode_decode(64, 128, 2)
        self.encoder_2 = _encode_decode(128, 256, 2)
        self.decoder_1 = _encode_decode(2 * self.max_displacement + 1 + 128, 128)
        self.decoder_0 = _encode_decode(128 + 64, 128)
        self.flow_reg = _pw_conv(128 + in_channel, 2)
        self.loss_fn = flow_loss
        for m in self.modules():                                         # :51-56
            if isinstance(m, (nn.Conv1d, nn.Conv2d)):
                nn.init.kaiming_normal_(m.weight, a=0.1, nonlinearity="leaky_relu")
            elif isinstance(m, (nn.BatchNorm1d, nn.BatchNorm2d)):
                nn.init.constant_(m.weight, 1)
                nn.init.constant_(m.bias, 0)

    def forward(self, scan1, scan2=None):
        if scan2 is None:
            scan2 = scan1
        scan1 = scan1.permute(0, 2, 1)                                   # [B, in_channel, N]
        scan2 = scan2.permute(0, 2, 1)
        f1_0, f2_0 = self.encoder_0(scan1), self.encoder_0(scan2)        # :71-81, both scans through the same encoder
        f1_1, f2_1 = self.encoder_1(f1_0), self.encoder_1(f2_0)
        f1_2, f2_2 = self.encoder_2(f1_1), self.encoder_2(f2_1)
        feat = self._fusion(f1_2, f2_2, max_displacement=self.max_displacement)        # :84
        up1 = self.decoder_1(torch.cat((f1_1, self._upsample(feat, f1_1.shape[-1])), dim=1))       # :88-91
        up0 = self.decoder_0(torch.cat((f1_0, self._upsample(up1, f1_0.shape[-1])), dim=1))        # :93-99
        out = self.flow_reg(torch.cat((scan1, self._upsample(up0, scan1.shape[-1])), dim=1))      # :101-105
        return out.permute(0, 2, 1)

    def _upsample(self, x, size):
        return F.interpolate(x, size=size, mode="nearest")

    def _fusion(self, feat1, feat2, kernel_size=3, max_displacement=5):
        return ops.patch_corr(feat1, feat2, kernel_size=kernel_size, max_displacement=max_displacement)
