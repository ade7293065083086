"""CPU oracle for the scan-pair flow prototype.  TEST INFRASTRUCTURE ONLY.

Restates /root/reference/src/depracted/model/prototype.py on the CPU:
  * `fusion_dense`     <- `Prototype._fusion` :118-156, the reference's own dense form (patch gather,
                          [N, N] matmul, +-max_displacement gather) without its stray `.cuda()` scratch tensor
  * `fusion_windowed`  <- the same numbers from the windowed definition the CUDA kernel implements
  * `prototype_forward`<- `Prototype.forward` :58-110, functional over a state dict
Pinned against the imported reference module (tests/test_oracle_vs_reference.py) and
tests/golden/prototype_*.npz.
"""
import torch
import torch.nn.functional as F


def fusion_dense(feat1, feat2, kernel_size=3, max_displacement=5):
    """[B, C, N] x [B, C, N] -> [B, 2D+1, N]   (prototype.py:118-156)."""
    b, c, n = feat1.shape
    half = kernel_size // 2
    patch_ids = (torch.arange(n).unsqueeze(-1) + torch.arange(-half, half + 1).unsqueeze(0)).clamp(0, n - 1)   # :124-127

    def patches(f):                                                                                             # :129-135
        p = f[:, :, patch_ids.reshape(-1)].reshape(b, c, n, kernel_size).permute(0, 1, 3, 2)
        return p.reshape(b, c * kernel_size, n)

    corr = torch.matmul(patches(feat1).permute(0, 2, 1), patches(feat2))                                        # :137
    ids2 = (torch.arange(n).unsqueeze(-1) + torch.arange(-max_displacement, max_displacement + 1).unsqueeze(0)).clamp(0, n - 1)
    ids1 = torch.arange(n).unsqueeze(-1).expand_as(ids2)                                                        # :140-145
    out = corr[:, ids1.reshape(-1), ids2.reshape(-1)].reshape(b, n, -1)                                         # :151
    return out.permute(0, 2, 1)                                                                                 # :152


def fusion_windowed(feat1, feat2, kernel_size=3, max_displacement=5):
    """out[b, d+D, i] = sum_c sum_k f1[b,c,clamp(i+k)] * f2[b,c,clamp(clamp(i+d)+k)]."""
    b, c, n = feat1.shape
    half = kernel_size // 2
    i = torch.arange(n)
    rows = []
    for d in range(-max_displacement, max_displacement + 1):
        j = (i + d).clamp(0, n - 1)
        acc = 0
        for k in range(-half, half + 1):
            acc = acc + (feat1[:, :, (i + k).clamp(0, n - 1)] * feat2[:, :, (j + k).clamp(0, n - 1)]).sum(dim=1)
        rows.append(acc)
    return torch.stack(rows, dim=1)


def _block(x, sd, name, stride, padding, training=False):
    y = F.conv1d(x, sd[name + ".0.weight"], sd[name + ".0.bias"], stride=stride, padding=padding)
    y = F.batch_norm(y, sd[name + ".1.running_mean"], sd[name + ".1.running_var"], sd[name + ".1.weight"], sd[name + ".1.bias"],
                     training=training, momentum=0.1, eps=1e-5)
    return F.leaky_relu(y, 0.01)


def prototype_forward(scan1, scan2, sd, max_displacement=5, training=False, fusion=fusion_dense):
    """[B, N, in_channel] pair -> flow [B, N, 2]   (prototype.py:58-110)."""
    s1, s2 = scan1.permute(0, 2, 1), scan2.permute(0, 2, 1)
    f1_0, f2_0 = _block(s1, sd, "encoder_0", 2, 1, training), _block(s2, sd, "encoder_0", 2, 1, training)
    f1_1, f2_1 = _block(f1_0, sd, "encoder_1", 2, 1, training), _block(f2_0, sd, "encoder_1", 2, 1, training)
    f1_2, f2_2 = _block(f1_1, sd, "encoder_2", 2, 1, training), _block(f2_1, sd, "encoder_2", 2, 1, training)
    feat = fusion(f1_2, f2_2, 3, max_displacement)
    up = lambda x, size: F.interpolate(x, size=size, mode="nearest")     # noqa: E731
    up1 = _block(torch.cat((f1_1, up(feat, f1_1.shape[-1])), dim=1), sd, "decoder_1", 1, 1, training)
    up0 = _block(torch.cat((f1_0, up(up1, f1_0.shape[-1])), dim=1), sd, "decoder_0", 1, 1, training)
    out = _block(torch.cat((s1, up(up0, s1.shape[-1])), dim=1), sd, "flow_reg", 1, 0, training)
    return out.permute(0, 2, 1)


def init_state_dict(in_channel=2, max_displacement=5, seed=0):
    """A reference-shaped state dict with kaiming-normal convolutions and non-trivial BN statistics."""
    g = torch.Generator().manual_seed(seed)
    shapes = {"encoder_0": (64, in_channel, 3), "encoder_1": (128, 64, 3), "encoder_2": (256, 128, 3),
              "decoder_1": (128, 2 * max_displacement + 1 + 128, 3), "decoder_0": (128, 128 + 64, 3),
              "flow_reg": (2, 128 + in_channel, 1)}
    sd = {}
    for name, (co, ci, k) in shapes.items():
        std = (2.0 / (1 + 0.1 ** 2) / (ci * k)) ** 0.5
        sd[name + ".0.weight"] = torch.randn(co, ci, k, generator=g) * std
        sd[name + ".0.bias"] = torch.randn(co, generator=g) * 0.05
        sd[name + ".1.weight"] = 1.0 + 0.1 * torch.randn(co, generator=g)
        sd[name + ".1.bias"] = 0.1 * torch.randn(co, generator=g)
        sd[name + ".1.running_mean"] = 0.1 * torch.randn(co, generator=g)
        sd[name + ".1.running_var"] = 0.75 + 0.5 * torch.rand(co, generator=g)
        sd[name + ".1.num_batches_tracked"] = torch.tensor(0, dtype=torch.long)
    return sd
