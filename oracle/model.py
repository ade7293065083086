"""CPU oracle for the DR-SPAAM network and its spatial-attention memory update.
TEST INFRASTRUCTURE ONLY.

Functional (state-dict driven) torch-CPU restatement of
/root/reference/src/depracted/model/dr_spaam.py:
  * `conv_bn_lrelu`, `backbone_front`, `backbone_back`   <- `_conv` :8-12,
    `DROW._forward_cutout` :87-97, `DROW._forward_fused_cutout` :102-114
  * `gate_dense`                                         <- `_SpatialAttention.forward` :163-217
    with `_generate_neighbor_mask` :145-160 (the reference's dense N x N form:
    this is what gets timed as the CPU baseline)
  * `gate_windowed`                                      <- the same maths restricted to the
    2*hw+1 window (SURVEY.md §8 a-bis); what the CUDA kernel implements
  * `spatial_drow_stream` / `spatial_drow_sequence`      <- `SpatialDROW.forward` :237-250 / :262-277
  * `drow_forward`                                       <- `DROW.forward` :116-121

Only `tests/`, `__graft_entry__.smoke()` and `bench.py`'s CPU-baseline legs may
import it.  Pinned against the imported reference modules (same state dict,
same inputs) by `tests/test_oracle_vs_reference.py` and against
`tests/golden/model_*.npz`.  It takes a plain `state_dict` whose keys are the
reference's (SURVEY.md §8b "Checkpoint compatibility"), so one set of weights
drives the reference, this oracle and the product module.
"""
import math

import torch
import torch.nn.functional as F

LRELU_SLOPE = 0.1  # dr_spaam.py:12


def conv_bn_lrelu(x, sd, prefix, padding, training=False, momentum=0.1, eps=1e-5):
    """Conv1d -> BatchNorm1d -> LeakyReLU(0.1)   (`_conv`, dr_spaam.py:8-12).

    `prefix` addresses the reference's nn.Sequential: `<prefix>.0` is the conv,
    `<prefix>.1` the batch norm.  In training mode the running statistics in
    `sd` are updated in place, as nn.BatchNorm1d does.
    """
    y = F.conv1d(x, sd[prefix + ".0.weight"], sd[prefix + ".0.bias"], padding=padding)
    y = F.batch_norm(y, sd[prefix + ".1.running_mean"], sd[prefix + ".1.running_var"],
                     sd[prefix + ".1.weight"], sd[prefix + ".1.bias"],
                     training=training, momentum=momentum, eps=eps)
    if training and (prefix + ".1.num_batches_tracked") in sd:
        sd[prefix + ".1.num_batches_tracked"] += 1
    return F.leaky_relu(y, LRELU_SLOPE)


def _block(x, sd, name, n_layers, training):
    for k in range(n_layers):
        x = conv_bn_lrelu(x, sd, "%s.%d" % (name, k), padding=1, training=training)
    return x


def backbone_front(cutouts, sd, training=False):
    """[B, N, P] single-scan cutouts -> [B, N, 256, P/4]   (dr_spaam.py:87-97)."""
    b, n, p = cutouts.shape
    y = cutouts.reshape(b * n, 1, p)
    y = F.max_pool1d(_block(y, sd, "conv_block_1", 3, training), 2)
    y = F.max_pool1d(_block(y, sd, "conv_block_2", 3, training), 2)
    return y.reshape(b, n, y.shape[-2], y.shape[-1])


def backbone_back(feat, sd, training=False):
    """[B, N, 256, L] -> (pred_cls [B,N,C], pred_reg [B,N,2])   (dr_spaam.py:102-114)."""
    b, n, c, l = feat.shape
    y = feat.reshape(b * n, c, l)
    y = F.max_pool1d(_block(y, sd, "conv_block_3", 3, training), 2)
    y = _block(y, sd, "conv_block_4", 2, training)
    y = F.avg_pool1d(y, y.shape[-1])
    cls = F.conv1d(y, sd["conv_cls.weight"], sd["conv_cls.bias"]).reshape(b, n, -1)
    reg = F.conv1d(y, sd["conv_reg.weight"], sd["conv_reg.bias"]).reshape(b, n, 2)
    return cls, reg


def gate_embed(feat, sd, training=False):
    """[B,N,256,L] -> [B,N,128] similarity embedding   (dr_spaam.py:130-133,176-181)."""
    b, n, c, l = feat.shape
    e = conv_bn_lrelu(feat.reshape(b * n, c, l), sd, "gate.conv", padding=0, training=training)
    return e.reshape(b, n, -1)


def neighbour_table(n, window_size):
    """Clamped neighbour indices [n, 2*hw+1] and the 0/1 window mask [n, n]  (:145-160)."""
    hw = int(window_size / 2)
    cols = (torch.arange(n)[:, None] + torch.arange(-hw, hw + 1)[None, :]).clamp(0, n - 1)
    mask = torch.zeros(n, n)
    mask.scatter_(1, cols, 1.0)
    return cols, mask


_TABLES = {}


def _cached_table(n, window_size, device):
    """The reference builds its neighbour table once per module (:171-173); same here, per (n, window, device)."""
    key = (n, int(window_size), str(device))
    if key not in _TABLES:
        cols, mask = neighbour_table(n, window_size)
        _TABLES[key] = (cols.to(device), mask.to(device))
    return _TABLES[key]


def gate_dense(x, template, sd, alpha, window_size, training=False):
    """The reference's dense formulation.  Returns (new_template, feat_fused, weights[B,N,N])."""
    b, n, c, l = x.shape
    cols, mask = _cached_table(n, window_size, x.device)
    e_x = gate_embed(x, sd, training)                                # :176-177
    e_t = gate_embed(template, sd, training)                         # :180-181
    sim = torch.matmul(e_x, e_t.transpose(1, 2))                     # :184
    feat_fused = torch.gather(sim, 2, cols.unsqueeze(0).expand(b, -1, -1))   # :187
    sim = sim - 1e10 * (1.0 - mask)                                  # :197
    top = sim.max(dim=-1, keepdim=True)[0]                           # :198
    w = torch.exp(sim - top) * mask                                  # :199
    w = w / w.sum(dim=-1, keepdim=True)                              # :200-201
    mixed = torch.matmul(w, template.reshape(b, n, c * l)).reshape(b, n, c, l)   # :210-212
    return alpha * x + (1.0 - alpha) * mixed, feat_fused, w          # :215


def gate_windowed(x, template, e_x, e_t, alpha, window_size):
    """Windowed statement of the same update, from precomputed embeddings.

    sim[b,i,k] = <e_x[b,i], e_t[b,clamp(i-hw+k)]>; softmax over the UNIQUE
    in-range neighbours j in [max(0,i-hw), min(N-1,i+hw)]; weighted sum of the
    template rows; alpha blend.  Returns (new_template, feat_fused[B,N,W], w[B,N,W])
    where w is zero on clamped duplicates.
    """
    b, n, c, l = x.shape
    hw = int(window_size / 2)
    W = 2 * hw + 1
    raw = torch.arange(n)[:, None] + torch.arange(-hw, hw + 1)[None, :]
    valid = (raw >= 0) & (raw <= n - 1)
    cols = raw.clamp(0, n - 1)
    nb = e_t[:, cols]                                                # [B,N,W,E]
    feat_fused = torch.einsum("bne,bnwe->bnw", e_x, nb)
    s = feat_fused.masked_fill(~valid[None], -float("inf"))
    w = torch.softmax(s, dim=-1)
    flat = template.reshape(b, n, c * l)
    mixed = torch.einsum("bnw,bnwf->bnf", w, flat[:, cols]).reshape(b, n, c, l)
    return alpha * x + (1.0 - alpha) * mixed, feat_fused, w


def spatial_drow_stream(cutouts, sd, alpha, window_size, fea_template=None):
    """`SpatialDROW.forward(x, testing=True, fea_template=...)`   (dr_spaam.py:239-250).

    cutouts: [B, N, S, P]; only scan 0 is used (:240).
    Returns (pred_cls, pred_reg, out_template, feat_fused).
    """
    feat = backbone_front(cutouts[:, :, 0, :], sd)
    if fea_template is None:
        template = feat.clone()                                      # :243
        _, feat_fused, _ = gate_dense(feat, template, sd, alpha, window_size)   # :244
    else:
        template, feat_fused, _ = gate_dense(feat, fea_template, sd, alpha, window_size)  # :246
    cls, reg = backbone_back(template, sd)
    return cls, reg, template, feat_fused


def spatial_drow_sequence(cutouts, sd, alpha, window_size, training=False):
    """`SpatialDROW.forward(x)` training/eval branch   (dr_spaam.py:262-277)."""
    n_scan = cutouts.shape[2]
    template = backbone_front(cutouts[:, :, 0, :], sd, training)     # :264-265
    feat_fused = None
    for s in range(1, n_scan):                                       # :266-273
        feat = backbone_front(cutouts[:, :, s, :], sd, training)
        template, feat_fused, _ = gate_dense(feat, template, sd, alpha, window_size, training)
    cls, reg = backbone_back(template, sd, training)
    return cls, reg, feat_fused


def drow_forward(cutouts, sd, training=False):
    """`DROW.forward`: per-scan features summed over scans   (dr_spaam.py:116-121)."""
    b, n, s, p = cutouts.shape
    feat = backbone_front(cutouts.reshape(b, n * s, p), sd, training)
    feat = feat.reshape(b, n, s, feat.shape[-2], feat.shape[-1]).sum(dim=2)
    return backbone_back(feat, sd, training)


def init_state_dict(num_pts=56, pedestrian_only=True, seed=0, dtype=torch.float32):
    """Random-init weights with the reference's key names, shapes and init rule
    (kaiming-normal a=0.1 on convs, BN weight 1 / bias 0; dr_spaam.py:72-77,138-143)."""
    g = torch.Generator().manual_seed(seed)
    sd = {}

    def conv(prefix, cin, cout, k, with_bn=True):
        gain = math.sqrt(2.0 / (1.0 + 0.1 ** 2))
        std = gain / math.sqrt(cin * k)
        sd[prefix + (".0.weight" if with_bn else ".weight")] = torch.randn(cout, cin, k, generator=g, dtype=dtype) * std
        bound = 1.0 / math.sqrt(cin * k)
        sd[prefix + (".0.bias" if with_bn else ".bias")] = (torch.rand(cout, generator=g, dtype=dtype) * 2 - 1) * bound
        if with_bn:
            sd[prefix + ".1.weight"] = torch.ones(cout, dtype=dtype)
            sd[prefix + ".1.bias"] = torch.zeros(cout, dtype=dtype)
            sd[prefix + ".1.running_mean"] = torch.zeros(cout, dtype=dtype)
            sd[prefix + ".1.running_var"] = torch.ones(cout, dtype=dtype)
            sd[prefix + ".1.num_batches_tracked"] = torch.zeros((), dtype=torch.long)

    chans = {"conv_block_1": (1, 64, 64, 128), "conv_block_2": (128, 128, 128, 256),
             "conv_block_3": (256, 256, 256, 512), "conv_block_4": (512, 256, 128)}
    for name, cs in chans.items():
        for k in range(len(cs) - 1):
            conv("%s.%d" % (name, k), cs[k], cs[k + 1], 3)
    conv("conv_cls", 128, 1 if pedestrian_only else 4, 1, with_bn=False)
    conv("conv_reg", 128, 2, 1, with_bn=False)
    conv("gate.conv", 256, 128, int(math.ceil(num_pts / 4)))
    return sd


def randomize_bn_stats(sd, seed=1):
    """Give the BN layers non-trivial affine parameters and running statistics
    so eval-mode parity tests exercise the whole normalisation formula."""
    g = torch.Generator().manual_seed(seed)
    for k in list(sd):
        if k.endswith(".1.running_mean"):
            sd[k] = torch.randn(sd[k].shape, generator=g) * 0.1
        elif k.endswith(".1.running_var"):
            sd[k] = torch.rand(sd[k].shape, generator=g) * 0.5 + 0.75
        elif k.endswith(".1.weight"):
            sd[k] = torch.rand(sd[k].shape, generator=g) * 0.5 + 0.75
        elif k.endswith(".1.bias"):
            sd[k] = torch.randn(sd[k].shape, generator=g) * 0.05
    return sd
