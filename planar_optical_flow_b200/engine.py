"""Batched streaming DR-SPAAM inference with the whole per-scan chain on the device.

The reference's streaming loop (depracted_scripts/infer_person_flow.py:101-139) handles ONE
sequence, one scan at a time, and crosses the host boundary twice per scan: CPU cutout ->
H2D -> network -> D2H -> CPU NMS.  `StreamingDetector` runs B independent sequences in
lock-step with nothing but the raw ranges going up and the detections coming down:

    ranges [B, N] --H2D--> cutout kernel -> conv blocks 1-2 -> embeddings
        -> fused attention-memory kernel (memory [B, N, 256, 14] stays resident, ping-pong)
        -> conv blocks 3-4 + heads -> sigmoid -> NMS kernels --D2H--> detections

(the convolutions run on libpof's tcgen05 kernel in the default precision, on cuDNN in the library modes below)

Sequences never interact, so multi-GPU use is one detector per rank over a
disjoint set of sequences with no data-path collective.

Numerics (BatchNorm is folded into the convolutions in eval mode in every mode, which
re-associates one multiply):
  * `precision="fp32"`   (default) fp32-ACCURATE convolutions on the tcgen05 tensor cores
                         (csrc/pof_conv_tc.cu): operands split x = hi + lo into two 11-bit-significand parts,
                         the three significant partial products lo*w_hi + hi*w_lo + hi*w_hi accumulated in
                         short chains that are promoted to fp32 registers with rounded adds.  The parts are
                         float16 (kind::f16: twice the TF32 rate, half the operand bytes); each layer's weights
                         are pre-scaled by a power of two so both parts are normal numbers.  Channels-last
                         activations; bias + LeakyReLU + max-pool + split are the kernel's epilogue.
  * `precision="fp32-tf32"`  the same kernel with TF32 parts in fp32 containers (kind::tf32): no range limit
                         on the activations (float16 parts need |x| <= 65504, checked on the device), half
                         the speed.  Per layer 5-7e-7 of the fp64 result; cuDNN's fp32 SIMT kernels: 1-2e-6.
  * `precision="fp32-simt"`  every convolution in IEEE fp32 on cuDNN's SIMT kernels, NCL layout as in the
                         reference.  The library baseline of the mode above (5x slower).
  * `precision="tf32x3"` the same split run through cuDNN's TF32 convolutions ([hi | lo | hi] x
                         [w_hi | w_hi | w_lo] channels): exact products, but the tensor core accumulates
                         with truncation over the whole 4608-deep reduction -> 1e-4, not parity.
  * `precision="tf32"`   plain TF32 on cuDNN, channels-last (what PyTorch's defaults give the reference on
                         a GPU, ~1e-3).  Throughput mode, reported separately.
"""
import contextlib

import numpy as np
import torch
import torch.nn.functional as F

from . import ops

_SLOPE = 0.1


def fold_conv_bn(seq):
    """(Conv1d, BatchNorm1d, LeakyReLU) in eval mode -> (weight, bias) of the equivalent conv."""
    conv, bn = seq[0], seq[1]
    scale = bn.weight / torch.sqrt(bn.running_var + bn.eps)
    w = conv.weight * scale[:, None, None]
    b = (conv.bias - bn.running_mean) * scale + bn.bias
    return w.detach().contiguous(), b.detach().contiguous()


class _FoldedStack:
    """A conv block with BN folded: conv1d(+bias) -> leaky_relu_, per layer."""

    def __init__(self, block, padding):
        self.layers = [fold_conv_bn(layer) for layer in block]
        self.padding = padding

    def __call__(self, x):
        for w, b in self.layers:
            x = F.leaky_relu_(F.conv1d(x, w, b, padding=self.padding), _SLOPE)
        return x


def split_tf32(w):
    """w = hi + lo, hi = w rounded to TF32 (nearest even on the 13 dropped bits); both exact in fp32."""
    u = w.contiguous().view(torch.int32)
    hi = ((u + 0x0fff + ((u >> 13) & 1)) & -8192).view(torch.float32)
    return hi, w - hi


class _ChannelsLastBackbone:
    """The DROW/SpatialDROW convolutions over channels-last activations [rows = cutout x position, C].

    split=True: operands carry the [hi | lo | hi] TF32 split (3C channels), weights [w_hi | w_hi | w_lo].
    The convolutions themselves are cuDNN (conv2d NHWC with H = 1), as north_star prescribes for the
    only dense contraction; everything between them is libpof's fused glue.
    """

    def __init__(self, model, split, tc=False, f16=False):
        self.split = bool(split)
        self.tc = bool(tc)                  # convolutions on libpof's tcgen05 kernel instead of cuDNN
        self.f16 = bool(f16) and self.tc    # ... with float16 hi / lo parts (kind::f16) instead of TF32 parts
        self.parts = ops.SPLIT_F16 if self.f16 else (2 if tc else 3)
        self.timer = None                   # optional: callable(name, flops) -> context manager around each tcgen05 launch
        self.status = None                  # the owner's device status word (ops.new_status), passed to every libpof launch
        blocks = [model.conv_block_1, model.conv_block_2, model.conv_block_3, model.conv_block_4]
        folded = [[fold_conv_bn(layer) for layer in blk] for blk in blocks]
        w0, b0 = folded[0][0]
        self.w_first, self.b_first = w0.reshape(w0.shape[0], 3).contiguous(), b0
        self.layers = [[(self._conv_weight(w), b, w.shape[0]) for w, b in blk] for blk in folded]
        we, be = fold_conv_bn(model.gate.conv)                      # [E, C, L]
        self.emb_b = be
        self.emb_w = self._embed_weight(we)                         # [L * Ceff, E]
        self.w_head = torch.cat([model.conv_cls.weight.detach()[:, :, 0], model.conv_reg.weight.detach()[:, :, 0]]).contiguous()
        self.b_head = torch.cat([model.conv_cls.bias.detach(), model.conv_reg.bias.detach()]).contiguous()
        self.n_cls = int(model.conv_cls.weight.shape[0])

    def _cat(self, w, dim):
        if not self.split:
            return w
        hi, lo = split_tf32(w)
        return torch.cat([hi, hi, lo], dim=dim)

    def _tc_weight(self, w):                                         # [Cout, Cin, taps] -> ([taps, 2, Cout, Cin], out_scale)
        if self.f16:
            # a power of two brings the largest weight to [2^12, 2^13): hi and lo are then normal float16 numbers
            # (lo ~ 2^-11 hi), and the epilogue undoes it exactly
            top = float(w.abs().max())
            s = 13 - int(np.ceil(np.log2(top))) if top > 0 else 0
            ws = w * (2.0 ** s)
            hi = ws.to(torch.float16)
            lo = (ws - hi.float()).to(torch.float16)
            return torch.stack([hi, lo], dim=0).permute(3, 0, 1, 2).contiguous(), 2.0 ** -s
        hi, lo = split_tf32(w)
        lo, _ = split_tf32(lo)
        return torch.stack([hi, lo], dim=0).permute(3, 0, 1, 2).contiguous(), 1.0

    def _conv_weight(self, w):                                       # [Cout, Cin, 3] -> [Cout, Ceff, 1, 3] NHWC
        if self.tc:
            return self._tc_weight(w)
        return self._cat(w, 1).unsqueeze(2).contiguous(memory_format=torch.channels_last)

    def _embed_weight(self, w):                                      # [E, C, L] -> [L * Ceff, E]
        if self.tc:
            return self._tc_weight(w)
        return self._cat(w, 1).permute(2, 1, 0).reshape(-1, w.shape[0]).contiguous()

    def _conv(self, a, M, L, w4, cout):
        x4 = a.view(M, 1, L, a.shape[1]).permute(0, 3, 1, 2)         # [M, Ceff, 1, L], NHWC strides, no copy
        y = F.conv2d(x4, w4, None, padding=(0, 1)).permute(0, 2, 3, 1)
        if not y.is_contiguous():
            raise RuntimeError("cuDNN returned a non-channels-last convolution output")
        return y.view(M * L, cout)

    def _conv_tc(self, a, w, b, M, LA, Lout, taps, pad, **kw):
        """ops.conv_tc, optionally bracketed by the owner's event timer (algorithmic FLOPs: 2 * rows * Cin * Cout * taps)."""
        w, out_scale = w
        if self.timer is None:
            return ops.conv_tc(a, w, b, M, LA, Lout, taps, pad, out_scale=out_scale, status=self.status, **kw)
        cin, cout = a.shape[1] // 2, w.shape[2]
        with self.timer("conv%d" % min(cout, 256), 2.0 * M * Lout * cin * cout * taps):
            return ops.conv_tc(a, w, b, M, LA, Lout, taps, pad, out_scale=out_scale, status=self.status, **kw)

    def _layer(self, a, M, L, w, b, cout, pool, plain=False, plain_out=None):
        """One conv + bias + LeakyReLU (+ max-pool) layer -> (plain or None, operand for the next convolution)."""
        if self.tc:
            return self._conv_tc(a, w, b, M, L, L, 3, 1, pool=pool, slope=_SLOPE, want_plain=plain, want_split=True,
                                 plain_out=plain_out if plain else None)
        y = self._conv(a, M, L, w, cout)
        p, s = ops.act(y, b, pool=pool, slope=_SLOPE, want_plain=plain or not self.split, want_split=self.split,
                       status=self.status)
        if plain and plain_out is not None:
            plain_out.view(p.shape).copy_(p)
            p = plain_out.view(p.shape)
        return p, (s if self.split else p)

    @property
    def gate_emits_operand(self):
        """float16 parts: the attention kernel writes the new memory's operand split itself (no extra pass)."""
        return self.f16

    def operand(self, plain_rows):
        """The convolution operand of already-activated rows (the attention memory)."""
        if not self.split:
            return plain_rows
        return ops.act(plain_rows, None, pool=1, slope=1.0, want_plain=False, want_split=True, parts=self.parts,
                       status=self.status)[1]

    def features(self, cutouts, plain_out=None):
        """cutouts [M, P] -> (features [M * P/4, 256] plain, the same rows as a convolution operand).
        `plain_out`: write the plain features there (the first frame's memory) instead of a fresh tensor."""
        M, L = cutouts.shape
        p, s = ops.conv_first(cutouts, self.w_first, self.b_first, slope=_SLOPE, want_plain=not self.split,
                              want_split=self.split, parts=self.parts, status=self.status)
        a = s if self.split else p
        plain = None
        for bi in (0, 1):
            todo = self.layers[bi][1:] if bi == 0 else self.layers[bi]
            for k, (w4, b, cout) in enumerate(todo):
                last = k == len(todo) - 1
                plain, a = self._layer(a, M, L, w4, b, cout, 2 if last else 1, plain=last and bi == 1,
                                       plain_out=plain_out)
                if last:
                    L //= 2
        return plain, a

    def embed(self, operand, M, out=None):
        """Gate embedding (Conv1d k = L, no padding == one GEMM over whole rows) + BN + LeakyReLU -> [M, E]."""
        if self.tc:
            L = self.emb_w[0].shape[0]
            return self._conv_tc(operand, self.emb_w, self.emb_b, M, L, 1, L, 0, pool=1, slope=_SLOPE, want_plain=True,
                                 want_split=False, plain_out=out)[0]
        e = F.leaky_relu_(torch.addmm(self.emb_b, operand.view(M, -1), self.emb_w), _SLOPE)
        if out is not None:
            out.view(e.shape).copy_(e)
            return out
        return e

    def votes(self, operand, M, L, out_cls, out_reg):
        """memory rows (as operand) [M * L, Ceff] -> sigmoid(cls) into out_cls [M, n_cls], reg into out_reg [M, 2]."""
        a = operand
        for bi in (2, 3):
            for k, (w4, b, cout) in enumerate(self.layers[bi]):
                last = k == len(self.layers[bi]) - 1
                if bi == 3 and last:
                    if self.tc:         # bias + LeakyReLU already applied by the convolution's epilogue
                        y = self._conv_tc(a, w4, b, M, L, L, 3, 1, pool=1, slope=_SLOPE, want_plain=True, want_split=False)[0]
                        return ops.head(y, None, M, L, self.w_head, self.b_head, n_sigmoid=self.n_cls, slope=1.0,
                                        out=out_cls, out_rest=out_reg)
                    y = self._conv(a, M, L, w4, cout)
                    return ops.head(y, b, M, L, self.w_head, self.b_head, n_sigmoid=self.n_cls, slope=_SLOPE,
                                    out=out_cls, out_rest=out_reg)
                _, a = self._layer(a, M, L, w4, b, cout, 2 if last else 1)
                if last:
                    L //= 2


class StreamingDetector:
    """B sequences in lock-step; owns their attention memory.

    model          planar_optical_flow_b200.model.SpatialDROW (weights are read at construction)
    scan_phi       [N] beam angles (NumPy float32/float64; dtype selects the cutout arithmetic)
    cutout_kwargs  the reference's `cutout_kwargs` (config/dr_spaam.yaml:21-28)
    """

    def __init__(self, model, scan_phi, cutout_kwargs, num_sequences, device=None, precision="fp32",
                 min_dist=0.5, seq_chunk=None, record_events=False, cutout_fast=False, cuda_graph=False):
        if not torch.cuda.is_available():
            raise RuntimeError("StreamingDetector needs a CUDA device; there is no CPU path")
        if precision not in ("fp32", "fp32-tf32", "fp32-simt", "tf32x3", "tf32"):
            raise ValueError("precision must be 'fp32', 'fp32-tf32', 'fp32-simt', 'tf32x3' or 'tf32'")
        self.device = torch.device(device) if device is not None else torch.device("cuda", torch.cuda.current_device())
        self.precision = precision
        self.cutout_kwargs = dict(cutout_kwargs)
        # The cutout is 0.1 % of a step, so the engine takes the reference's exact arithmetic (bit-equal samples given the
        # half-angles); cutout_fast=True selects the fixed-point kernel that the cutout-only sweep is about (1e-5 bar).
        self.cutout_fast = bool(cutout_fast)
        self.min_dist = float(min_dist)
        self.B = int(num_sequences)
        phi = np.ascontiguousarray(scan_phi)
        if phi.dtype not in (np.float32, np.float64):
            phi = phi.astype(np.float64)
        self.N = int(phi.shape[0])
        self.phi = torch.from_numpy(phi).to(self.device)
        self.P = int(self.cutout_kwargs.get("num_cutout_pts", 48))

        model = model.to(self.device).eval()
        self.alpha = float(model.gate._alpha)
        self.window = int(model.gate.window)
        self.channels_last = precision != "fp32-simt"
        with torch.no_grad():
            if self.channels_last:
                # two max-pools, then a third inside block 3: every pooled length must be even (the reference's
                # max_pool1d floors instead; precision="fp32-simt" keeps that behaviour for other sizes)
                if self.P % 8:
                    raise ValueError("precision=%r needs num_cutout_pts to be a multiple of 8 (got %d); "
                                     "precision='fp32-simt' takes any size" % (precision, self.P))
                self.net = _ChannelsLastBackbone(model, split=precision != "tf32", tc=precision in ("fp32", "fp32-tf32"),
                                                 f16=precision == "fp32")
            self.block1 = _FoldedStack(model.conv_block_1, 1)
            self.block2 = _FoldedStack(model.conv_block_2, 1)
            self.block3 = _FoldedStack(model.conv_block_3, 1)
            self.block4 = _FoldedStack(model.conv_block_4, 1)
            self.embed = _FoldedStack([model.gate.conv], 0)
            self.w_cls, self.b_cls = model.conv_cls.weight.detach().clone(), model.conv_cls.bias.detach().clone()
            self.w_reg, self.b_reg = model.conv_reg.weight.detach().clone(), model.conv_reg.bias.detach().clone()
        if self.w_cls.shape[0] != 1:
            raise ValueError("the NMS stage needs a one-class head (pedestrian_only=True), as utils.py:536 asserts")

        # sequences are processed in chunks so that the widest activation (128 ch x 56 x 4 B per
        # point) stays a few GB regardless of B
        if seq_chunk:
            self.seq_chunk = int(seq_chunk)
        else:
            # the split operands are 3x wider: keep every activation below 2^30 elements
            n_chunks = max(1, -(-self.B * self.N // (73728 if self.channels_last else 1 << 18)))
            self.seq_chunk = -(-self.B // n_chunks)
        C, L = 256, int(np.ceil(self.P / 4))
        # memory rows are [C, L] (reference layout) or, in the channels-last modes, [L, C]; the gate
        # kernel only needs x and the memory to agree
        shape = (self.B, self.N, L, C) if self.channels_last else (self.B, self.N, C, L)
        self.memory = [torch.empty(shape, dtype=torch.float32, device=self.device) for _ in range(2)]
        self.emb_memory = torch.empty((self.B, self.N, 128), dtype=torch.float32, device=self.device)
        self.cur = 0                       # index of the buffer holding the current memory
        self.has_memory = False
        self.steps_done = 0

        # pinned staging for the host-facing call
        self.h_scans = torch.empty((self.B, self.N), dtype=torch.float32).pin_memory()
        self.d_scans = torch.empty((self.B, self.N), dtype=torch.float32, device=self.device)
        self.h_out = {
            "n_keep": torch.empty((self.B,), dtype=torch.int32).pin_memory(),
            "keep_idx": torch.empty((self.B, self.N), dtype=torch.int32).pin_memory(),
            "instance_mask": torch.empty((self.B, self.N), dtype=torch.int32).pin_memory(),
            "det_xy": torch.empty((self.B, self.N, 2), dtype=torch.float64).pin_memory(),
            "det_cls": torch.empty((self.B, self.N), dtype=torch.float32).pin_memory(),
        }
        self.h2d_bytes_per_step = self.h_scans.numel() * 4
        self.d2h_bytes_per_step = sum(t.numel() * t.element_size() for t in self.h_out.values())
        # one status word per detector: float16-range overflow of the operand split, or a bounded device wait that timed out
        self.status = ops.new_status(self.device)
        self.h_status = torch.zeros(1, dtype=torch.int32).pin_memory()
        if self.channels_last:
            self.net.status = self.status
        # cuda_graph=True: from the third step on, the whole chain of a step (~70 launches per chunk of sequences) is
        # replayed as ONE CUDA graph per memory parity instead of being issued launch by launch from Python - what a
        # small batch (a robot's single scanner: B = 1) is bound by
        self.cuda_graph = bool(cuda_graph)
        if self.cuda_graph and record_events:
            raise ValueError("record_events places CUDA events between the launches; it cannot be combined with cuda_graph")
        self._graphs = {}
        self._graph_pool = None
        self.g_scans = torch.empty((self.B, self.N), dtype=torch.float32, device=self.device) if self.cuda_graph else None
        self.record_events = record_events
        self.events = {"cutout": [], "gate": [], "nms": []}
        self.event_work = {}               # stage -> algorithmic FLOPs of each recorded launch (tcgen05 convolutions)
        if self.channels_last and record_events:
            self.net.timer = self._timed
        self.kernel_launches = 0           # launches of libpof kernels (not cuDNN / torch)

    # ------------------------------------------------------------------ helpers
    def reset(self):
        self.has_memory = False
        self.steps_done = 0
        self.cur = 0

    @contextlib.contextmanager
    def _precision(self):
        old = (torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32)
        allow = self.precision in ("tf32x3", "tf32")
        torch.backends.cudnn.allow_tf32 = allow
        torch.backends.cuda.matmul.allow_tf32 = allow
        try:
            yield
        finally:
            torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32 = old

    @contextlib.contextmanager
    def _timed(self, name, work=None):
        if not self.record_events:
            yield
            return
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record()
        yield
        e.record()
        self.events.setdefault(name, []).append((s, e))
        if work is not None:
            self.event_work.setdefault(name, []).append(work)

    def event_ms(self, name):
        """Durations (ms) of the recorded launches of one stage; call after a synchronize."""
        return [s.elapsed_time(e) for s, e in self.events.get(name, [])]

    # ------------------------------------------------------------------ one chunk of sequences
    def _chunk_channels_last(self, cutouts, b0, b1, first, prev, nxt, pred_cls, pred_reg, feat_fused):
        N = self.N
        nb, M, L = b1 - b0, (b1 - b0) * N, self.P // 4
        # every stage writes straight into the step's own tensors (slices along the sequence axis are contiguous)
        if first:                # memory := features; similarities against itself (dr_spaam.py:242-244)
            feat, x_op = self.net.features(cutouts[b0:b1].view(M, self.P), plain_out=nxt[b0:b1])
            emb_x = self.net.embed(x_op, M, out=self.emb_memory[b0:b1]).view(nb, N, -1)
            with self._timed("gate"):
                ops.gate_forward(nxt[b0:b1], nxt[b0:b1], emb_x, emb_x, self.alpha, self.window, out=prev[b0:b1],
                                 feat_out=feat_fused[b0:b1], status=self.status)      # blended output is discarded
            t_op = x_op
        else:
            feat, x_op = self.net.features(cutouts[b0:b1].view(M, self.P))
            feat = feat.view(nb, N, L, -1)
            emb_x = self.net.embed(x_op, M).view(nb, N, -1)
            fused = self.net.gate_emits_operand
            t_op = torch.empty((M * L, 2 * feat.shape[-1]), dtype=torch.float16, device=self.device) if fused else None
            with self._timed("gate"):
                ops.gate_forward(feat, prev[b0:b1], emb_x, self.emb_memory[b0:b1], self.alpha, self.window, out=nxt[b0:b1],
                                 feat_out=feat_fused[b0:b1], split_out=t_op, split_channels=feat.shape[-1],
                                 status=self.status)
            if not fused:
                t_op = self.net.operand(nxt[b0:b1].view(M * L, -1))
                self.kernel_launches += 1 if self.net.split else 0
            # the embedding of the NEW memory is what the next step's gate needs (dr_spaam.py:180-181)
            self.net.embed(t_op, M, out=self.emb_memory[b0:b1])
            self.kernel_launches += 1 if self.net.tc else 0
        self.net.votes(t_op, M, L, pred_cls[b0:b1], pred_reg[b0:b1])
        # cuDNN modes: first layer + 9 activation passes + head + gate; tcgen05 mode: first layer + 10 convolutions
        # + embedding + head + gate
        self.kernel_launches += 14 if self.net.tc else 12

    def _chunk_ncl(self, cutouts, b0, b1, first, prev, nxt, pred_cls, pred_reg, feat_fused):
        N = self.N
        nb = b1 - b0
        y = cutouts[b0:b1].view(nb * N, 1, self.P)
        y = F.max_pool1d(self.block1(y), 2)
        y = F.max_pool1d(self.block2(y), 2)                                                 # [nb*N, 256, L]
        feat = y.view(nb, N, y.shape[-2], y.shape[-1])
        emb_x = self.embed(y).view(nb, N, -1)
        if first:
            # first frame: memory := features, similarities against itself (dr_spaam.py:242-244)
            nxt[b0:b1].copy_(feat)
            with self._timed("gate"):
                _, ff, _ = ops.gate_forward(feat, nxt[b0:b1], emb_x, emb_x, self.alpha, self.window,
                                            out=prev[b0:b1])          # blended output is discarded
        else:
            emb_t = self.embed(prev[b0:b1].view(nb * N, y.shape[-2], y.shape[-1])).view(nb, N, -1)
            with self._timed("gate"):
                _, ff, _ = ops.gate_forward(feat, prev[b0:b1], emb_x, emb_t, self.alpha, self.window,
                                            out=nxt[b0:b1])
        self.kernel_launches += 1
        feat_fused[b0:b1] = ff
        z = nxt[b0:b1].view(nb * N, y.shape[-2], y.shape[-1])
        z = F.max_pool1d(self.block3(z), 2)
        z = self.block4(z)
        z = F.avg_pool1d(z, z.shape[-1])
        pred_cls[b0:b1] = torch.sigmoid(F.conv1d(z, self.w_cls, self.b_cls).view(nb, N))
        pred_reg[b0:b1] = F.conv1d(z, self.w_reg, self.b_reg).view(nb, N, 2)

    # ------------------------------------------------------------------ the step
    @torch.no_grad()
    def step_device(self, scans):
        """One scan per sequence, all on the device.  scans: [B, N] float32 CUDA tensor.

        Returns a dict of device tensors: pred_cls [B, N] (post-sigmoid), pred_reg [B, N, 2],
        feat_fused [B, N, W] and the NMS outputs of `ops.nms_centers`.  With cuda_graph=True the tensors belong to the
        graph of the step's memory parity: they are overwritten two steps later.
        """
        B, N = self.B, self.N
        if tuple(scans.shape) != (B, N):
            raise ValueError("scans must be [%d, %d]" % (B, N))
        if self.cuda_graph and self.steps_done >= 2:           # steps 0 and 1 run eagerly: first-frame branch, lazy set-up
            if scans.data_ptr() != self.g_scans.data_ptr():
                self.g_scans.copy_(scans)
            entry = self._graphs.get(self.cur)
            if entry is None:
                graph = torch.cuda.CUDAGraph()
                before = self.kernel_launches
                torch.cuda.current_stream(self.device).synchronize()
                with torch.cuda.graph(graph, pool=self._graph_pool):
                    res = self._step_body(self.g_scans)
                if self._graph_pool is None:
                    self._graph_pool = graph.pool()
                entry = self._graphs[self.cur] = (graph, res, self.kernel_launches - before)
                self.kernel_launches = before
            graph, res, launches = entry
            graph.replay()
            self.kernel_launches += launches
        else:
            res = self._step_body(scans)
        self.cur = 1 - self.cur
        self.has_memory = True
        self.steps_done += 1
        self._last = res
        return res

    def _step_body(self, scans):
        B, N = self.B, self.N
        prev, nxt = self.memory[self.cur], self.memory[1 - self.cur]
        pred_cls = torch.empty((B, N), dtype=torch.float32, device=self.device)
        pred_reg = torch.empty((B, N, 2), dtype=torch.float32, device=self.device)
        feat_fused = torch.empty((B, N, self.window), dtype=torch.float32, device=self.device)
        first = not self.has_memory
        with self._precision():
            with self._timed("cutout"):      # one launch for all B sequences: [B, N, 1, P]
                cutouts = ops.cutout(scans.unsqueeze(1), self.phi, fast=self.cutout_fast, **self.cutout_kwargs)
            # one CTA per scan does the span reduction, the half-angles and the samples in one launch (both numerics); EXACT on
            # a scan too long for its shared-memory staging (16 B per beam) falls back to the span + pieces kernels
            P = int(self.cutout_kwargs.get("num_cutout_pts", 48))
            one_launch = self.cutout_fast or 16 * (N + 1) + 8 * P + 4 * ((N + 3) & ~3) + 512 * P <= 110 * 1024
            self.kernel_launches += 1 if one_launch else (2 if self.cutout_kwargs.get("area_mode") else 1)
            chunk = self._chunk_channels_last if self.channels_last else self._chunk_ncl
            for b0 in range(0, B, self.seq_chunk):
                chunk(cutouts, b0, min(B, b0 + self.seq_chunk), first, prev, nxt, pred_cls, pred_reg, feat_fused)
            with self._timed("nms"):
                res = ops.nms_centers(scans, self.phi, pred_cls, pred_reg, min_dist=self.min_dist)
            self.kernel_launches += 3
        res.update(pred_cls=pred_cls, pred_reg=pred_reg, feat_fused=feat_fused)
        return res

    def _raise_status(self, code):
        if code == 16:
            raise RuntimeError("an activation left the float16 range of precision='fp32' (|x| > 65504) on %s: "
                               "use precision='fp32-tf32' for this checkpoint" % self.device)
        if code:
            raise RuntimeError("a libpof kernel reported device status %d on %s (a bounded pipeline wait timed out; "
                               "the results of that step are invalid)" % (code, self.device))

    def check(self):
        """Synchronise and raise if a step since the last check left the float16 range or timed out on the device.
        The status word belongs to this detector and is cleared by the read, so a new step starts clean."""
        self._raise_status(ops.read_status(self.status))

    @property
    def template(self):
        """The current attention memory [B, N, 256, L] (what the reference returns as out_template)."""
        m = self.memory[self.cur]
        return m.permute(0, 1, 3, 2) if self.channels_last else m

    def step(self, scans_host):
        """Host-facing call: NumPy/CPU-tensor ranges [B, N] in, detections (NumPy views of pinned
        buffers) out.  Includes the H2D copy of the ranges and the D2H copy of the results."""
        src = torch.as_tensor(scans_host, dtype=torch.float32)
        self.h_scans.copy_(src)
        d_scans = self.g_scans if self.cuda_graph else self.d_scans        # the graphs read their own input buffer
        d_scans.copy_(self.h_scans, non_blocking=True)
        res = self.step_device(d_scans)
        for k, h in self.h_out.items():
            h.copy_(res[k], non_blocking=True)
        self.h_status.copy_(self.status, non_blocking=True)
        torch.cuda.current_stream(self.device).synchronize()
        if int(self.h_status[0]):            # never hand back detections computed from an invalid step
            self.status.zero_()
            self._raise_status(int(self.h_status[0]))
        return {k: v.numpy() for k, v in self.h_out.items()}

    def detections(self, host_result, b):
        """Unpack sequence b of a `step` result into the reference's (det_xys, det_cls, instance_mask)."""
        k = int(host_result["n_keep"][b])
        return (host_result["det_xy"][b, :k].copy(), host_result["det_cls"][b, :k].reshape(k, 1).copy(),
                host_result["instance_mask"][b].copy())
