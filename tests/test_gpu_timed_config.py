"""Parity of the CUDA path AT THE SHAPES bench.py TIMES (BASELINE.json configs[2]): launches large enough that every
persistent CTA of the tcgen05 convolution walks many tiles (ring wrap-around, tensor-memory double-buffer phase
flips, resident-weight reuse), and the streaming engine on JRDB-shaped scans with more sequences than one backbone
chunk.  Needs a B200: run with `-m gpu`."""
import numpy as np
import pytest
import torch
import torch.nn.functional as F

from oracle import cutout as ocut
from oracle import model as omodel
from oracle import nms as onms
from planar_optical_flow_b200 import ops, synth
from planar_optical_flow_b200.model import SpatialDROW
from tests.helpers import REL_TOL, assert_rel, f64_state_dict, rel_err

pytestmark = pytest.mark.gpu

CFG = dict(fixed=True, centered=True, window_width=1.0, window_depth=0.5,
           num_cutout_pts=56, padding_val=29.99, area_mode=True)
STREAM_W = 0x20000          # include/pof.h POF_CONV_TC_STREAM_W
SINGLE_CTA = 0x10000        # include/pof.h POF_CONV_TC_SINGLE_CTA
NO_SPLIT = 0x80000          # include/pof.h POF_CONV_TC_NO_SPLIT_TILE
HALO = 0x100000             # include/pof.h POF_CONV_TC_HALO


@pytest.fixture(scope="module", autouse=True)
def strict_fp32():
    old = (torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32)
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False
    yield
    torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32 = old


def _tc_weights(w, f16=True):
    from planar_optical_flow_b200.engine import _ChannelsLastBackbone

    holder = _ChannelsLastBackbone.__new__(_ChannelsLastBackbone)
    holder.f16 = f16
    return holder._tc_weight(w)


# every layer shape of the engine: (LA, Cin, Cout, taps, pad, pool) -> instantiation <BN, 2, F16>
ENGINE_LAYERS = [
    (56, 64, 64, 3, 1, 1),       # <64>   resident weights
    (56, 64, 128, 3, 1, 2),      # <128>  resident weights, pooled
    (28, 128, 128, 3, 1, 1),     # <128>
    (28, 128, 256, 3, 1, 2),     # <256>
    (14, 256, 256, 3, 1, 1),     # <256>
    (14, 256, 512, 3, 1, 2),     # <256>  two column tiles: the dominant kernel of the bench
    (7, 512, 256, 3, 1, 1),      # <256>
    (7, 256, 128, 3, 1, 1),      # <128>
    (14, 256, 128, 14, 0, 1),    # <128>  the gate embedding: one GEMM over whole rows
]


@pytest.mark.parametrize("LA,Cout,pool", [(56, 64, 1), (56, 128, 2), (56, 128, 1), (40, 64, 2), (62, 128, 1)])
@pytest.mark.parametrize("M", [1, 2, 3, 4, 5, 130, 2001])
def test_conv_tc_halo_tiles_are_bit_equal(M, LA, Cout, pool):
    """With POF_CONV_TC_HALO a 64-channel k = 3 layer loads a tile ONCE, with the two padding rows of every cutout, and runs the
    three taps on row-shifted views of it (instead of three loads): same MMAs in the same order, so the same bits as the default
    form, for one and two cutouts per tile, odd cutout counts, with and without the pool; and the float64 convolution to 1.5e-6."""
    Cin, taps, pad = 64, 3, 1
    g = torch.Generator(device="cuda").manual_seed(M * 13 + LA + Cout)
    x = torch.randn(M, LA, Cin, generator=g, device="cuda")
    w = torch.randn(Cout, Cin, taps, generator=g, device="cuda") * (2.0 / (Cin * taps)) ** 0.5
    b = torch.randn(Cout, generator=g, device="cuda") * 0.1
    _, a = ops.act(x.view(M * LA, Cin), None, pool=1, slope=1.0, want_plain=False, want_split=True, parts=ops.SPLIT_F16)
    ws, out_scale = _tc_weights(w)
    outs = []
    for flags in (HALO, 0):
        status = ops.new_status(x.device)
        plain, split = ops.conv_tc(a, ws, b, M, LA, LA, taps, pad, pool=pool, slope=0.1, want_plain=True, want_split=True,
                                   out_scale=out_scale, chain_channels=flags, status=status)
        assert ops.read_status(status) == 0
        outs.append((plain.clone(), split.clone()))
    y = F.conv1d(x.permute(0, 2, 1).double(), w.double(), b.double(), padding=pad)
    if pool == 2:
        y = F.max_pool1d(y, 2)
    want = torch.where(y > 0, y, y * 0.1).permute(0, 2, 1)
    assert_rel(outs[0][0].view(M, LA // pool, Cout).double().cpu(), want.cpu(), tol=1.5e-6, what="halo tiles vs fp64")
    assert torch.equal(outs[0][0], outs[1][0]), int((outs[0][0] != outs[1][0]).sum())
    assert torch.equal(outs[0][1].view(torch.int16), outs[1][1].view(torch.int16))


@pytest.mark.parametrize("LA", [12, 48, 36])
@pytest.mark.parametrize("M", [1, 3, 5, 6, 7, 20, 21, 22, 130, 1001])
def test_conv_tc_split_tiles_other_row_counts(M, LA):
    """The same for the other cutout heights that leave room for one more cutout per SM pair (num_cutout_pts = 48 gives 48- and
    12-row layers: 5 and 21 cutouts per pair; 36 rows: 7): the peer's part of the shared cutout can start anywhere in its tile,
    including on a warp boundary."""
    Cin, Cout, taps, pad, pool = 64, 128, 3, 1, 2
    g = torch.Generator(device="cuda").manual_seed(M * 11 + LA)
    x = torch.randn(M, LA, Cin, generator=g, device="cuda")
    w = torch.randn(Cout, Cin, taps, generator=g, device="cuda") * (2.0 / (Cin * taps)) ** 0.5
    b = torch.randn(Cout, generator=g, device="cuda") * 0.1
    _, a = ops.act(x.view(M * LA, Cin), None, pool=1, slope=1.0, want_plain=False, want_split=True, parts=ops.SPLIT_F16)
    ws, out_scale = _tc_weights(w)
    outs = []
    for flags in (0, NO_SPLIT):
        status = ops.new_status(x.device)
        plain, split = ops.conv_tc(a, ws, b, M, LA, LA, taps, pad, pool=pool, slope=0.1, want_plain=True, want_split=True,
                                   out_scale=out_scale, chain_channels=flags, status=status)
        assert ops.read_status(status) == 0
        outs.append((plain.clone(), split.clone()))
    assert torch.equal(outs[0][0], outs[1][0]), int((outs[0][0] != outs[1][0]).sum())
    assert torch.equal(outs[0][1].view(torch.int16), outs[1][1].view(torch.int16))
    y = F.max_pool1d(F.conv1d(x.permute(0, 2, 1).double(), w.double(), b.double(), padding=pad), 2)
    want = torch.where(y > 0, y, y * 0.1).permute(0, 2, 1)
    assert_rel(outs[0][0].view(M, LA // pool, Cout).double().cpu(), want.cpu(), tol=1.5e-6, what="split tiles vs fp64")


@pytest.mark.parametrize("Cout,pool", [(128, 1), (256, 2), (512, 1)])
@pytest.mark.parametrize("M", [1, 4, 5, 8, 9, 10, 17, 130, 4001])
def test_conv_tc_split_tiles_are_bit_equal(M, Cout, pool):
    """Lout = 28: an SM pair's 256 accumulator rows take nine cutouts, the middle one split between the two CTAs (252 busy rows
    instead of 224).  Tiling must not change a bit: same outputs, plain and operand split, as with POF_CONV_TC_NO_SPLIT_TILE, for
    cutout counts around every multiple of 9 and 8, one and two column tiles, with and without the pool."""
    LA, Cin, taps, pad = 28, 128, 3, 1
    g = torch.Generator(device="cuda").manual_seed(M * 7 + Cout)
    x = torch.randn(M, LA, Cin, generator=g, device="cuda")
    w = torch.randn(Cout, Cin, taps, generator=g, device="cuda") * (2.0 / (Cin * taps)) ** 0.5
    b = torch.randn(Cout, generator=g, device="cuda") * 0.1
    _, a = ops.act(x.view(M * LA, Cin), None, pool=1, slope=1.0, want_plain=False, want_split=True, parts=ops.SPLIT_F16)
    ws, out_scale = _tc_weights(w)
    outs = []
    for flags in (0, NO_SPLIT):
        status = ops.new_status(x.device)
        plain, split = ops.conv_tc(a, ws, b, M, LA, LA, taps, pad, pool=pool, slope=0.1, want_plain=True, want_split=True,
                                   out_scale=out_scale, chain_channels=flags, status=status)
        assert ops.read_status(status) == 0
        outs.append((plain.clone(), split.clone()))
    assert torch.equal(outs[0][0], outs[1][0]), int((outs[0][0] != outs[1][0]).sum())
    assert torch.equal(outs[0][1].view(torch.int16), outs[1][1].view(torch.int16))
    y = F.conv1d(x.permute(0, 2, 1).double(), w.double(), b.double(), padding=pad)
    if pool == 2:
        y = F.max_pool1d(y, 2)
    want = torch.where(y > 0, y, y * 0.1).permute(0, 2, 1)
    assert_rel(outs[0][0].view(M, LA // pool, Cout).double().cpu(), want.cpu(), tol=1.5e-6, what="split tiles vs fp64")


@pytest.mark.parametrize("flags", [0, STREAM_W, SINGLE_CTA], ids=["default", "streamed-weights", "single-cta"])
@pytest.mark.parametrize("LA,Cin,Cout,taps,pad,pool", ENGINE_LAYERS)
def test_conv_tc_many_tiles_per_cta(LA, Cin, Cout, taps, pad, pool, flags):
    """M = 20,480 cutouts: 70-280 tiles per CTA pair.  Checked against (1) an fp64 convolution on a sample of cutouts that
    includes the first and last tile and tile-boundary neighbours, at 1.5e-6, and (2) a strict-fp32 cuDNN convolution of
    EVERY row at 6e-6 (so a wrong or missing tile anywhere fails)."""
    if flags == STREAM_W and not (Cin == 64):
        pytest.skip("weights are streamed by default for this shape")
    if flags == SINGLE_CTA and (LA, Cin, Cout) not in ((56, 64, 64), (14, 256, 512), (14, 256, 128)):
        pytest.skip("single-CTA form: one shape per instantiation is enough")
    M = 20480
    g = torch.Generator(device="cuda").manual_seed(LA * 1000 + Cin + Cout)
    x = torch.randn(M, LA, Cin, generator=g, device="cuda").abs() * torch.rand(M, LA, Cin, generator=g, device="cuda")
    x = torch.where(torch.rand(M, LA, Cin, generator=g, device="cuda") < 0.3, -0.1 * x, x)
    w = torch.randn(Cout, Cin, taps, generator=g, device="cuda") * (2.0 / (Cin * taps)) ** 0.5
    b = torch.randn(Cout, generator=g, device="cuda") * 0.1
    Lout = LA if pad else LA - taps + 1
    _, a = ops.act(x.view(M * LA, Cin), None, pool=1, slope=1.0, want_plain=False, want_split=True, parts=ops.SPLIT_F16)
    ws, out_scale = _tc_weights(w)
    status = ops.new_status(x.device)
    plain, split = ops.conv_tc(a, ws, b, M, LA, Lout, taps, pad, pool=pool, slope=0.1, want_plain=True, want_split=True,
                               out_scale=out_scale, chain_channels=flags, status=status)
    assert ops.read_status(status) == 0
    rows = Lout // pool
    plain = plain.view(M, rows, Cout)

    def ref(xs, dtype):
        y = F.conv1d(xs.permute(0, 2, 1).to(dtype), w.to(dtype), b.to(dtype), padding=pad)
        if pool == 2:
            y = F.max_pool1d(y, 2)
        return torch.where(y > 0, y, y * 0.1).permute(0, 2, 1)

    mt = 128 // Lout                                                  # cutouts per tile
    sample = sorted(set([0, 1, mt - 1, mt, 2 * mt - 1, 2 * mt, M // 2, M - mt - 1, M - mt, M - 1] +
                        [int(v) for v in torch.randint(0, M, (54,), generator=torch.Generator().manual_seed(7))]))
    idx = torch.tensor(sample, device="cuda")
    want64 = ref(x[idx], torch.float64)
    assert_rel(plain[idx].double().cpu(), want64.cpu(), tol=1.5e-6, what="sampled cutouts vs fp64")
    scale = float(want64.abs().max())
    worst = 0.0
    for m0 in range(0, M, 4096):                                      # every row, strict fp32 on cuDNN
        want32 = ref(x[m0:m0 + 4096], torch.float32)
        worst = max(worst, float((plain[m0:m0 + 4096] - want32).abs().max()) / scale)
    assert worst <= 6e-6, "some row differs from the fp32 convolution by %.3g of the output range" % worst
    assert_rel((split[:, :Cout].double() + split[:, Cout:].double()).cpu(), plain.view(-1, Cout).double().cpu(), tol=2.0 ** -21,
               what="output split")


def test_streaming_engine_jrdb_shape_across_chunks():
    """StreamingDetector exactly as bench.py builds it (JRDB-shaped scans: 1091 points, float32 angles; default
    precision and chunking) with B = 130 sequences, which spans two backbone chunks, over 4 steps.  Three sequences
    (first, the one either side of the chunk boundary, last) are replayed through the oracle's reference loop in
    float32 AND in float64: the engine must be within 1e-5 of the float32 oracle, or - where the float32 oracle itself
    is further than that from the float64 truth - at least as close to the truth as the oracle is.
    The network is compared on identical inputs (the oracle is fed the cutout kernel's output for the same ranges; the
    kernel's own parity is tests/test_gpu_parity.py's subject); the whole chain from the oracle's NumPy cutouts is
    checked too for every sequence whose device cutouts have matched NumPy's at every sample so far (a <= 2-ulp
    half-angle difference can flip the nearest beam of an area-resampled sample, which changes that sample by up to the
    clip range and everything downstream of it)."""
    from planar_optical_flow_b200.engine import StreamingDetector

    B, steps = 130, 4
    phi = synth.phi_for("jrdb")
    assert phi.dtype == np.float32 and len(phi) == 1091
    n = len(phi)
    base = [synth.structured_sequence(steps, n, seed=300 + k, phi=phi) for k in range(8)]
    rs = np.random.RandomState(5)
    scans = np.stack([np.clip(base[b % 8] + rs.normal(0, 0.02, (steps, n)).astype(np.float32), 0.05, 29.99)
                      for b in range(B)], axis=1)                      # [T, B, N]
    sd = omodel.randomize_bn_stats(omodel.init_state_dict(56, True, seed=11))
    m = SpatialDROW(num_scans=10, num_pts=56, alpha=0.5, window_size=11, pedestrian_only=True)
    m.load_state_dict(sd, strict=True)
    det = StreamingDetector(m.cuda(), phi, CFG, B)
    assert det.seq_chunk < B, "the test must span more than one chunk"
    picks = [0, det.seq_chunk - 1, det.seq_chunk, B - 1]
    sd64 = f64_state_dict(sd)
    tmpl32 = {k: None for k in picks}
    tmpl64 = {k: None for k in picks}
    tmpl_np = {k: None for k in picks}
    same_input, n_e2e = {k: True for k in picks}, 0
    report = {}
    phi_d = torch.from_numpy(phi).cuda()
    for t in range(steps):
        host = det.step(scans[t])
        ct_dev = ops.cutout(torch.from_numpy(scans[t, picks]).cuda().unsqueeze(1), phi_d, **CFG).cpu().numpy()     # what the engine saw
        for j, k in enumerate(picks):
            ct = ct_dev[j]
            with torch.no_grad():
                ct_np = ocut.scans_to_cutout(scans[t, k][None], phi, **CFG)
                c_np, r_np, tmpl_np[k], _ = omodel.spatial_drow_stream(torch.from_numpy(ct_np)[None], sd, 0.5, 11, tmpl_np[k])
                c32, r32, tmpl32[k], f32 = omodel.spatial_drow_stream(torch.from_numpy(ct)[None], sd, 0.5, 11, tmpl32[k])
                c64, r64, tmpl64[k], f64 = omodel.spatial_drow_stream(torch.from_numpy(ct)[None].double(), sd64, 0.5, 11, tmpl64[k])
            got = {"scores": det._last["pred_cls"][k].cpu().numpy().reshape(-1, 1),
                   "votes": det._last["pred_reg"][k].cpu().numpy(),
                   "similarities": det._last["feat_fused"][k].cpu().numpy(),
                   "memory": det.template[k].cpu().numpy()}
            want32 = {"scores": torch.sigmoid(c32[0]).numpy(), "votes": r32[0].numpy(), "similarities": f32[0].numpy(),
                      "memory": tmpl32[k][0].numpy()}
            want64 = {"scores": torch.sigmoid(c64[0]).numpy(), "votes": r64[0].numpy(), "similarities": f64[0].numpy(),
                      "memory": tmpl64[k][0].numpy()}
            same_input[k] = same_input[k] and float(np.abs(ct - ct_np).max()) <= REL_TOL
            if same_input[k]:        # no nearest-beam flip from a 1-ulp half-angle so far: the whole chain from NumPy's cutouts agrees too
                n_e2e += 1
                for name, w_np in (("scores", torch.sigmoid(c_np[0]).numpy()), ("votes", r_np[0].numpy()), ("memory", tmpl_np[k][0].numpy())):
                    e = rel_err(got[name], w_np)
                    assert e <= 2 * REL_TOL or rel_err(got[name], want64[name]) <= max(REL_TOL, rel_err(w_np, want64[name])), \
                        "end-to-end %s step %d seq %d: %.3g" % (name, t, k, e)
            for name in got:
                e_ref = rel_err(got[name], want32[name])              # engine vs the reference's own float32 path
                e_true = rel_err(got[name], want64[name])             # engine vs float64 truth
                o_true = rel_err(want32[name], want64[name])          # reference float32 path vs float64 truth
                report[name] = max(report.get(name, (0, 0, 0)), (e_ref, e_true, o_true))
                assert e_ref <= REL_TOL or e_true <= max(REL_TOL, o_true), \
                    "%s step %d seq %d: engine-vs-oracle %.3g, engine-vs-fp64 %.3g, oracle-vs-fp64 %.3g" % (name, t, k, e_ref, e_true, o_true)
                assert e_true <= 1.5 * REL_TOL, "%s step %d seq %d: %.3g from the float64 result" % (name, t, k, e_true)
            # detections: bit-exact against the NMS specification on the engine's own scores
            xy, c, mask = det.detections(host, k)
            mine = onms.nms_sweep_spec(scans[t, k], phi, got["scores"], got["votes"])
            assert np.array_equal(mask, mine["instance_mask"])
            assert np.array_equal(host["keep_idx"][k, :len(xy)], mine["keep_idx"])
            want = onms.nms_sweep_spec(scans[t, k], phi, want32["scores"], want32["votes"])
            if np.array_equal(mine["order"], want["order"]) and want["margin"] > 1e-4:
                assert np.array_equal(mask, want["instance_mask"])
    det.check()
    print("engine parity at the bench shape (engine-vs-oracle32, engine-vs-fp64, oracle32-vs-fp64):", report,
          "; end-to-end checks on identical cutouts: %d of %d" % (n_e2e, steps * len(picks)))
