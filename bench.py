#!/usr/bin/env python
"""Benchmark of the DR-SPAAM per-point scan hot path on B200 (see BASELINE.json).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]
                    [--sequences 256] [--shape jrdb|drow] [--precision fp32|fp32-tf32|fp32-simt|tf32x3|tf32]
                    [--scaling weak|strong] [--graph] [--workload stream|train]

Other modes (same JSON contract, `config.workload` says which):
  --scaling strong      BASELINE.json configs[2] as written: `--sequences` (256) sequences IN TOTAL, sharded round-robin over
                        the ranks (planar_optical_flow_b200.parallel.shard_sequences); default is weak (256 per GPU)
  --sequences 1|8 ...   latency mode: the line carries `latency` = ms per step = ms per scan of a sequence.  Every run replays
                        each step as one CUDA graph per memory parity (`StreamingDetector(cuda_graph=True)`; `--no-graph`
                        issues the launches from Python: 48.2 instead of 46.9 ms per step of 256 scans, 0.77 instead of 0.51 ms
                        for one sequence); the per-stage CUDA events come from an extra eager pass over the same steps
  --workload train      BASELINE.json configs[3]: the training step of bin/train_dr_spaam.py (SpatialDROW, per-GPU batch 8 x
                        11 scans x 450 points, cutouts on the device, fused gate forward/backward, Adam; DDP all-reduce
                        over NCCL when N > 1); metric = training samples/s

Workload (BASELINE.json configs[2], the one the metric is quoted on): DR-SPAAM streaming
inference with spatial-attention memory over 256 independent JRDB-shaped sequences
(1091 points, 360 deg) per GPU; one "step" = one scan of every sequence through
cutout -> conv backbone -> attention memory update -> heads -> sigmoid -> NMS, memory carried.

One JSON line on stdout (rank 0):
  value      scans/s, whole job, ranges already resident in HBM when the timed region starts
  e2e        the same through StreamingDetector.step(host buffers): H2D of the ranges and D2H
             of the detections inside the timed region
  roofline   the dominant kernel of the step, the tcgen05 convolution (conv_tc_kernel<256,2>): algorithmic
             fp32 FLOPs (45.9 MFLOP per point) / CUDA-event time around every launch
             inside the timed steps, against the measured sustained bf16 peak (kind::f16 runs at the bf16 rate; the
             kernel issues three MMAs per algorithmic product, reported beside it as mma_tflops)
  roofline_gate / roofline_cutout   the two HBM-bound hot-path kernels: algorithmic bytes (SURVEY.md
             §8d: N*44,076 B per sequence-step, N*228 B per scan row) / CUDA-event launch time,
             against the measured copy bandwidth (MEASURED_PEAKS.json)
             (`traffic` = ncu dram bytes of one launch of the current build, profiles/r2_ncu_traffic.json; the gate line also
             carries the figure that counts the float16 operand split the kernel writes in the same pass)
  parity_spot   two of the timed sequences replayed through the oracle's reference loop AFTER the timed region: max error of
             the last step's scores / votes / memory relative to each tensor's magnitude, NMS masks equal
  cpu_baseline  the UNMODIFIED reference (baseline/_ref: its scans_to_cutout, SpatialDROW module, nms_predicted_center;
             kind "reference") - or, if that directory is missing, the oracle port of the same algorithm (kind "port") -
             timed on this box's host cores on a bounded sample of the same workload
`--impl reference` times that CPU path alone, as the reference arm.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

CUTOUT_KW = dict(fixed=True, centered=True, window_width=1.0, window_depth=0.5, num_cutout_pts=56,
                 padding_val=29.99, area_mode=True)          # config/dr_spaam.yaml:21-28
ALPHA, WINDOW = 0.5, 11                                      # config/dr_spaam.yaml:16-18
METRIC = "DR-SPAAM scans/sec (JRDB 1091-pt)"
FALLBACK_HBM_GBS = 6650.0                                    # /opt/skills/guides/B200_PROFILING.md


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=8)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--sequences", type=int, default=256, help="sequences per GPU")
    ap.add_argument("--shape", default="jrdb", choices=["jrdb", "drow"])
    ap.add_argument("--precision", default="fp32", choices=["fp32", "fp32-tf32", "fp32-simt", "tf32x3", "tf32"])
    ap.add_argument("--cpu-scans", type=int, default=24, help="scans in the CPU-baseline sample")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-parity-spot", action="store_true", help="skip the oracle replay of two of the timed sequences")
    ap.add_argument("--extra-precisions", default="", help="comma list of further engine precisions to time (device only)")
    ap.add_argument("--seq-chunk", type=int, default=0, help="sequences per backbone chunk (0 = engine default)")
    ap.add_argument("--scaling", default="weak", choices=["weak", "strong"],
                    help="weak: --sequences per GPU; strong: --sequences in total, sharded over the ranks")
    ap.add_argument("--graph", dest="graph", action="store_true", default=True,
                    help="replay each step as one CUDA graph per memory parity (default; +2.6 %% at 256 sequences, 1.5x at one)")
    ap.add_argument("--no-graph", dest="graph", action="store_false", help="issue every launch of a step from Python")
    ap.add_argument("--workload", default="stream", choices=["stream", "train", "cutout"],
                    help="stream: the headline metric; train: BASELINE configs[3]; cutout: the scans_to_cutout-only sweep of configs[1]")
    return ap.parse_args()


# ----------------------------------------------------------------------------- synthetic workload
def make_sequences(shape, n_seq, n_steps, seed0):
    """[n_steps, n_seq, N] float32 ranges; sequence id seeds the generator."""
    import numpy as np

    from planar_optical_flow_b200 import synth

    phi = synth.phi_for(shape)
    n = len(phi)
    # a handful of structured walks, tiled with per-sequence noise: generation cost stays bounded
    base = [synth.structured_sequence(n_steps, n, seed=seed0 + k, phi=phi) for k in range(min(n_seq, 16))]
    rs = np.random.RandomState(seed0)
    out = np.empty((n_steps, n_seq, n), dtype=np.float32)
    for b in range(n_seq):
        jitter = rs.normal(0.0, 0.02, size=(n_steps, n)).astype(np.float32)
        out[:, b] = np.clip(base[b % len(base)] + jitter, 0.05, 29.99)
    return phi, out


def build_model(seed=0):
    import torch

    from planar_optical_flow_b200.model import SpatialDROW

    torch.manual_seed(seed)
    m = SpatialDROW(num_scans=10, num_pts=CUTOUT_KW["num_cutout_pts"], alpha=ALPHA, window_size=WINDOW,
                    pedestrian_only=True)
    # non-trivial BN statistics so folding is exercised (random-init weights, no checkpoint offline)
    g = torch.Generator().manual_seed(seed + 1)
    for mod in m.modules():
        if isinstance(mod, torch.nn.BatchNorm1d):
            mod.running_mean.copy_(torch.randn(mod.running_mean.shape, generator=g) * 0.1)
            mod.running_var.copy_(torch.rand(mod.running_var.shape, generator=g) * 0.5 + 0.75)
    return m.eval()


def stream_config(args, B, total_seqs, N, world, use_graph):
    """The `config` object of the streaming workload: ONE definition for our arm and the reference arm (which times a bounded
    sample of the same workload on the host cores and says so in `cpu_baseline.sample`)."""
    return {"workload": "DR-SPAAM streaming inference, %s %s-shaped sequences (%d pts), "
                        "cutout+backbone+attention memory+heads+NMS per scan"
                        % ("%d independent sequences in total, sharded round-robin over %d GPU(s)," % (total_seqs, world)
                           if args.scaling == "strong" else "%d independent sequences per GPU," % B, args.shape.upper(), N),
            "sequences_per_gpu": B, "sequences_total": total_seqs, "cuda_graph": use_graph, "points": N,
            "cutout_pts": CUTOUT_KW["num_cutout_pts"], "window": WINDOW, "alpha": ALPHA, "precision": args.precision,
            "weights": "random-init",
            "l2_policy": "inputs larger than L2: per step the path streams %.1f GB of attention memory "
                         "and features (L2 = 126 MB)" % (3 * B * N * 3584 * 4 / 1e9)}


# ----------------------------------------------------------------------------- CPU reference path
def reference_modules():
    """(utils module, dr_spaam module) of the UNMODIFIED reference (baseline/_ref, installed by baseline/install_reference.py
    from the build container's read-only mount), or None when it is not there (then the oracle port is timed)."""
    try:
        from oracle import ref_shim

        if ref_shim.available():
            return ref_shim.load()
    except Exception as e:      # noqa: BLE001
        sys.stderr.write("reference tree not usable (%r): timing the oracle port\n" % (e,))
    return None


def cpu_reference_kind():
    return "reference" if reference_modules() is not None else "port"


def cpu_reference_scans_per_s(shape, n_scans, warmup=2, sequences=1, model_device="cpu"):
    """The reference's own code on the host cores when baseline/_ref holds it (its `scans_to_cutout`, its `SpatialDROW`
    module, its `nms_predicted_center`, driven like depracted_scripts/infer_person_flow.py:101-139), else the oracle port of
    the same algorithm (oracle/); all threads torch can use.

    Streams `sequences` independent sequences one scan at a time exactly like the reference loop
    (depracted_scripts/infer_person_flow.py:101-139): NumPy cutout -> torch SpatialDROW with
    dense attention -> sigmoid -> NumPy NMS.  model_device="cuda" is the reference's PyTorch-GPU
    path (:66-77,129-135: cutout and NMS stay on the host, the network runs on the GPU in strict
    fp32, one H2D and one D2H per scan) - a reported baseline like the CPU one.
    """
    import numpy as np
    import torch

    from oracle import cutout as ocut
    from oracle import model as omodel
    from oracle import nms as onms

    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    model = build_model()
    sd = {k: v.detach().clone().to(model_device) for k, v in model.state_dict().items()}
    on_gpu = model_device != "cpu"
    ref = reference_modules()
    if ref is not None:
        ref_utils, ref_net = ref
        ref_model = ref_net.SpatialDROW(num_scans=10, num_pts=CUTOUT_KW["num_cutout_pts"], alpha=ALPHA, window_size=WINDOW,
                                        pedestrian_only=True)
        ref_model.load_state_dict(model.state_dict(), strict=True)
        ref_model = ref_model.to(model_device).eval()

        def cutout_fn(scan, phi):
            return ref_utils.scans_to_cutout(scan[None], phi, stride=1, **CUTOUT_KW)

        def model_fn(ct, tmpl):
            return ref_model(torch.from_numpy(ct)[None].to(model_device), testing=True, fea_template=tmpl)

        nms_fn = ref_utils.nms_predicted_center
    else:
        def cutout_fn(scan, phi):
            return ocut.scans_to_cutout(scan[None], phi, stride=1, **CUTOUT_KW)

        def model_fn(ct, tmpl):
            return omodel.spatial_drow_stream(torch.from_numpy(ct)[None].to(model_device), sd, ALPHA, WINDOW, tmpl)

        nms_fn = onms.nms_predicted_center
    if on_gpu:
        tf32_was = (torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32)
        torch.backends.cudnn.allow_tf32 = torch.backends.cuda.matmul.allow_tf32 = False
    per_seq = -(-(n_scans + warmup * sequences) // sequences)
    phi, scans = make_sequences(shape, sequences, per_seq, seed0=900)
    tmpl = [None] * sequences
    done, t0 = 0, None
    stage = {"cutout": 0.0, "model": 0.0, "nms": 0.0}
    with torch.no_grad():
        for t in range(per_seq):
            for b in range(sequences):
                if done == warmup * sequences:
                    t0 = time.perf_counter()
                    stage = {k: 0.0 for k in stage}
                a = time.perf_counter()
                ct = cutout_fn(scans[t, b], phi)
                c = time.perf_counter()
                cls, reg, tmpl[b], _ = model_fn(ct, tmpl[b])
                conf = torch.sigmoid(cls[0]).cpu().numpy()
                reg_h = reg[0].cpu().numpy()
                d = time.perf_counter()
                nms_fn(scans[t, b], phi, conf, reg_h)
                e = time.perf_counter()
                stage["cutout"] += c - a
                stage["model"] += d - c
                stage["nms"] += e - d
                done += 1
    timed = done - warmup * sequences
    dt = time.perf_counter() - t0
    if on_gpu:
        torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32 = tf32_was
    return timed / dt, cores, timed, {k: 1e3 * v / timed for k, v in stage.items()}


def parity_spot(det, phi, scans, picks):
    """Replay `picks` of the sequences the detector just streamed through the oracle's reference loop (NumPy cutout ->
    torch-CPU SpatialDROW with dense attention -> sigmoid -> NumPy NMS; the checker, after the timed region) and compare
    the detector's LAST step: scores, votes, attention memory (max error relative to the tensor's magnitude) and the
    NMS instance masks.  scans: [T, B, N] = every step the detector has seen."""
    import numpy as np
    import torch

    from oracle import cutout as ocut
    from oracle import model as omodel
    from oracle import nms as onms

    from planar_optical_flow_b200 import ops

    sd = {k: v.detach().cpu() for k, v in build_model().state_dict().items()}
    worst = {"scores": 0.0, "votes": 0.0, "memory": 0.0, "similarities": 0.0}
    worst_same = dict(worst)
    mask_self = mask_ref = True
    phi_d = torch.from_numpy(np.ascontiguousarray(phi)).to(det.device)
    for b in picks:
        tmpl = tmpl_s = None
        ct_dev = ops.cutout(torch.from_numpy(np.ascontiguousarray(scans[:, b])).to(det.device).unsqueeze(1), phi_d,
                            **CUTOUT_KW).cpu()                       # [T, N, 1, P]: the cutouts the detector computed for this sequence
        with torch.no_grad():
            for t in range(scans.shape[0]):
                ct = ocut.scans_to_cutout(scans[t, b][None], phi, stride=1, **CUTOUT_KW)
                cls, reg, tmpl, ff = omodel.spatial_drow_stream(torch.from_numpy(ct)[None], sd, ALPHA, WINDOW, tmpl)
                cls_s, reg_s, tmpl_s, ff_s = omodel.spatial_drow_stream(ct_dev[t][None], sd, ALPHA, WINDOW, tmpl_s)
        conf = torch.sigmoid(cls[0]).numpy()
        got = {"scores": det._last["pred_cls"][b].cpu().numpy().reshape(-1, 1), "votes": det._last["pred_reg"][b].cpu().numpy(),
               "memory": det.template[b].cpu().numpy(), "similarities": det._last["feat_fused"][b].cpu().numpy()}
        want = {"scores": conf, "votes": reg[0].numpy(), "memory": tmpl[0].numpy(), "similarities": ff[0].numpy()}
        same = {"scores": torch.sigmoid(cls_s[0]).numpy(), "votes": reg_s[0].numpy(), "memory": tmpl_s[0].numpy(), "similarities": ff_s[0].numpy()}
        for k in worst:
            worst[k] = max(worst[k], float(np.abs(got[k].astype(np.float64) - want[k]).max() / np.abs(want[k]).max()))
            worst_same[k] = max(worst_same[k], float(np.abs(got[k].astype(np.float64) - same[k]).max() / np.abs(same[k]).max()))
        mask = det._last["instance_mask"][b].cpu().numpy()
        mine = onms.nms_sweep_spec(scans[-1, b], phi, got["scores"], got["votes"])       # NMS spec on the detector's own scores
        ref = onms.nms_predicted_center(scans[-1, b], phi, conf, reg[0].numpy())[2]       # the reference loop end to end
        mask_self = mask_self and bool(np.array_equal(mask, mine["instance_mask"]))
        mask_ref = mask_ref and bool(np.array_equal(mask, ref))
    return {"sequences": list(picks), "steps_replayed": int(scans.shape[0]), "max_rel": max(worst.values()), "per_tensor": worst,
            "max_rel_network_on_same_cutouts": max(worst_same.values()), "per_tensor_network_on_same_cutouts": worst_same,
            "mask_equal": mask_self, "mask_equal_reference_loop": mask_ref,
            "note": "max_rel: detector vs the oracle's float32 reference loop (NumPy cutout included) on the same ranges after the same "
                    "history, relative to each tensor's magnitude (bar 1e-5; tests/test_gpu_timed_config.py arbitrates with float64); "
                    "network_on_same_cutouts: the same with the oracle network fed the cutout kernel's own output, i.e. without the "
                    "<= 2-ulp difference between NumPy's float32 arctan and the device's in the window half-angles; mask_equal: "
                    "device NMS == the NMS specification on the detector's own scores (bit-exact bar); "
                    "mask_equal_reference_loop: == the reference loop's masks (differs only if two scores or a distance "
                    "sit within float32 noise of each other)"}


# ----------------------------------------------------------------------------- clocks
class ClockSampler(threading.Thread):
    Q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        super().__init__(daemon=True)
        self.index, self.samples, self.reasons, self.max_mhz = index, [], set(), None
        self._stop_evt = threading.Event()

    def run(self):
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        while not self._stop_evt.is_set():
            try:
                out = subprocess.run(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q,
                                      "--format=csv,noheader,nounits"], capture_output=True, text=True, timeout=5).stdout
                f = [x.strip() for x in out.strip().split(",")]
                self.samples.append(float(f[0]))
                self.max_mhz = float(f[1])
                for n, v in zip(names, f[2:6]):
                    if v.lower().startswith("active"):
                        self.reasons.add(n)
            except Exception:      # noqa: BLE001  (nvidia-smi hiccup: skip the sample)
                pass
            self._stop_evt.wait(0.2)

    def stop(self):
        self._stop_evt.set()
        self.join(timeout=6)
        s = sorted(self.samples)
        return {"sm_mhz": s[len(s) // 2] if s else None, "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons),
                "samples": len(s)}


def measured_hbm_peak():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
    except Exception:      # noqa: BLE001
        return FALLBACK_HBM_GBS, "fallback (B200_PROFILING.md)"


def ncu_traffic(key):
    """dram__bytes_read.sum + dram__bytes_write.sum of one launch of a hot-path kernel, from the `ncu --set full` capture of the
    CURRENT build summarised in profiles/r2_ncu_traffic.json (tools/ncu_traffic.py); None when that launch shape was not captured."""
    try:
        with open(os.path.join(ROOT, "profiles", "r2_ncu_traffic.json")) as f:
            e = json.load(f).get(key)
        return None if e is None else {"bytes": e["traffic_bytes"], "source": "profiles/r2_ncu_traffic.json:%s (%s)" % (key, e["report"])}
    except Exception:      # noqa: BLE001
        return None


def measured_tensor_peak():
    """Dense bf16 TFLOP/s sustained inside a long step (MEASURED_PEAKS.json), else the recipe's fallback."""
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            d = json.load(f)
            return float(d.get("bf16_tflops_sustained", d["bf16_tflops"])), "measured bf16 sustained (MEASURED_PEAKS.json)"
    except Exception:      # noqa: BLE001
        return 1400.0, "fallback bf16 sustained (B200_PROFILING.md)"


# ----------------------------------------------------------------------------- our arm
def run_ours(args):
    import numpy as np
    import torch
    import torch.distributed as dist

    from planar_optical_flow_b200.engine import StreamingDetector

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world != args.gpus and world > 1:
        raise SystemExit("--gpus %d but WORLD_SIZE=%d" % (args.gpus, world))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)

    K, W = args.steps, args.warmup
    if args.scaling == "strong":                   # configs[2] as written: the sequences of the job, sharded round-robin
        from planar_optical_flow_b200.parallel import shard_sequences

        mine = shard_sequences(args.sequences, rank, world)
        phi, scans = make_sequences(args.shape, args.sequences, W + K, seed0=0)
        scans = np.ascontiguousarray(scans[:, mine])
        B, total_seqs = len(mine), args.sequences
        if B == 0:
            raise SystemExit("rank %d owns no sequence: --sequences %d < --gpus %d" % (rank, args.sequences, world))
    else:
        B, total_seqs = args.sequences, world * args.sequences
        phi, scans = make_sequences(args.shape, B, W + K, seed0=1000 * rank)      # every rank owns ITS sequences
    N = len(phi)
    model = build_model()
    use_graph = args.graph or B <= 16
    if use_graph and W < 4:                        # two eager steps, then one capture per memory parity, all before the timed region
        W = 4
        phi, scans = make_sequences(args.shape, total_seqs if args.scaling == "strong" else B, W + K,
                                    seed0=0 if args.scaling == "strong" else 1000 * rank)
        if args.scaling == "strong":
            scans = np.ascontiguousarray(scans[:, mine])

    def timed_run(precision, host_path, record, graph=None):
        graph = use_graph if graph is None else graph
        det = StreamingDetector(model, phi, CUTOUT_KW, B, device=dev, precision=precision, record_events=record and not graph,
                                seq_chunk=args.seq_chunk or None, cuda_graph=graph)
        d_scans = torch.from_numpy(scans).to(dev)
        run = (lambda t: det.step(scans[t])) if host_path else (lambda t: det.step_device(d_scans[t]))
        checksum = 0
        for t in range(W):
            run(t)
        det.events = {k: [] for k in det.events}
        det.event_work = {}
        launches0 = det.kernel_launches
        torch.cuda.synchronize(dev)
        if world > 1:
            dist.barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for t in range(W, W + K):
            res = run(t)
            if host_path:
                checksum += int(res["n_keep"].sum())
        e1.record()
        torch.cuda.synchronize(dev)
        det.check()
        if world > 1:
            dist.barrier()
        ms = torch.tensor([e0.elapsed_time(e1)], device=dev)
        if world > 1:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        if not host_path:
            checksum = int(res["n_keep"].sum().item())
        return float(ms.item()), det, det.kernel_launches - launches0, checksum

    sampler = ClockSampler(local) if rank == 0 else None
    if sampler:
        sampler.start()
    try:
        ms_dev, det, launches, _ = timed_run(args.precision, host_path=False, record=True)
    except Exception as e:      # noqa: BLE001
        if not use_graph:
            raise
        # a failed stream capture must not cost the run its number: say so and issue the launches from Python instead
        sys.stderr.write("CUDA-graph replay failed (%r); falling back to launch-by-launch steps\n" % (e,))
        use_graph = False
        torch.cuda.synchronize(dev)
        torch.cuda.empty_cache()
        ms_dev, det, launches, _ = timed_run(args.precision, host_path=False, record=True)
    spot = None
    if rank == 0 and not args.no_parity_spot:
        spot = parity_spot(det, phi, scans, sorted({0, B - 1}))
    if use_graph:          # per-stage CUDA events cannot sit inside a graph: the stage times come from an eager run of the same steps
        del det
        torch.cuda.empty_cache()
        ms_eager, det, _, _ = timed_run(args.precision, host_path=False, record=True, graph=False)
    gate_ms = det.event_ms("gate")
    cut_ms = det.event_ms("cutout")
    nms_ms = det.event_ms("nms")
    conv_ms = {k: det.event_ms(k) for k in ("conv64", "conv128", "conv256")}
    conv_flops = {k: list(det.event_work.get(k, [])) for k in conv_ms}
    chunk_seqs = det.seq_chunk
    h2d, d2h = det.h2d_bytes_per_step, det.d2h_bytes_per_step
    del det
    torch.cuda.empty_cache()
    ms_e2e, det2, _, n_det = timed_run(args.precision, host_path=True, record=False)
    del det2
    torch.cuda.empty_cache()
    clocks = sampler.stop() if sampler else None
    extra = {}
    for prec in [p for p in args.extra_precisions.split(",") if p]:
        ms_x, d3, _, _ = timed_run(prec, host_path=False, record=False)
        del d3
        torch.cuda.empty_cache()
        extra[prec] = {"value": total_seqs * K / (ms_x / 1e3), "unit": "scans/s", "ms_per_step": ms_x / K}

    # BASELINE.json configs[1]: cutout-only sweep, largest batch (4096 JRDB-shaped scans, 1 GB of output per
    # launch, far larger than L2), both arithmetic policies; CUDA events around each launch
    cut_sweep = None
    if rank == 0:
        from planar_optical_flow_b200 import ops

        cb = 4096
        reps = -(-cb // B)
        big = torch.from_numpy(np.ascontiguousarray(np.tile(scans[W], (reps, 1))[:cb])).to(dev).unsqueeze(1)
        phi_d = torch.from_numpy(np.ascontiguousarray(phi)).to(dev)
        buf = torch.empty((cb, N, 1, CUTOUT_KW["num_cutout_pts"]), dtype=torch.float32, device=dev)
        cut_sweep = {}
        for name, fast in (("fast", True), ("exact", False)):
            # the sweep follows seconds of power-capped tensor-core work: let the SM clock settle at THIS kernel's power
            # level (~0.15 s of its own launches) before timing it - it is issue bound, so its time follows the clock
            for _ in range(400 if fast else 200):
                ops.cutout(big, phi_d, out=buf, fast=fast, **CUTOUT_KW)
            torch.cuda.synchronize(dev)
            # 20 back-to-back launches between ONE event pair: the average launch duration of the kernel in steady state
            # (an event pair per launch adds the 3-5 us a start event waits for the launch that follows it)
            s_ev, e_ev = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            s_ev.record()
            for _ in range(20):
                ops.cutout(big, phi_d, out=buf, fast=fast, **CUTOUT_KW)
            e_ev.record()
            torch.cuda.synchronize(dev)
            cut_sweep[name] = s_ev.elapsed_time(e_ev) / 20
        cut_sweep["batch"] = cb
        # what a kernel that ONLY writes can reach on this board: the same 1 GB buffer filled by torch's vectorised fill kernel
        # and by cudaMemsetAsync (the copy peak in MEASURED_PEAKS.json is half reads, half writes)
        fills = {}
        for name, fn in (("fill_kernel", lambda: buf.fill_(1.0)), ("memset", lambda: buf.zero_())):
            for _ in range(5):
                fn()
            torch.cuda.synchronize(dev)
            s_ev, e_ev = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            s_ev.record()
            for _ in range(20):
                fn()
            e_ev.record()
            torch.cuda.synchronize(dev)
            fills[name] = buf.numel() * 4 / (s_ev.elapsed_time(e_ev) / 20 * 1e-3) / 1e9
        cut_sweep["write_only_gbs"] = fills
        del big, buf
        torch.cuda.empty_cache()

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    peak, peak_src = measured_hbm_peak()
    # algorithmic bytes per launch: SURVEY.md §8d per-unit figure x sequences in one launch
    gate_bytes_per_seq = N * 44076
    per_launch_seqs = [min(chunk_seqs, B - b0) for b0 in range(0, B, chunk_seqs)]
    seqs_per_launch = sum(per_launch_seqs) / len(per_launch_seqs)
    gate_avg_ms = sum(gate_ms) / len(gate_ms)
    gate_gbs = gate_bytes_per_seq * seqs_per_launch / (gate_avg_ms * 1e-3) / 1e9
    cut_avg_ms = sum(cut_ms) / len(cut_ms)
    cut_gbs = N * 228 * B / (cut_avg_ms * 1e-3) / 1e9
    sweep_bytes = N * 228 * cut_sweep["batch"]
    sweep_gbs = {k: sweep_bytes / (cut_sweep[k] * 1e-3) / 1e9 for k in ("fast", "exact")}
    # the dominant kernel of the step: the tcgen05 convolution, BN = 256 instantiation (layers with >= 256 output channels)
    conv_roof = None
    if conv_ms["conv256"]:
        t_ms, fl = sum(conv_ms["conv256"]), sum(conv_flops["conv256"])
        tpeak, tsrc = measured_tensor_peak()
        all_ms = sum(sum(v) for v in conv_ms.values())
        all_fl = sum(sum(v) for v in conv_flops.values())
        ach = fl / (t_ms * 1e-3) / 1e12
        f16 = args.precision == "fp32"
        pipe_peak = tpeak if f16 else tpeak / 2.0
        conv_roof = {"kernel": "conv_tc_kernel<256,2,%s> (tcgen05 3x%s split convolution, SM pairs; layers with >= 256 output channels)"
                               % (("true", "F16") if f16 else ("false", "TF32")),
                     "bound": "tensor", "achieved": ach, "peak": pipe_peak, "unit": "TFLOP/s", "frac": ach / pipe_peak,
                     "traffic": (ncu_traffic("conv_tc_256_layer_256to512_M%d" % (chunk_seqs * N)) or {}).get("bytes"),
                     "traffic_note": "ncu dram bytes of the 256->512 layer's launch at this launch size (the largest of the BN = 256 launches; "
                                     "algorithmic: 1.00 GB of split activations in, 0.50 GB out + 0.50 GB of weights and re-reads served by L2)",
                     "peak_source": tsrc + (" (kind::f16 runs at the bf16 rate)" if f16 else " / 2 (TF32 runs at half the bf16 rate)"),
                     "avg_launch_ms": t_ms / len(conv_ms["conv256"]), "launches_timed": len(conv_ms["conv256"]),
                     "algorithmic_flops_per_launch": fl / len(conv_ms["conv256"]),
                     "note": "achieved = fp32 convolution FLOPs (2*rows*Cin*Cout*taps) / CUDA-event time; the kernel issues THREE "
                             "MMAs per algorithmic product (hi*hi, lo*hi, hi*lo) to stay fp32-accurate",
                     "mma_tflops": 3 * ach, "mma_frac_of_pipe_peak": 3 * ach / pipe_peak,
                     "share_of_step": t_ms / K / (ms_dev / K),
                     "all_conv_tc_launches": {"achieved": all_fl / (all_ms * 1e-3) / 1e12, "ms_per_step": all_ms / K,
                                              "share_of_step": all_ms / ms_dev}}
    out = {
        "metric": METRIC, "value": total_seqs * K / (ms_dev / 1e3), "unit": "scans/s", "n_gpus": world,
        "steps": K, "warmup": W, "ms_per_step": ms_dev / K, "higher_is_better": True, "scaling": args.scaling,
        "vs_baseline": None, "dtype": {"fp32": "f32 (operands split into two float16 parts, three kind::f16 products per term on tcgen05, 128-channel chains promoted to fp32 registers; 6-7e-7 per layer vs fp64, cuDNN fp32: 1-2e-6)",
                                      "fp32-tf32": "f32 (3xTF32 split products on tcgen05, 64-channel chains promoted to fp32 registers; 5-7e-7 per layer vs fp64)",
                                      "fp32-simt": "f32", "tf32x3": "f32 operands, TF32 tensor-core accumulation (1e-4)",
                                      "tf32": "tf32"}[args.precision], "data": "synthetic",
        "impl": "ours",
        "config": stream_config(args, B, total_seqs, N, world, use_graph),
        "e2e": {"value": total_seqs * K / (ms_e2e / 1e3), "unit": "scans/s", "h2d_bytes_per_step": h2d,
                "d2h_bytes_per_step": d2h, "ms_per_step": ms_e2e / K, "detections_in_timed_region": n_det,
                "api": "StreamingDetector.step(host ranges) -> host detections"},
        "gpu_launches": launches,
        "parity_spot": spot,
        "roofline": conv_roof,
        "roofline_gate": {"kernel": "gate_stream_kernel<11,0> (attention memory update)", "bound": "hbm",
                     "achieved": gate_gbs, "peak": peak, "unit": "GB/s", "frac": gate_gbs / peak,
                     "traffic": (ncu_traffic("gate_stream_split_B%d" % chunk_seqs) or {}).get("bytes"),
                     "peak_source": peak_src, "avg_launch_ms": gate_avg_ms, "launches_timed": len(gate_ms),
                     "algorithmic_bytes_per_launch": gate_bytes_per_seq * seqs_per_launch,
                     "frac_of_nominal_8TBs": gate_gbs / 8000.0,
                     "with_operand_split": {
                         "bytes_per_launch": (gate_bytes_per_seq + (N * 3584 * 4 if args.precision == "fp32" else 0)) * seqs_per_launch,
                         "achieved": (gate_bytes_per_seq + (N * 3584 * 4 if args.precision == "fp32" else 0)) * seqs_per_launch / (gate_avg_ms * 1e-3) / 1e9,
                         "frac": (gate_bytes_per_seq + (N * 3584 * 4 if args.precision == "fp32" else 0)) * seqs_per_launch / (gate_avg_ms * 1e-3) / 1e9 / peak,
                         "note": "in the engine's default precision the same launch also writes the new memory as the float16 [hi | lo] operand of the "
                                 "convolutions that consume it (N*3584*4 B per sequence-step on top of SURVEY 8d's N*44,076 B; it replaced a separate "
                                 "read-and-split pass): `achieved` / `frac` above count the SURVEY bytes only"}},
        "roofline_cutout": {"kernel": "cutout_scan_kernel (one CTA per scan, FAST numerics), cutout-only sweep at batch %d "
                                      "(BASELINE.json configs[1])" % cut_sweep["batch"],
                            "bound": "hbm", "achieved": sweep_gbs["fast"], "peak": peak, "unit": "GB/s",
                            "frac": sweep_gbs["fast"] / peak, "avg_launch_ms": cut_sweep["fast"],
                            "traffic": (ncu_traffic("cutout_scan_B%d" % cut_sweep["batch"]) or {}).get("bytes"),
                            "algorithmic_bytes_per_launch": sweep_bytes,
                            "write_only_reference": {"measured_gbs": cut_sweep["write_only_gbs"],
                                                     "frac_of_best": sweep_gbs["fast"] / max(cut_sweep["write_only_gbs"].values()),
                                                     "note": "the kernel only writes (reads are 1/57 of its traffic): bandwidth of torch's fill kernel and of "
                                                             "a memset over the same 1 GB output buffer, measured in this run, for context; `frac` above "
                                                             "stays against the read+write copy peak"},
                            "exact_arithmetic": {"kernel": "cutout_scan_exact_kernel (what scans_to_cutout and the engine run)",
                                                 "achieved": sweep_gbs["exact"], "frac": sweep_gbs["exact"] / peak,
                                                 "avg_launch_ms": cut_sweep["exact"]},
                            "in_streaming_step": {"achieved": cut_gbs, "frac": cut_gbs / peak, "avg_launch_ms": cut_avg_ms,
                                                  "note": "the engine's own cutout call (EXACT arithmetic, one launch) per step "
                                                          "over %d sequences (64 MB): launch-latency bound at this size" % B}},
        "stage_ms_per_step": {"cutout": sum(cut_ms) / K, "gate": sum(gate_ms) / K, "nms": sum(nms_ms) / K,
                              "convolutions_tcgen05": sum(sum(v) for v in conv_ms.values()) / K,
                              "rest": ms_dev / K - (sum(cut_ms) + sum(gate_ms) + sum(nms_ms) + sum(sum(v) for v in conv_ms.values())) / K},
        "clocks": clocks,
    }
    if use_graph:
        out["stage_ms_per_step"]["note"] = ("stage times: CUDA events of an eager pass over the same steps (events cannot sit inside a graph), "
                                            "%.2f ms per step; `rest` = the graph-replayed step minus those stages" % (ms_eager / K))
        out["latency"] = {"ms_per_step": ms_dev / K, "ms_per_scan_of_a_sequence": ms_dev / K, "e2e_ms_per_step": ms_e2e / K,
                          "eager_ms_per_step": ms_eager / K, "launches_per_step_in_graph": launches / K,
                          "note": "each step = one scan of each of the %d sequence(s), replayed as one CUDA graph per memory parity; "
                                  "eager = the same steps issued launch by launch from Python" % B}
    if extra:
        out["other_precisions"] = extra
    if out["roofline"] is None:          # library-convolution modes: the attention kernel is the dominant libpof kernel
        out["roofline"] = out["roofline_gate"]
    if not args.no_cpu_baseline and world == 1:
        v, cores, n, stage = cpu_reference_scans_per_s(args.shape, args.cpu_scans)
        kind = cpu_reference_kind()
        out["cpu_baseline"] = {"value": v, "unit": "scans/s", "cores": cores, "kind": kind,
                               "sample": "%d scans of one %s-shaped sequence streamed through %s "
                                         "(NumPy cutout, torch-CPU SpatialDROW with dense attention, NumPy NMS)"
                                         % (n, args.shape.upper(), "the unmodified reference (baseline/_ref)" if kind == "reference" else "the oracle port"),
                               "stage_ms_per_scan": stage}
        # the reference's PyTorch-GPU path (north_star's second bar): same loop, the network on this GPU
        vg, _, ng, stage_g = cpu_reference_scans_per_s(args.shape, 4 * args.cpu_scans, warmup=4, model_device="cuda")
        out["torch_gpu_baseline"] = {"value": vg, "unit": "scans/s", "kind": kind,
                                     "sample": "%d scans of one %s-shaped sequence: NumPy cutout on the host, the %s "
                                               "SpatialDROW (dense attention) on cuda in strict fp32, NumPy NMS; batch 1 as in "
                                               "depracted_scripts/infer_person_flow.py" % (ng, args.shape.upper(),
                                                                                         "reference's own" if kind == "reference" else "oracle's"),
                                     "stage_ms_per_scan": stage_g}
    _JSON_OUT.write(json.dumps(out) + "\n")
    _JSON_OUT.flush()
    if world > 1:
        dist.destroy_process_group()


# ----------------------------------------------------------------------------- training workload (configs[3])
TRAIN_METRIC = "DR-SPAAM training samples/sec (DROW 450-pt, 11 scans per sample)"


def run_train(args):
    """BASELINE.json configs[3]: the step of bin/train_dr_spaam.py with config/dr_spaam.yaml - SpatialDROW, per-GPU batch 8,
    11 scans x 450 points per sample, cutouts on the device, fused attention-memory forward/backward, detector loss, Adam;
    DistributedDataParallel (one 7.9 MB bucket, NCCL over NVLink) when N > 1.  `value`: batches resident on the device;
    `e2e`: every step takes a fresh host batch through the loader's pinned staging buffers (H2D inside the timed region)
    and reads the loss back."""
    import numpy as np
    import torch
    import torch.distributed as dist
    import yaml
    from torch import optim

    from planar_optical_flow_b200 import parallel
    from planar_optical_flow_b200.dataset_dr_spaam import DeviceBatches, create_dataloader
    from planar_optical_flow_b200.eval_utils import make_model_fn_obj_det
    from planar_optical_flow_b200.model import SpatialDROW

    rank, local, world = parallel.env_rank_world()
    if world != args.gpus and world > 1:
        raise SystemExit("--gpus %d but WORLD_SIZE=%d" % (args.gpus, world))
    dev = torch.device("cuda", local)
    torch.cuda.set_device(dev)
    parallel.init(device=dev)
    with open(os.path.join(ROOT, "config", "dr_spaam.yaml")) as f:
        cfg = yaml.safe_load(f)
    bs, K, W = cfg["batch_size"], args.steps, max(args.warmup, 3)
    n_batches = 4
    loader, _ = create_dataloader(data_path=os.path.join(ROOT, "no-such-dir"), num_scans=cfg["num_scans"], batch_size=bs, num_workers=0,
                                  network_type=cfg["network"], use_data_augumentation=cfg["use_data_augumentation"],
                                  cutout_kwargs=cfg["cutout_kwargs"], polar_grid_kwargs=cfg["polar_grid_kwargs"],
                                  pedestrian_only=cfg["pedestrian_only"], num_samples=bs * n_batches * max(world, 1))
    host_batches = [b for b in loader][:n_batches]

    class _Replay:                                         # the loader's batches, again and again (synthesis is not the measured path)
        sampler = dataset = None

        def __len__(self):
            return 1 << 30

        def __iter__(self):
            while True:
                yield from host_batches

    staged = DeviceBatches(_Replay(), dev)
    dev_batches = []
    it = iter(staged)
    for _ in range(n_batches):
        b = next(it)
        dev_batches.append({k: (v.clone() if isinstance(v, torch.Tensor) else v) for k, v in b.items()})
    torch.manual_seed(0)
    net = SpatialDROW(num_scans=cfg["num_scans"], num_pts=cfg["cutout_kwargs"]["num_cutout_pts"],
                      focal_loss_gamma=cfg["focal_loss_gamma"], alpha=cfg["similarity_kwargs"]["alpha"],
                      window_size=cfg["similarity_kwargs"]["window_size"], pedestrian_only=cfg["pedestrian_only"]).to(dev)
    opt = optim.Adam(net.parameters(), lr=0.01)
    n_params = sum(p.numel() for p in net.parameters())
    net = parallel.wrap_ddp(net, dev)
    fn = make_model_fn_obj_det(cfg["cutout_kwargs"])
    net.train()

    def one(batch):
        opt.zero_grad(set_to_none=True)
        loss = fn(net, batch)[0]
        loss.backward()
        opt.step()
        return loss

    def timed(source, read_loss):
        for i in range(W):
            one(source(i))
        torch.cuda.synchronize(dev)
        if world > 1:
            dist.barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        last = 0.0
        for i in range(K):
            loss = one(source(W + i))
            if read_loss:
                last = float(loss.item())                 # the D2H read of the step's result
        e1.record()
        torch.cuda.synchronize(dev)
        if world > 1:
            dist.barrier()
        ms = torch.tensor([e0.elapsed_time(e1)], device=dev)
        if world > 1:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        return float(ms.item()), last if read_loss else float(loss.item())

    sampler = ClockSampler(local) if rank == 0 else None
    if sampler:
        sampler.start()
    ms_dev, loss_dev = timed(lambda i: dev_batches[i % n_batches], read_loss=False)
    staged.h2d_bytes = 0
    ms_e2e, loss_e2e = timed(lambda i: next(it), read_loss=True)
    h2d = staged.h2d_bytes // (W + K)
    clocks = sampler.stop() if sampler else None

    # the gate's backward in isolation at the training shape (HBM bound): g_out and the memory are read once, g_x and g_tmpl written
    gate_roof = None
    if rank == 0:
        from planar_optical_flow_b200 import ops

        Bn, N, CL, E, Wn = bs, 450, 3584, 128, int(2 * int(cfg["similarity_kwargs"]["window_size"] / 2) + 1)
        g = torch.Generator(device=dev).manual_seed(0)
        tmpl = torch.randn(Bn, N, CL, device=dev, generator=g)
        ex, et = torch.randn(Bn, N, E, device=dev, generator=g) * 0.2, torch.randn(Bn, N, E, device=dev, generator=g) * 0.2
        _, _, attn = ops.gate_forward(tmpl, tmpl.clone(), ex, et, 0.5, Wn, want_weights=True)
        g_out, g_feat = torch.randn(Bn, N, CL, device=dev, generator=g), torch.randn(Bn, N, Wn, device=dev, generator=g)
        for _ in range(5):
            ops.gate_backward(tmpl, ex, et, attn, g_out, g_feat, 0.5, Wn)
        torch.cuda.synchronize(dev)
        s_ev, e_ev = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s_ev.record()
        for _ in range(20):
            ops.gate_backward(tmpl, ex, et, attn, g_out, g_feat, 0.5, Wn)
        e_ev.record()
        torch.cuda.synchronize(dev)
        bwd_ms = s_ev.elapsed_time(e_ev) / 20
        bwd_bytes = Bn * N * (4 * CL * 4 + 4 * E * 4 + 3 * Wn * 4)        # read g_out, tmpl; write g_x, g_tmpl; embeddings and their gradients; weights, g_feat, g_s
        peak, peak_src = measured_hbm_peak()
        gbs = bwd_bytes / (bwd_ms * 1e-3) / 1e9
        gate_roof = {"kernel": "pof_spaam_gate_bwd (gate_bwd_scores + gate_bwd_embed + gate_stream_kernel<11,1>), per-GPU training batch "
                               "(%d x %d points), isolated launches" % (Bn, N),
                     "bound": "hbm", "achieved": gbs, "peak": peak, "unit": "GB/s", "frac": gbs / peak, "traffic": None,
                     "avg_launch_ms": bwd_ms, "algorithmic_bytes_per_launch": bwd_bytes, "peak_source": peak_src,
                     "note": "small launch (%.0f MB): latency matters as much as bandwidth; 10 such calls per training step" % (bwd_bytes / 1e6)}
    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return
    out = {"metric": TRAIN_METRIC, "value": world * bs * K / (ms_dev / 1e3), "unit": "samples/s", "n_gpus": world, "steps": K,
           "warmup": W, "ms_per_step": ms_dev / K, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
           "dtype": "f32 parameters and activations; cuDNN convolutions with TF32 tensor-core products (PyTorch's training default, "
                    "as the reference trains); attention memory forward/backward in fp32 (libpof)",
           "data": "synthetic", "impl": "ours",
           "config": {"workload": "DR-SPAAM training step (bin/train_dr_spaam.py, config/dr_spaam.yaml): SpatialDROW, per-GPU batch %d x %d scans x "
                                  "450 points, device cutouts, fused gate fwd/bwd, Adam%s" % (bs, cfg["num_scans"] + 1,
                                                                                            ", DDP all-reduce over NCCL" if world > 1 else ""),
                      "per_gpu_batch": bs, "scans_per_sample": cfg["num_scans"] + 1, "points": 450, "parameters": n_params,
                      "allreduce_bytes_per_step": 4 * n_params if world > 1 else 0,
                      "l2_policy": "activations of a step (several GB) are far larger than L2"},
           "e2e": {"value": world * bs * K / (ms_e2e / 1e3), "unit": "samples/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": 4,
                   "ms_per_step": ms_e2e / K, "api": "model_fn_obj_det(model, host batch through DeviceBatches) + backward + Adam, loss.item()"},
           "gpu_launches": K * (2 + cfg["num_scans"] + 3 * cfg["num_scans"]),
           "gpu_launches_note": "libpof launches per step: the cutout call (2 kernels) + %d gate forwards + %d gate backwards of 3 kernels; "
                                "the convolutions are cuDNN" % (cfg["num_scans"], cfg["num_scans"]),
           "roofline": gate_roof, "last_loss": {"device_batches": loss_dev, "e2e": loss_e2e}, "clocks": clocks}
    _JSON_OUT.write(json.dumps(out) + "\n")
    _JSON_OUT.flush()
    if world > 1:
        dist.destroy_process_group()


# ----------------------------------------------------------------------------- BASELINE configs[1]: the cutout-only sweep
CUTOUT_SWEEP_BATCHES = (1, 4, 16, 64, 256, 1024, 4096)


def cutout_sweep_scans(shape, batch):
    """[batch, N] structured ranges (16 seeded walks tiled with per-scan jitter) and phi."""
    import numpy as np

    phi, seq = make_sequences(shape, min(batch, 256), 1, seed0=4000)
    scans = seq[0]
    if batch > scans.shape[0]:
        rs = np.random.RandomState(7)
        scans = np.tile(scans, (-(-batch // scans.shape[0]), 1))[:batch]
        scans = np.clip(scans + rs.normal(0.0, 0.02, size=scans.shape).astype(np.float32), 0.05, 29.99)
    return phi, np.ascontiguousarray(scans, dtype=np.float32)


def cpu_reference_cutout_rows_per_s(shape, n_scans):
    """The reference's NumPy `scans_to_cutout` (baseline/_ref, else the oracle port) looped over single scans on ONE host
    thread, as SURVEY.md 8d prescribes for this sweep (the function is single-threaded NumPy)."""
    from oracle import cutout as ocut

    ref = reference_modules()
    fn = ref[0].scans_to_cutout if ref is not None else ocut.scans_to_cutout
    phi, scans = cutout_sweep_scans(shape, n_scans + 2)
    for b in range(2):
        fn(scans[b:b + 1], phi, stride=1, **CUTOUT_KW)
    t0 = time.perf_counter()
    for b in range(2, n_scans + 2):
        fn(scans[b:b + 1], phi, stride=1, **CUTOUT_KW)
    return n_scans / (time.perf_counter() - t0)


def run_cutout(args):
    """`scans_to_cutout`-only sweep on JRDB-shaped scans, batch 1-4096 on one GPU (BASELINE.json configs[1]).  A step = one
    batched call over B scans.  `value` = scan rows/s at B = 4096 with EXACT arithmetic (what `scans_to_cutout` computes), ranges
    resident in HBM; `e2e` = the same call from host ranges to host cutouts (the 1 GB result crosses PCIe)."""
    import numpy as np
    import torch

    from planar_optical_flow_b200 import ops

    if args.impl == "reference":
        v = cpu_reference_cutout_rows_per_s(args.shape, max(args.steps, 1) * 32)
        kind = cpu_reference_kind()
        _JSON_OUT.write(json.dumps({
            "impl": "reference", "metric": "scans_to_cutout scan rows/s", "value": v, "unit": "scans/s", "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": 32e3 / v, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": "scans_to_cutout-only sweep, %s-shaped scans, batch 1-4096" % args.shape.upper()},
            "cpu_baseline": {"value": v, "unit": "scans/s", "cores": 1, "kind": kind,
                             "sample": "each step = 32 single-scan calls of the reference's NumPy scans_to_cutout on one host thread"},
            "e2e": {"value": v, "unit": "scans/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}) + "\n")
        _JSON_OUT.flush()
        return
    dev = torch.device("cuda", int(os.environ.get("LOCAL_RANK", "0")))
    torch.cuda.set_device(dev)
    K, W = max(args.steps, 1), max(args.warmup, 3)
    peak, peak_src = measured_hbm_peak()
    P = CUTOUT_KW["num_cutout_pts"]
    sampler = ClockSampler(dev.index or 0)
    sampler.start()
    sweep, e2e = [], None
    for B in CUTOUT_SWEEP_BATCHES:
        phi, scans = cutout_sweep_scans(args.shape, B)
        N = scans.shape[1]
        phi_d = torch.from_numpy(np.ascontiguousarray(phi)).to(dev)
        s_d = torch.from_numpy(scans).to(dev).unsqueeze(1)
        out = torch.empty((B, N, 1, P), dtype=torch.float32, device=dev)
        row = {"batch": B}
        for name, fast in (("exact", False), ("fast", True)):
            for _ in range(W + (200 if B >= 1024 else 20)):          # warm-up long enough for the clock to settle at this kernel's power
                ops.cutout(s_d, phi_d, out=out, fast=fast, **CUTOUT_KW)
            torch.cuda.synchronize(dev)
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(K):
                ops.cutout(s_d, phi_d, out=out, fast=fast, **CUTOUT_KW)
            e1.record()
            torch.cuda.synchronize(dev)
            ms = e0.elapsed_time(e1) / K
            row[name] = {"ms_per_call": ms, "scan_rows_per_s": B / (ms * 1e-3), "gbs": B * N * 228 / (ms * 1e-3) / 1e9,
                         "frac": B * N * 228 / (ms * 1e-3) / 1e9 / peak}
        sweep.append(row)
        if B == CUTOUT_SWEEP_BATCHES[-1]:
            # end to end: pinned host ranges -> device -> kernel -> pinned host cutouts, every step
            h_in = torch.from_numpy(scans).pin_memory()
            h_out = torch.empty((B, N, 1, P), dtype=torch.float32).pin_memory()
            for _ in range(W):
                s_d.copy_(h_in.unsqueeze(1), non_blocking=True)
                ops.cutout(s_d, phi_d, out=out, **CUTOUT_KW)
                h_out.copy_(out, non_blocking=True)
            torch.cuda.synchronize(dev)
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(K):
                s_d.copy_(h_in.unsqueeze(1), non_blocking=True)
                ops.cutout(s_d, phi_d, out=out, **CUTOUT_KW)
                h_out.copy_(out, non_blocking=True)
            e1.record()
            torch.cuda.synchronize(dev)
            ms = e0.elapsed_time(e1) / K
            e2e = {"value": B / (ms * 1e-3), "unit": "scans/s", "h2d_bytes_per_step": int(h_in.numel() * 4),
                   "d2h_bytes_per_step": int(h_out.numel() * 4), "ms_per_step": ms,
                   "api": "ops.cutout on pinned host ranges, result copied back to pinned host memory (PCIe bound: 1 GB per call)"}
            # parity spot: three of the timed scans against the oracle (reference half-angles fed in: every other operation is exact)
            from oracle import cutout as ocut

            worst, exact = 0.0, 1.0
            for b in (0, B // 2, B - 1):
                ha = ocut.window_half_angle(scans[b:b + 1], 1, CUTOUT_KW["fixed"], CUTOUT_KW["window_width"])
                want = ocut.scans_to_cutout(scans[b:b + 1], phi, **CUTOUT_KW)
                got = ops.cutout(s_d[b:b + 1], phi_d, half_alpha=torch.from_numpy(np.ascontiguousarray(ha, np.float32)).to(dev).unsqueeze(0),
                                 **CUTOUT_KW)[0].cpu().numpy()
                worst = max(worst, float(np.abs(got.astype(np.float64) - want).max() / max(np.abs(want).max(), 1e-30)))
                exact = min(exact, float((got == want).mean()))
    clocks = sampler.stop()
    last = sweep[-1]
    out_line = {
        "metric": "scans_to_cutout scan rows/s (JRDB 1091-pt, batch 4096)", "value": last["exact"]["scan_rows_per_s"], "unit": "scans/s",
        "n_gpus": 1, "steps": K, "warmup": W, "ms_per_step": last["exact"]["ms_per_call"], "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": "scans_to_cutout-only sweep, %s-shaped scans (%d pts), batch 1-4096, EXACT arithmetic (the API default); "
                               "FAST beside it" % (args.shape.upper(), N), "cutout_pts": P,
                   "l2_policy": "1 GB of output per call at the quoted batch (L2 = 126 MB)"},
        "e2e": e2e, "gpu_launches": K * 2 * len(CUTOUT_SWEEP_BATCHES), "sweep": sweep,
        "parity_spot": {"max_rel_with_reference_half_angles": worst, "bit_equal_fraction": exact},
        "roofline": {"kernel": "cutout_scan_exact_kernel", "bound": "hbm", "achieved": last["exact"]["gbs"], "peak": peak, "unit": "GB/s",
                     "frac": last["exact"]["frac"], "traffic": None, "peak_source": peak_src,
                     "algorithmic_bytes_per_launch": CUTOUT_SWEEP_BATCHES[-1] * N * 228,
                     "fast_arithmetic": {"kernel": "cutout_scan_kernel", "achieved": last["fast"]["gbs"], "frac": last["fast"]["frac"]}},
        "clocks": clocks}
    if not args.no_cpu_baseline:
        v = cpu_reference_cutout_rows_per_s(args.shape, 256)
        out_line["cpu_baseline"] = {"value": v, "unit": "scans/s", "cores": 1, "kind": cpu_reference_kind(),
                                    "ideal_all_cores": v * (os.cpu_count() or 1), "host_cores": os.cpu_count(),
                                    "sample": "256 single-scan calls of the reference's NumPy scans_to_cutout on one host thread "
                                              "(the function is single-threaded; x cores = the ideal multi-process bound)"}
    _JSON_OUT.write(json.dumps(out_line) + "\n")
    _JSON_OUT.flush()


# ----------------------------------------------------------------------------- reference arm
def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    from planar_optical_flow_b200 import synth

    # the workload is the other arm's (same `config`); what is TIMED is a bounded sample of it: one step = one scan of 4 sequences
    world_ref = max(args.gpus, 1)
    B_ref = len(range(0, args.sequences, world_ref)) if args.scaling == "strong" else args.sequences
    total_ref = args.sequences if args.scaling == "strong" else world_ref * args.sequences
    n_pts = len(synth.phi_for(args.shape))
    seqs_per_step = 4
    n = args.steps * seqs_per_step
    v, cores, timed, stage = cpu_reference_scans_per_s(args.shape, n, warmup=args.warmup, sequences=seqs_per_step)
    kind = cpu_reference_kind()
    sample = ("each step = one scan of %d of the %d %s-shaped sequences, streamed through %s"
              % (seqs_per_step, args.sequences, args.shape.upper(),
                 "the unmodified reference's CPU path (baseline/_ref: its scans_to_cutout, SpatialDROW module and nms_predicted_center)"
                 if kind == "reference" else "the oracle port of the reference's CPU path"))
    _JSON_OUT.write(json.dumps({
        "impl": "reference", "metric": METRIC, "value": v, "unit": "scans/s", "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": 1e3 * seqs_per_step / v, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": stream_config(args, B_ref, total_ref, n_pts, world_ref, B_ref <= 16 or args.graph),
        "cpu_baseline": {"value": v, "unit": "scans/s", "cores": cores, "kind": kind, "sample": sample,
                         "stage_ms_per_scan": stage},
        "e2e": {"value": v, "unit": "scans/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }) + "\n")
    _JSON_OUT.flush()


def _claim_stdout():
    """Keep stdout for the ONE JSON line: libraries (NCCL's version banner, cuDNN notes) get stderr."""
    real = os.fdopen(os.dup(1), "w")
    sys.stdout.flush()
    os.dup2(2, 1)
    return real


if __name__ == "__main__":
    a = parse()
    _JSON_OUT = _claim_stdout()
    if a.workload == "cutout":
        run_cutout(a)
    elif a.impl == "reference":
        run_reference(a)
    elif a.workload == "train":
        run_train(a)
    else:
        run_ours(a)
