"""Training-step throughput of BASELINE.json configs[3] and configs[4] on synthetic batches.  Needs a B200.

    python tools/train_step_bench.py [--steps 20] [--warmup 5]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 tools/train_step_bench.py

configs[3]: bin/train_dr_spaam.py's step - SpatialDROW (config/dr_spaam.yaml: per-GPU batch 8, 11 scans of 450 points,
            cutouts on the device, detector loss, Adam), data parallel with the DDP gradient all-reduce over NCCL.
configs[4]: bin/train_prototype.py's step - Prototype(in_channel=2) on [100, 450, 2] scan pairs, forward + backward + Adam.
One collated batch per rank is reused for every step (the loaders are host I/O, not the measured path); time is CUDA
events around the K steps, max over ranks; rank 0 prints one JSON line per config.
"""
import argparse
import json
import os
import sys

import torch
import torch.distributed as dist
import yaml
from torch import optim

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from planar_optical_flow_b200 import parallel                                   # noqa: E402
from planar_optical_flow_b200.dataset import FlowDataset                         # noqa: E402
from planar_optical_flow_b200.dataset_dr_spaam import create_dataloader         # noqa: E402
from planar_optical_flow_b200.eval_utils import make_model_fn_obj_det, model_fn  # noqa: E402
from planar_optical_flow_b200.model import SpatialDROW                           # noqa: E402
from planar_optical_flow_b200.model.prototype import Prototype                   # noqa: E402


def timed_steps(model, fn, batch, optimizer, steps, warmup, device, world):
    def one():
        optimizer.zero_grad(set_to_none=True)
        out = fn(model, batch)
        loss = out[0] if isinstance(out, tuple) else out
        loss.backward()
        optimizer.step()
        return loss

    model.train()
    for _ in range(warmup):
        one()
    torch.cuda.synchronize(device)
    if world > 1:
        dist.barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps):
        loss = one()
    e1.record()
    torch.cuda.synchronize(device)
    ms = torch.tensor([e0.elapsed_time(e1)], device=device)
    if world > 1:
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    return float(ms.item()) / steps, float(loss.item())


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    args = ap.parse_args()
    rank, local, world = parallel.env_rank_world()
    device = torch.device("cuda", local)
    torch.cuda.set_device(device)
    parallel.init(device=device)

    with open(os.path.join(ROOT, "config", "dr_spaam.yaml")) as f:
        cfg = yaml.safe_load(f)
    loader, _ = create_dataloader(data_path="./no-such-dir", num_scans=cfg["num_scans"], batch_size=cfg["batch_size"], num_workers=0,
                                  network_type=cfg["network"], train_with_val=False,
                                  use_data_augumentation=cfg["use_data_augumentation"], cutout_kwargs=cfg["cutout_kwargs"],
                                  polar_grid_kwargs=cfg["polar_grid_kwargs"], pedestrian_only=cfg["pedestrian_only"],
                                  num_samples=cfg["batch_size"] * max(world, 1))
    batch = next(iter(loader))
    torch.manual_seed(0)
    net = SpatialDROW(num_scans=cfg["num_scans"], num_pts=cfg["cutout_kwargs"]["num_cutout_pts"],
                      focal_loss_gamma=cfg["focal_loss_gamma"], alpha=cfg["similarity_kwargs"]["alpha"],
                      window_size=cfg["similarity_kwargs"]["window_size"], pedestrian_only=cfg["pedestrian_only"]).to(device)
    opt = optim.Adam(net.parameters(), lr=0.01)
    net = parallel.wrap_ddp(net, device)
    ms, loss = timed_steps(net, make_model_fn_obj_det(cfg["cutout_kwargs"]), batch, opt, args.steps, args.warmup, device, world)
    if rank == 0:
        bs = cfg["batch_size"]
        print(json.dumps({"config": "configs[3] DR-SPAAM training step (bin/train_dr_spaam.py, config/dr_spaam.yaml)",
                          "n_gpus": world, "per_gpu_batch": bs, "scans_per_sample": cfg["num_scans"] + 1, "points": 450,
                          "ms_per_step": ms, "samples_per_s": world * bs / (ms / 1e3), "last_loss": loss,
                          "parallelism": "DDP gradient all-reduce over NCCL" if world > 1 else "single GPU"}), flush=True)
    del net, opt

    ds = FlowDataset(None, split="train", num_samples=100)
    pairs = ds.collate_batch([ds[i] for i in range(100)])
    torch.manual_seed(0)
    proto = Prototype(in_channel=2).to(device)
    opt = optim.Adam(proto.parameters(), lr=0.01)
    proto = parallel.wrap_ddp(proto, device)
    ms, loss = timed_steps(proto, model_fn, pairs, opt, args.steps, args.warmup, device, world)
    if rank == 0:
        print(json.dumps({"config": "configs[4] scan-pair flow prototype step (bin/train_prototype.py)", "n_gpus": world,
                          "per_gpu_batch": 100, "points": 450, "ms_per_step": ms, "pairs_per_s": world * 100 / (ms / 1e3),
                          "last_loss": loss, "parallelism": "DDP gradient all-reduce over NCCL" if world > 1 else "single GPU"}),
              flush=True)
    if dist.is_available() and dist.is_initialized():
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
