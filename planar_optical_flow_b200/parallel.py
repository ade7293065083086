"""Multi-GPU plumbing: one process per GPU, independent scan sequences sharded across ranks.

The hot path shards by sequence: a sequence's cutouts, attention memory and NMS
never touch another sequence, so the data path has NO collective.  torch.distributed (NCCL on the
GPU box, gloo in the CPU tests) is used only for
  * metric gathers at the end of a run (scan counts, detection counts, max-over-ranks timings), and
  * the gradient all-reduce of the training step (DistributedDataParallel, one 7.9 MB bucket).
"""
import os

import torch
import torch.distributed as dist


def env_rank_world():
    """(rank, local_rank, world_size) from the torchrun environment; (0, 0, 1) without it."""
    return (int(os.environ.get("RANK", "0")), int(os.environ.get("LOCAL_RANK", "0")),
            int(os.environ.get("WORLD_SIZE", "1")))


def init(backend=None, device=None):
    """Initialise the default process group if the job is multi-process.  Returns (rank, local_rank, world)."""
    rank, local, world = env_rank_world()
    if world > 1 and not dist.is_initialized():
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        os.environ.setdefault("MASTER_PORT", "29511")
        if backend is None:
            backend = "nccl" if torch.cuda.is_available() else "gloo"
        kwargs = {}
        if backend == "nccl" and device is not None:
            kwargs["device_id"] = device
        dist.init_process_group(backend, rank=rank, world_size=world, **kwargs)
    return rank, local, world


def shard_sequences(n_total, rank, world):
    """Indices of the sequences rank `rank` owns: round-robin, b -> rank b % world.

    Every sequence is owned by exactly one rank; the sizes differ by at most one.
    """
    if not (0 <= rank < world):
        raise ValueError("rank %d outside world of %d" % (rank, world))
    return list(range(rank, n_total, world))


def gather_metrics(local, device=None, reduce_max=("elapsed_ms",)):
    """Combine per-rank counters: keys in `reduce_max` take the max over ranks (timings are never
    wall-clock averaged), everything else is summed.  Works on NCCL (device tensors) and gloo."""
    keys = sorted(local)
    if not dist.is_initialized() or dist.get_world_size() == 1:
        return {k: float(local[k]) for k in keys}
    dev = device if device is not None else (torch.device("cuda", torch.cuda.current_device())
                                             if dist.get_backend() == "nccl" else torch.device("cpu"))
    sums = torch.tensor([float(local[k]) for k in keys if k not in reduce_max], dtype=torch.float64, device=dev)
    maxs = torch.tensor([float(local[k]) for k in keys if k in reduce_max], dtype=torch.float64, device=dev)
    if sums.numel():
        dist.all_reduce(sums, op=dist.ReduceOp.SUM)
    if maxs.numel():
        dist.all_reduce(maxs, op=dist.ReduceOp.MAX)
    out, si, mi = {}, 0, 0
    for k in keys:
        if k in reduce_max:
            out[k] = float(maxs[mi]); mi += 1
        else:
            out[k] = float(sums[si]); si += 1
    return out


def wrap_ddp(model, device=None):
    """DistributedDataParallel with a single bucket: DR-SPAAM has 1,977,667 fp32 parameters (7.9 MB);
    over NVLink 5 the all-reduce is latency bound, so one bucket beats several.  BatchNorm stays
    per-rank, as in the (single-GPU) reference."""
    if not dist.is_initialized() or dist.get_world_size() == 1:
        return model
    ids = [device.index] if (device is not None and device.type == "cuda") else None
    return torch.nn.parallel.DistributedDataParallel(model, device_ids=ids, bucket_cap_mb=16,
                                                     broadcast_buffers=False, gradient_as_bucket_view=True)
