"""World-size-2 gloo tests of the multi-process host logic (sharding, metric gathers, DDP wiring).
The data path itself has no collective; the GPU kernels are not involved here."""
import os
import socket
import sys

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, fn_name, q):
    sys.path.insert(0, ROOT)
    os.environ.update(RANK=str(rank), LOCAL_RANK=str(rank), WORLD_SIZE=str(world),
                      MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    from planar_optical_flow_b200 import parallel

    parallel.init(backend="gloo")
    try:
        q.put((rank, globals()[fn_name](rank, world, parallel)))
    finally:
        dist.destroy_process_group()


def _run(fn_name, world=2):
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, fn_name, q)) for r in range(world)]
    for p in procs:
        p.start()
    out = dict(q.get(timeout=120) for _ in range(world))
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    return out


def _metrics(rank, world, parallel):
    mine = parallel.shard_sequences(257, rank, world)
    local = {"scans": len(mine) * 10, "detections": 100 + rank, "elapsed_ms": 5.0 + rank}
    return parallel.gather_metrics(local), len(mine)


def _ddp(rank, world, parallel):
    torch.manual_seed(0)                      # same init on every rank
    model = torch.nn.Sequential(torch.nn.Conv1d(1, 4, 3, padding=1), torch.nn.BatchNorm1d(4), torch.nn.Conv1d(4, 1, 1))
    ddp = parallel.wrap_ddp(model)
    torch.manual_seed(100 + rank)             # different data per rank
    x = torch.randn(8, 1, 16)
    ddp(x).square().mean().backward()
    flat = torch.cat([p.grad.flatten() for p in model.parameters()])
    gathered = [torch.empty_like(flat) for _ in range(world)]
    dist.all_gather(gathered, flat)
    return bool(all(torch.equal(g, gathered[0]) for g in gathered)), float(flat.abs().sum())


def test_shard_sequences_partition():
    sys.path.insert(0, ROOT)
    from planar_optical_flow_b200 import parallel

    for total, world in ((256, 8), (257, 2), (3, 4), (0, 2)):
        shards = [parallel.shard_sequences(total, r, world) for r in range(world)]
        assert sorted(i for s in shards for i in s) == list(range(total))
        assert max(len(s) for s in shards) - min(len(s) for s in shards) <= 1
    with pytest.raises(ValueError):
        parallel.shard_sequences(4, 2, 2)


def test_metric_gather_sum_and_max_world2():
    out = _run("_metrics")
    for rank in (0, 1):
        merged, n_mine = out[rank]
        assert merged["scans"] == 2570            # 129 + 128 sequences x 10 scans
        assert merged["detections"] == 201
        assert merged["elapsed_ms"] == 6.0        # max over ranks, never the mean
    assert {out[0][1], out[1][1]} == {129, 128}


def test_ddp_gradients_identical_across_ranks_world2():
    out = _run("_ddp")
    assert out[0][0] and out[1][0]
    assert out[0][1] == out[1][1] > 0


def _sharded_loader(rank, world, parallel):
    """Under an initialised process group `create_dataloader` shards every epoch: the ranks' samples are disjoint, cover the
    set, and the order changes with the epoch (Trainer.train calls sampler.set_epoch)."""
    from planar_optical_flow_b200.dataset_dr_spaam import create_dataloader

    loader, _ = create_dataloader("/no-such-dir", num_scans=2, batch_size=4, num_workers=0, num_samples=32,
                                  cutout_kwargs=dict(num_cutout_pts=56))
    seen = []
    for epoch in (0, 1):
        loader.sampler.set_epoch(epoch)
        seen.append([i for b in loader for i in b["idx"]])
    return seen, type(loader.sampler).__name__


def test_training_loader_is_sharded_across_ranks_world2():
    out = _run("_sharded_loader")
    (a0, a1), name = out[0]
    (b0, b1), _ = out[1]
    assert name == "DistributedSampler"
    assert len(a0) == len(b0) == 16 and not set(a0) & set(b0) and sorted(a0 + b0) == list(range(32))
    assert not set(a1) & set(b1) and a0 != a1                     # re-shuffled, still disjoint
