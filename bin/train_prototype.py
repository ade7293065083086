#!/usr/bin/env python
"""Scan-pair flow prototype training:  python bin/train_prototype.py [--epochs 300] [--batch-size 100] [--ckpt x.pth]

The reference's bin/train_prototype.py (:20-95) hard-codes its configuration (300 epochs, batch 100,
Adam lr 0.01, Prototype(in_channel=2)) and cannot run at HEAD.  This is the
same run made coherent, with those values as defaults.  Forward + backward go through the windowed
patch-correlation kernels (csrc/pof_corr.cu).  Under torchrun it is data parallel: one process per GPU,
DistributedDataParallel gradient all-reduce over NCCL, rank 0 writes checkpoints.
"""
import argparse
import os
import sys

import torch
from torch import optim
from torch.utils.data.distributed import DistributedSampler

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

import src.utils.train_utils as tu  # noqa: E402
from planar_optical_flow_b200 import parallel  # noqa: E402
from src.depracted.model import Prototype  # noqa: E402
from src.utils.dataset import FlowDataset  # noqa: E402
from src.utils.eval_utils import model_fn, model_fn_eval  # noqa: E402
from src.utils.train_utils import Trainer, create_tb_logger, load_checkpoint  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--data", default="./../data/DROWv2-data")
    ap.add_argument("--out", default=os.path.join("./..", "output"))
    ap.add_argument("--ckpt", default=None)
    ap.add_argument("--epochs", type=int, default=300)
    ap.add_argument("--batch-size", type=int, default=100, help="per GPU")
    ap.add_argument("--ckpt-save-interval", type=int, default=100)
    ap.add_argument("--num-samples", type=int, default=1000, help="synthetic pairs per epoch")
    ap.add_argument("--max-iters", type=int, default=None)
    args = ap.parse_args()

    rank, local, world = parallel.env_rank_world()
    device = torch.device("cuda", local)
    torch.cuda.set_device(device)
    parallel.init(device=device)
    ckpt_dir = os.path.join(args.out, "ckpts")
    if rank == 0:
        os.makedirs(ckpt_dir, exist_ok=True)

    print("Prepare data")
    train = FlowDataset(args.data, split="train", num_samples=args.num_samples)
    sampler = DistributedSampler(train, num_replicas=world, rank=rank, shuffle=True) if world > 1 else None
    train_loader = torch.utils.data.DataLoader(train, batch_size=args.batch_size, shuffle=sampler is None, sampler=sampler,
                                               num_workers=0, collate_fn=train.collate_batch, drop_last=True)
    print("Prepare model")
    model = Prototype(in_channel=2).to(device)
    optimizer = optim.Adam(model.parameters(), lr=tu.lr_scheduler())
    it0, ep0 = 0, 0
    if args.ckpt is not None:
        it0, ep0 = load_checkpoint(model=model, optimizer=optimizer, filename=args.ckpt)
    elif os.path.isfile(os.path.join(ckpt_dir, "sigterm_ckpt.pth")):           # bin/train_prototype.py:68-69
        it0, ep0 = load_checkpoint(model=model, optimizer=optimizer, filename=os.path.join(ckpt_dir, "sigterm_ckpt.pth"))
    model = parallel.wrap_ddp(model, device)

    print("Start training")
    trainer = Trainer(model, model_fn, optimizer, ckpt_dir, tu.ConstantLR(optimizer), model_fn_eval=model_fn_eval,
                      grad_norm_clip=0.0, tb_logger=create_tb_logger(args.out) if rank == 0 else None, is_main=rank == 0)
    last = trainer.train(num_epochs=args.epochs, train_loader=train_loader, ckpt_save_interval=args.ckpt_save_interval,
                         starting_iteration=int(it0), starting_epoch=max(int(ep0), 0), max_iters=args.max_iters)
    if rank == 0:
        print("Analysis finished (last loss %.6f)" % (last if last is not None else float("nan")))


if __name__ == "__main__":
    main()
