#!/usr/bin/env python
"""DR-SPAAM training entry point:  python bin/train_dr_spaam.py --cfg config/dr_spaam.yaml [--ckpt x.pth]

Same CLI and YAML schema as the reference's bin/train_dr_spaam.py (:22-35).  What it runs is the
coherent version of that script: SpatialDROW + the detector loss
(`model_fn_obj_det`) + Adam(lr=0.01) + the Trainer, with cutouts generated on the GPU.  Under
torchrun it is data parallel: one process per GPU, the samples of an epoch sharded by a DistributedSampler,
DistributedDataParallel gradient all-reduce over NCCL, per-rank BatchNorm, rank 0 writes checkpoints.  `--data` is
a DROWv2 directory (train/*.csv + .wc/.wa/.wp/.odom2); if it does not exist, seeded synthetic sequences are served.  Extra flags (--max-iters, --out) only bound
the run; they do not change the training step.
"""
import argparse
import os
import sys
from shutil import copyfile

import torch
import yaml
from torch import optim

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

import src.utils.train_utils as tu  # noqa: E402
from planar_optical_flow_b200 import parallel  # noqa: E402
from src.depracted.model import SpatialDROW  # noqa: E402
from src.utils.dataset_dr_spaam import create_dataloader  # noqa: E402
from src.utils.eval_utils import make_model_fn_obj_det  # noqa: E402
from src.utils.train_utils import Trainer, create_tb_logger, load_checkpoint  # noqa: E402


def main():
    parser = argparse.ArgumentParser(description="arg parser")
    parser.add_argument("--cfg", type=str, required=True, help="configuration of the experiment")
    parser.add_argument("--ckpt", type=str, required=False, default=None)
    parser.add_argument("--data", type=str, default="./../data/DROWv2-data")
    parser.add_argument("--out", type=str, default=os.path.join("./..", "output"))
    parser.add_argument("--max-iters", type=int, default=None, help="stop after this many iterations")
    parser.add_argument("--num-samples", type=int, default=256, help="synthetic samples per epoch")
    args = parser.parse_args()

    with open(args.cfg, "r") as f:
        cfg = yaml.safe_load(f)
    cfg["name"] = os.path.basename(args.cfg).split(".")[0] + cfg["tag"]

    rank, local, world = parallel.env_rank_world()
    device = torch.device("cuda", local)
    torch.cuda.set_device(device)
    parallel.init(device=device)

    root_result_dir = os.path.join(args.out, cfg["name"])
    ckpt_dir = os.path.join(root_result_dir, "ckpts")
    if rank == 0:
        os.makedirs(ckpt_dir, exist_ok=True)
        copyfile(args.cfg, os.path.join(root_result_dir, os.path.basename(args.cfg)))

    print("Prepare data")
    train_loader, eval_loader = create_dataloader(
        data_path=args.data, num_scans=cfg["num_scans"], batch_size=cfg["batch_size"], num_workers=cfg["num_workers"],
        network_type=cfg["network"], train_with_val=cfg["train_with_val"],
        use_data_augumentation=cfg["use_data_augumentation"], cutout_kwargs=cfg["cutout_kwargs"],
        polar_grid_kwargs=cfg["polar_grid_kwargs"], pedestrian_only=cfg["pedestrian_only"],
        num_samples=args.num_samples, device=device)      # sharded across ranks when world > 1; pinned staging + async H2D

    model = SpatialDROW(num_scans=cfg["num_scans"], num_pts=cfg["cutout_kwargs"]["num_cutout_pts"],
                        focal_loss_gamma=cfg["focal_loss_gamma"], alpha=cfg["similarity_kwargs"]["alpha"],
                        window_size=cfg["similarity_kwargs"]["window_size"], pedestrian_only=cfg["pedestrian_only"])
    model.to(device)

    print("Prepare training")
    optimizer = optim.Adam(model.parameters(), lr=tu.lr_scheduler())
    starting_iteration, starting_epoch = 0, 0
    if args.ckpt is not None:
        starting_iteration, starting_epoch = load_checkpoint(model=model, optimizer=optimizer, filename=args.ckpt)
    model = parallel.wrap_ddp(model, device)

    tb_logger = create_tb_logger(root_result_dir) if rank == 0 else None
    trainer = Trainer(model, make_model_fn_obj_det(cfg["cutout_kwargs"]), optimizer, ckpt_dir, tu.ConstantLR(optimizer),
                      grad_norm_clip=cfg["grad_norm_clip"], tb_logger=tb_logger, is_main=rank == 0)
    last = trainer.train(num_epochs=cfg["epochs"], train_loader=train_loader,
                         ckpt_save_interval=max(int(cfg["epochs"] / 10), 1), starting_iteration=int(starting_iteration),
                         starting_epoch=max(int(starting_epoch), 0), max_iters=args.max_iters)
    if rank == 0:
        print("final loss %.6f" % last if last is not None else "no training iteration ran (empty loader)")
        if tb_logger is not None:
            tb_logger.close()


if __name__ == "__main__":
    main()
