#!/usr/bin/env python
"""DR-SPAAM evaluation entry point:  python bin/eval_dr_spaam.py --cfg config/dr_spaam.yaml [--ckpt x.pth]

Same CLI and YAML schema as the reference's bin/eval_dr_spaam.py (:24-31).  Loads
`ckpt["model_state"]` (reference checkpoints load unchanged; random init if no --ckpt), streams the
test sequences through SpatialDROW(testing=True) with the attention memory carried from scan to
scan, and post-processes every scan with the NMS kernel.
"""
import argparse
import os
import sys
from shutil import copyfile

import torch
import yaml

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

from src.depracted.model import SpatialDROW  # noqa: E402
from src.utils.dataset_dr_spaam import create_test_dataloader  # noqa: E402
from src.utils.eval_utils import eval_dr_spaam  # noqa: E402


def main():
    parser = argparse.ArgumentParser(description="arg parser")
    parser.add_argument("--cfg", type=str, required=True, help="configuration of the experiment")
    parser.add_argument("--ckpt", type=str, required=False, default=None)
    parser.add_argument("--data", type=str, default="./../data/DROWv2-data")
    parser.add_argument("--out", type=str, default=os.path.join("./..", "output"))
    parser.add_argument("--num-samples", type=int, default=16)
    args = parser.parse_args()

    with open(args.cfg, "r") as f:
        cfg = yaml.safe_load(f)
    cfg["name"] = os.path.basename(args.cfg).split(".")[0] + cfg["tag"]
    root_result_dir = os.path.join(args.out, cfg["name"])
    os.makedirs(root_result_dir, exist_ok=True)
    copyfile(args.cfg, os.path.join(root_result_dir, os.path.basename(args.cfg)))

    print("Prepare data")
    test_loader = create_test_dataloader(data_path=args.data, num_scans=cfg["num_scans"], network_type=cfg["network"],
                                         cutout_kwargs=cfg["cutout_kwargs"], polar_grid_kwargs=cfg["polar_grid_kwargs"],
                                         pedestrian_only=cfg["pedestrian_only"], split="test",
                                         num_samples=args.num_samples)
    print("Prepare model")
    model = SpatialDROW(num_scans=cfg["num_scans"], num_pts=cfg["cutout_kwargs"]["num_cutout_pts"],
                        focal_loss_gamma=cfg["focal_loss_gamma"], alpha=cfg["similarity_kwargs"]["alpha"],
                        window_size=cfg["similarity_kwargs"]["window_size"], pedestrian_only=cfg["pedestrian_only"])
    model.cuda()
    if args.ckpt is not None:
        model.load_state_dict(torch.load(args.ckpt, map_location="cuda")["model_state"])
    model.eval()

    print("Start testing")
    summary, _ = eval_dr_spaam(model, test_loader=test_loader, cfg=cfg, output_dir=root_result_dir)
    print(summary)


if __name__ == "__main__":
    main()
