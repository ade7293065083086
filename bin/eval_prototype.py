#!/usr/bin/env python
"""Scan-pair flow prototype evaluation:  python bin/eval_prototype.py [--ckpt x.pth]   (reference: bin/eval_prototype.py)

Loads `ckpt["model_state"]` (the reference's checkpoint layout), runs the test pairs through the
prototype and reports mean end-point / angular error.
"""
import argparse
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

from src.depracted.model import Prototype  # noqa: E402
from src.utils.dataset import FlowDataset  # noqa: E402
from src.utils.eval_utils import model_fn_eval  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--data", default="./../data/DROWv2-data")
    ap.add_argument("--ckpt", default=None)
    ap.add_argument("--num-samples", type=int, default=200)
    ap.add_argument("--batch-size", type=int, default=100)
    args = ap.parse_args()
    device = torch.device("cuda", 0)
    test = FlowDataset(args.data, split="test", num_samples=args.num_samples)
    loader = torch.utils.data.DataLoader(test, batch_size=args.batch_size, shuffle=False, num_workers=0,
                                         collate_fn=test.collate_batch)
    model = Prototype(in_channel=2).to(device)
    if args.ckpt is not None:
        model.load_state_dict(torch.load(args.ckpt, map_location=device)["model_state"])
    epe, aae = model_fn_eval(model, loader)
    print(json.dumps({"pairs": len(test), "epe": epe, "aae_deg": aae}))


if __name__ == "__main__":
    main()
