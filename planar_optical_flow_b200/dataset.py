"""Scan-pair data for the flow prototype (reference: src/utils/dataset.py:19-110, FlowDataset).

A sample is a pair of consecutive scans as xy point sets — the second one moved into the first
one's frame with the odometry increment (dataset.py:77-93) — and a per-point flow target
(dataset.py:74).  The batch keeps the reference's keys (`scan_pair` [B, 2, N, 2], `flow_target`
[B, N, 2]; the reference's training closure looks for `flow_target_flow`).

There is no DROW data in this environment (the `.flow` targets come out of the reference's offline
bin/data_prepare.py), so the loader serves seeded synthetic pairs: a structured scan, a small rigid
ego motion, and the flow that motion induces on every point.
"""
import os

import numpy as np
from torch.utils.data import DataLoader, Dataset

from . import synth
from .utils import rphi_to_xy


class SyntheticFlowDataset(Dataset):
    def __init__(self, split="train", num_samples=512, shape="drow", seed=0):
        self.scan_phi = synth.phi_for(shape)
        self.n = len(self.scan_phi)
        self.num_samples = num_samples
        self.seed = seed + {"train": 0, "val": 10_000, "test": 20_000}.get(split, 30_000)

    def __len__(self):
        return self.num_samples

    def __getitem__(self, idx):
        rs = np.random.RandomState(self.seed + idx)
        scans = synth.structured_sequence(2, self.n, seed=self.seed + idx, phi=self.scan_phi)
        xy = np.stack(rphi_to_xy(scans[0], self.scan_phi), axis=1).astype(np.float32)          # dataset.py:79-81
        xy_next = np.stack(rphi_to_xy(scans[1], self.scan_phi), axis=1).astype(np.float32)
        dtheta = rs.uniform(-0.05, 0.05)
        trans = rs.uniform(-0.1, 0.1, size=2).astype(np.float32)
        rot = np.array([[np.cos(dtheta), np.sin(dtheta)], [-np.sin(dtheta), np.cos(dtheta)]], dtype=np.float32)   # :84-85
        xy_next_rot = xy_next @ rot.T + trans                                                   # :91
        flow = (xy @ rot.T + trans - xy).astype(np.float32)          # what the ego motion does to every point of scan 1
        return {"scan_pair": [xy, xy_next_rot.astype(np.float32)], "flow_target": flow, "phi_grid": self.scan_phi,
                "odom": np.array([trans[0], trans[1], dtheta], dtype=np.float32)}

    @staticmethod
    def collate_batch(batch):                                         # dataset.py:100-108
        out = {}
        for k in batch[0]:
            if k in ("scan_pair", "flow_target"):
                out[k] = np.stack([s[k] for s in batch], axis=0)
            else:
                out[k] = [s[k] for s in batch]
        return out


def FlowDataset(data_path=None, split="train", testing=False, train_with_val=False, num_samples=512):
    """Same call as the reference's FlowDataset(data_path, split, ...)."""
    if data_path and os.path.isdir(data_path):
        raise NotImplementedError("reading DROWv2 .csv/.odom2/.flow files from %s is not implemented in this build; "
                                  "omit the data directory to run on synthetic scan pairs" % data_path)
    return SyntheticFlowDataset(split=split, num_samples=num_samples)


def create_flow_dataloader(batch_size, split="train", num_samples=512, sampler=None, data_path=None):
    ds = FlowDataset(data_path, split=split, num_samples=num_samples)
    return DataLoader(ds, batch_size=batch_size, shuffle=sampler is None and split == "train", sampler=sampler, num_workers=0,
                      collate_fn=ds.collate_batch, drop_last=split == "train")
