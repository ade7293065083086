"""Scan-pair data for the flow prototype (reference: src/utils/dataset.py:19-110, FlowDataset).

A sample is a pair of consecutive scans as xy point sets — the second one moved into the first
one's frame with the odometry increment (dataset.py:77-93) — and a per-point flow target
(dataset.py:74).  The batch keeps the reference's keys (`scan_pair` [B, 2, N, 2], `flow_target`
[B, N, 2]; the reference's training closure looks for `flow_target_flow`).

`DROWFlowDataset` reads the real files (`<split>/*.csv` scans, `.difodom` odometry increments, `.odom2` poses and
the `.flow` targets written by the reference's offline bin/data_prepare.py).  When the data directory does not exist
(there is no DROW data in the build environment) the loader serves seeded synthetic pairs: a structured scan, a small
rigid ego motion, and the flow that motion induces on every point.
"""
import os

import numpy as np
from torch.utils.data import DataLoader, Dataset

from . import drow_io, synth
from .utils import get_laser_phi, rphi_to_xy


def _collate(batch):                                                  # dataset.py:100-108
    out = {}
    for k in batch[0]:
        if k in ("scan_pair", "flow_target"):
            out[k] = np.stack([s[k] for s in batch], axis=0)
        else:
            out[k] = [s[k] for s in batch]
    return out


class DROWFlowDataset(Dataset):
    """The reference's `FlowDataset` (dataset.py:19-108) on real DROWv2 recordings: every scan of the first
    `max_sequences` recordings paired with its successor (the last scan with itself), the successor rotated by the
    odometry increment's angle and shifted by its translation turned into the scan direction (:77-93)."""

    def __init__(self, data_path, split="train", testing=False, train_with_val=False, max_sequences=5):
        stems = drow_io.sequence_stems(data_path, split)
        if train_with_val:
            stems += drow_io.sequence_stems(data_path, "val")
        self.seq_names = stems[:max_sequences]
        if not self.seq_names:
            raise FileNotFoundError("{}: No valid data".format(split))
        scans, nxt, odoms_t, odoms, dirs, flows = [], [], [], [], [], []
        for stem in self.seq_names:
            _, _, sc = drow_io.load_scan_file(stem)
            scans.append(sc)
            nxt.append(np.vstack([sc[1:], sc[-1].reshape(1, -1)]))                                    # :38
            t, d = drow_io.load_difodom(stem)
            odoms_t.append(t)
            odoms.append(d)
            dirs.append(drow_io.load_odom2(stem)[2][..., -1])                                          # :50-52
            flows.append(drow_io.load_flow_file(stem, sc.shape[-1]))
        self.scans, self.scans_next = np.vstack(scans), np.vstack(nxt)
        self.odoms_t, self.odoms = np.hstack(odoms_t), np.vstack(odoms)
        self.scan_dir = np.hstack(dirs)
        self.flow_targets = np.vstack(flows)

    def __len__(self):
        return len(self.scans)

    def __getitem__(self, idx):
        phi = get_laser_phi()
        odom, scan_dir = self.odoms[idx], self.scan_dir[idx]
        xy = np.stack(rphi_to_xy(self.scans[idx], phi), axis=1)
        xy_next = np.stack(rphi_to_xy(self.scans_next[idx], phi), axis=1)
        rot = np.array([[np.cos(odom[-1]), np.sin(odom[-1])], [-np.sin(odom[-1]), np.cos(odom[-1])]], dtype=np.float32)
        to_scan = np.array([[np.cos(scan_dir), -np.sin(scan_dir)], [np.sin(scan_dir), np.cos(scan_dir)]], dtype=np.float32)
        trans = np.matmul(odom[:-1], to_scan.T)
        return {"phi_grid": phi, "scan_pair": [xy, np.matmul(xy_next, rot.T) + trans], "odom_t": self.odoms_t[idx],
                "odom": odom, "flow_target": self.flow_targets[idx]}

    collate_batch = staticmethod(_collate)


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

    collate_batch = staticmethod(_collate)


def FlowDataset(data_path=None, split="train", testing=False, train_with_val=False, num_samples=512):
    """Same call as the reference's FlowDataset(data_path, split, ...): real recordings when `data_path` is a DROWv2
    directory, seeded synthetic pairs when it does not exist."""
    if drow_io.has_drow_files(data_path):
        return DROWFlowDataset(data_path, split=split, testing=testing, train_with_val=train_with_val)
    if data_path and os.path.isdir(data_path) and os.listdir(data_path):
        raise FileNotFoundError("%s exists but holds no <split>/*.csv DROWv2 recordings" % data_path)
    return SyntheticFlowDataset(split=split, num_samples=num_samples)


def create_flow_dataloader(batch_size, split="train", num_samples=512, sampler=None, data_path=None):
    ds = FlowDataset(data_path, split=split, num_samples=num_samples)
    return DataLoader(ds, batch_size=batch_size, shuffle=sampler is None and split == "train", sampler=sampler, num_workers=0,
                      collate_fn=ds.collate_batch, drop_last=split == "train")
