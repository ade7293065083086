"""Data side of the DR-SPAAM entry points (reference: src/utils/dataset_dr_spaam.py:12-68,339-471).

What changed (SURVEY.md §7 hard part 7, row N2): the reference computes cutouts inside forked
DataLoader workers on the CPU; CUDA cannot run there, so a sample now carries the RAW ranges
(`scans` [S, N], already part of the reference's batch: dataset_dr_spaam.py:366,467) and the
cutouts are produced on the device, one launch per batch, inside `model_fn` (eval_utils.py here).
The batch keeps the reference's keys (`scans`, `target_cls`, `target_reg`, ...).

There is no DROW data (and no network) in this environment: the loaders serve seeded synthetic
DROW-shaped sequences (planar_optical_flow_b200/synth.py).  Reading the real DROWv2 files
(.csv/.wc/.wa/.wp/.odom2) is file I/O outside the accelerated path and is not implemented yet; a
`data_path` that exists raises NotImplementedError instead of silently substituting data.
"""
import os

import numpy as np
from torch.utils.data import DataLoader, Dataset

from . import synth


class SyntheticDROWDataset(Dataset):
    """Sequences of DROW-shaped scans with person-like leg arcs and per-point vote targets."""

    def __init__(self, split="train", num_scans=10, num_samples=256, seq_len=None, shape="drow", cutout_kwargs=None,
                 pedestrian_only=True, seed=0):
        self.num_scans, self.shape = num_scans, shape
        self.scan_phi = synth.phi_for(shape)
        self.n = len(self.scan_phi)
        self.num_samples = num_samples
        self.seed = seed + {"train": 0, "val": 10_000, "test": 20_000}.get(split, 30_000)
        self.cutout_kwargs = cutout_kwargs or {}

    def __len__(self):
        return self.num_samples

    def __getitem__(self, idx):
        s = self.num_scans + 1                                        # dataset_dr_spaam.py:366
        scans = synth.structured_sequence(s, self.n, seed=self.seed + idx, phi=self.scan_phi)
        rs = np.random.RandomState(self.seed + idx)
        cur = scans[-1]
        # foreground = points on close, narrow structures; their vote points at the structure centre
        near = cur < np.percentile(cur, 15)
        target_cls = near.astype(np.int64)
        target_reg = np.zeros((self.n, 2), dtype=np.float32)
        target_reg[near] = rs.normal(0.0, 0.1, size=(int(near.sum()), 2)).astype(np.float32)
        return {"scans": scans, "target_cls": target_cls, "target_reg": target_reg, "scan_phi": self.scan_phi,
                "idx": idx}

    @staticmethod
    def collate_batch(batch):
        out = {}
        for k in batch[0]:
            if k in ("scans", "target_cls", "target_reg"):
                out[k] = np.array([s[k] for s in batch])
            elif k == "scan_phi":
                out[k] = batch[0][k]
            else:
                out[k] = [s[k] for s in batch]
        return out


def _dataset(data_path, split, num_scans, cutout_kwargs, pedestrian_only, num_samples):
    if data_path and os.path.isdir(data_path):
        raise NotImplementedError("reading DROWv2 files from %s is not implemented in this build; "
                                  "omit the data directory to run on synthetic DROW-shaped sequences" % data_path)
    return SyntheticDROWDataset(split=split, num_scans=num_scans, num_samples=num_samples,
                                cutout_kwargs=cutout_kwargs, pedestrian_only=pedestrian_only)


def create_dataloader(data_path, num_scans, batch_size, num_workers, network_type="cutout", train_with_val=False,
                      use_data_augumentation=False, cutout_kwargs=None, polar_grid_kwargs=None, pedestrian_only=False,
                      num_samples=256, sampler=None):
    """Returns (train_loader, eval_loader_or_None), as dataset_dr_spaam.py:12-45."""
    train = _dataset(data_path, "train", num_scans, cutout_kwargs, pedestrian_only, num_samples)
    train_loader = DataLoader(train, batch_size=batch_size, pin_memory=False, num_workers=0, shuffle=sampler is None,
                              sampler=sampler, collate_fn=train.collate_batch, drop_last=True)
    if not train_with_val:
        return train_loader, None
    val = _dataset(data_path, "val", num_scans, cutout_kwargs, pedestrian_only, max(num_samples // 8, batch_size))
    return train_loader, DataLoader(val, batch_size=batch_size, num_workers=0, shuffle=False, collate_fn=val.collate_batch)


def create_test_dataloader(data_path, num_scans, network_type="cutout", cutout_kwargs=None, polar_grid_kwargs=None,
                           pedestrian_only=False, split="test", scan_stride=1, pt_stride=1, num_samples=32):
    test = _dataset(data_path, split, num_scans, cutout_kwargs, pedestrian_only, num_samples)
    return DataLoader(test, batch_size=1, num_workers=0, shuffle=False, collate_fn=test.collate_batch)
