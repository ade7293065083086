"""Data side of the DR-SPAAM entry points (reference: src/utils/dataset_dr_spaam.py:12-68, 256-529).

What changed against the reference (SURVEY.md §7 hard part 7, row N2): the reference computes cutouts inside forked
DataLoader workers on the CPU; CUDA cannot run there, so a sample carries the RAW ranges (`scans` [S, N], already part
of the reference's batch: dataset_dr_spaam.py:366,467) and the cutouts are produced on the device, one launch per
batch, inside `model_fn` (eval_utils.py here).  The batch keeps the reference's keys (`scans`, `target_cls`,
`target_reg`, `target_flow`, `exclude_mask`, ...); only `input` is gone.

Two datasets behind the same `create_dataloader` call:

  * `DROWDataset`          the real DROWv2 recordings (`<split>/*.csv` + `.wc/.wa/.wp` annotations + `.odom2` poses):
                           the reference's `DROWDataset2` - static stretches removed, one sample per annotated scan,
                           `num_scans` history scans + the current one, detection and odometry-flow targets, dynamic /
                           valid-range masks - read by `drow_io`;
  * `SyntheticDROWDataset` seeded synthetic DROW-shaped sequences (`synth.py`), used when the data directory does not
                           exist (there is no DROW data, and no network, in the build environment).

`DeviceBatches` wraps a loader for training on a GPU: the large arrays of every batch are copied into rotating PINNED
staging buffers and sent to the device with non-blocking copies, so the H2D transfer of batch k+1 overlaps the step on
batch k; the batch then holds CUDA tensors under the same keys.
"""
import os

import numpy as np
import torch
from torch.utils.data import DataLoader, Dataset
from torch.utils.data.distributed import DistributedSampler

from . import drow_io, synth
from . import utils as u

_ARRAY_KEYS = ("scans", "target_cls", "target_reg", "input", "target_flow", "exclude_mask", "odom")     # :464-468


def collate_batch(batch):
    """Stack the per-point arrays, keep everything else as lists (dataset_dr_spaam.py:462-471)."""
    out = {}
    for k in batch[0]:
        if k in _ARRAY_KEYS:
            out[k] = np.array([s[k] for s in batch])
        elif k == "scan_phi":                      # one angle grid per batch (what the device cutout call takes)
            out[k] = batch[0][k]
        else:
            out[k] = [s[k] for s in batch]
    return out


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

    collate_batch = staticmethod(collate_batch)


class DROWDataset(Dataset):
    """The reference's `DROWDataset2` (dataset_dr_spaam.py:256-529) on real DROWv2 files, minus the CPU cutout.

    Construction (:266-337): the first `max_sequences` recordings of the split; odometry poses that do not change to
    the next one mark static stretches, which are dropped together with their scans (a recording that never moves is
    dropped entirely); one sample per annotated scan that survived.
    A sample (:342-459): the `num_scans` scans `scan_stride` apart that end `distance` = 5 strides before the annotated
    scan, plus the annotated scan itself; detection targets for the annotated scan; the apparent flow of static points
    between the last history scan and the annotated one from odometry, in each beam's canonical frame; `exclude_mask` =
    0 within 2.5 / 2.0 / 2.0 m of an annotated wheelchair / walker / person or at ranges >= 20 m.
    """

    DISTANCE = 5                  # the reference fixes it (:361), `max_scan_dist` is kept for the signature

    def __init__(self, data_path, split="train", num_scans=5, network_type="cutout", train_with_val=False, cutout_kwargs=None,
                 polar_grid_kwargs=None, use_data_augumentation=False, pedestrian_only=False, scan_stride=1, pt_stride=1,
                 max_scan_dist=6, max_sequences=5, seed=None):
        self._num_scans, self._scan_stride, self._pt_stride = num_scans, scan_stride, pt_stride
        self._use_data_augmentation = use_data_augumentation
        self._cutout_kwargs, self._polar_grid_kwargs = cutout_kwargs, polar_grid_kwargs
        self._network_type, self._pedestrian_only = network_type, pedestrian_only
        self.max_scan_dist = max_scan_dist
        self._rng = np.random.RandomState(seed) if seed is not None else np.random

        stems = drow_io.sequence_stems(data_path, split, limit=max_sequences)
        self.seq_names, self.odoms_t, self.odoms = [], [], []
        self.scans_ns, self.scans_t, self.scans = [], [], []
        self.dets_ns, self.dets_wc, self.dets_wa, self.dets_wp = [], [], [], []
        for stem in stems:
            _, odom_t, odom = drow_io.load_odom2(stem)
            if len(odom) == 0:
                continue
            moving = np.hstack([np.any((odom[1:] - odom[:-1]) != 0.0, axis=1), False])                  # :284
            if not np.any(moving):
                continue                                                                                # static scene
            ns, ts, sc = drow_io.load_scan_file(stem)
            d_ns, wc, wa, wp = drow_io.load_det_file(stem)
            self.seq_names.append(stem)
            self.odoms_t.append(odom_t[moving])
            self.odoms.append(odom[moving])
            self.scans_ns.append(ns[moving])
            self.scans_t.append(ts[moving])
            self.scans.append(sc[moving])
            self.dets_ns.append(d_ns)
            self.dets_wc.append(wc)
            self.dets_wa.append(wa)
            self.dets_wp.append(wp)
        if not self.seq_names:
            raise FileNotFoundError("{}: No valid data".format(split))                                # :295
        print("{}: {} valid files found".format(split, len(self.seq_names)))

        # annotated scans that survived the filter, in annotation order (:323-337).  NOTE the reference numbers the
        # SURVIVORS 0..k-1 and then indexes the unfiltered annotation lists with that number (so after the first
        # dropped annotation its samples pair a scan with another scan's boxes); here a sample keeps the annotation
        # index it came from.
        self.idet2iscan, self.flat_seq_inds, self.flat_det_inds = [], [], []
        for seq_idx, (ss, ds) in enumerate(zip(self.scans_ns, self.dets_ns)):
            pos = {int(n): i for i, n in reversed(list(enumerate(ss)))}                               # first occurrence
            table = [(j, pos[int(d)]) for j, d in enumerate(ds) if int(d) in pos]
            self.idet2iscan.append(table)
            self.flat_seq_inds += [seq_idx] * len(table)
            self.flat_det_inds += range(len(table))

    def __len__(self):
        return len(self.flat_det_inds)

    @property
    def scan_phi(self):
        return u.get_laser_phi()[::self._pt_stride]

    def __getitem__(self, idx):
        seq_idx = self.flat_seq_inds[idx]
        det_idx, scan_idx = self.idet2iscan[seq_idx][self.flat_det_inds[idx]]
        seq_scans = self.scans[seq_idx]
        out = {"seq_name": self.seq_names[seq_idx], "dets_ns": self.dets_ns[seq_idx][det_idx],
               "dets_wc": self.dets_wc[seq_idx][det_idx], "dets_wa": self.dets_wa[seq_idx][det_idx],
               "dets_wp": self.dets_wp[seq_idx][det_idx]}

        cur_scan = seq_scans[scan_idx]
        back = (np.arange(self._num_scans + self.DISTANCE) * self._scan_stride)[::-1]                  # :364
        scan_inds = [max(0, scan_idx - int(i)) for i in back[:self._num_scans]]
        scans = seq_scans[scan_inds][:, ::self._pt_stride]
        out["scans"] = np.vstack((scans, cur_scan))                                                    # :370
        out["scans_ns"] = [self.scans_ns[seq_idx][i] for i in scan_inds]

        t1, t0 = self.scans_t[seq_idx][scan_idx], self.scans_t[seq_idx][scan_inds[-1]]                  # :374-377
        i1 = int(np.argmin(np.abs(self.odoms_t[seq_idx] - t1)))
        i0 = int(np.argmin(np.abs(self.odoms_t[seq_idx] - t0)))
        odom1, odom0 = self.odoms[seq_idx][i1], self.odoms[seq_idx][i0]
        out["odom1_t"], out["odom1"] = self.odoms_t[seq_idx][i1], odom1

        scan_phi = self.scan_phi
        out["phi_grid"] = out["scan_phi"] = scan_phi
        out["target_cls"], out["target_reg"] = u.get_regression_target(
            cur_scan, scan_phi, out["dets_wc"], out["dets_wa"], out["dets_wp"], pedestrian_only=self._pedestrian_only)

        cur_xy = np.array(u.rphi_to_xy(cur_scan, scan_phi)).T                                          # :400-403
        out["target_flow"] = u.global_to_canonical_flow(u.get_displacement_from_odometry(cur_xy, odom0, odom1), scan_phi)
        out["exclude_mask"] = self._dynamic_mask(cur_xy, out["dets_wc"], out["dets_wa"], out["dets_wp"]) * self._valid_mask(cur_scan)

        if self._use_data_augmentation:
            out = u.data_augmentation(out, self._rng)
        return out

    collate_batch = staticmethod(collate_batch)

    @staticmethod
    def _dynamic_mask(scan_xy, dets_wc, dets_wa, dets_wp, radius_wc=2.5, radius_wa=2.0, radius_wp=2.0):
        """0 for points within `radius` of an annotated (possibly moving) object, 1 elsewhere (:510-522)."""
        mask = np.ones(len(scan_xy), dtype=np.float64)
        for dets, radius in ((dets_wc, radius_wc), (dets_wa, radius_wa), (dets_wp, radius_wp)):
            for det in dets:
                centre = np.hstack(u.rphi_to_xy(det[0], det[1]))
                mask[np.linalg.norm(scan_xy - centre, axis=-1) <= radius] = 0.0
        return mask

    @staticmethod
    def _valid_mask(scan, thresh=20.0):
        mask = np.ones_like(scan)                                                                      # :524-528
        mask[scan >= 20.0] = 0.0
        return mask


DROWDataset2 = DROWDataset          # the reference's class name


class DeviceBatches:
    """Iterate a loader and hand out batches whose array entries already live on `device`.

    Each array of a batch is copied into a pinned host buffer (one of `depth` rotating sets, allocated on first use and
    re-used while the shape stays the same) and sent with a non-blocking copy on the current stream; an event per set
    guards its re-use.  Keys and shapes are unchanged, values become CUDA tensors."""

    def __init__(self, loader, device, keys=_ARRAY_KEYS, depth=3):
        self.loader, self.device, self.keys, self.depth = loader, torch.device(device), tuple(keys), depth
        self._sets = [dict() for _ in range(depth)]
        self._events = [None] * depth
        self._turn = 0
        self.h2d_bytes = 0

    def __len__(self):
        return len(self.loader)

    @property
    def sampler(self):
        return self.loader.sampler

    @property
    def dataset(self):
        return self.loader.dataset

    def _stage(self, slot, key, arr):
        arr = np.ascontiguousarray(arr)
        if arr.dtype == np.float64 and key != "odom":
            arr = arr.astype(np.float32)
        buf = self._sets[slot].get(key)
        if buf is None or tuple(buf.shape) != arr.shape or buf.dtype != torch.from_numpy(arr[:0]).dtype:
            buf = self._sets[slot][key] = torch.empty(arr.shape, dtype=torch.from_numpy(arr[:0]).dtype).pin_memory()
        buf.numpy()[...] = arr
        self.h2d_bytes += buf.numel() * buf.element_size()
        return buf.to(self.device, non_blocking=True)

    def __iter__(self):
        for batch in self.loader:
            slot = self._turn
            self._turn = (self._turn + 1) % self.depth
            if self._events[slot] is not None:
                self._events[slot].synchronize()              # the copies that last read this pinned set are done
            out = dict(batch)
            for k in self.keys:
                if k in batch and isinstance(batch[k], np.ndarray) and batch[k].dtype != object:
                    out[k] = self._stage(slot, k, batch[k])
            ev = torch.cuda.Event()
            ev.record(torch.cuda.current_stream(self.device))
            self._events[slot] = ev
            yield out


def _dataset(data_path, split, num_scans, cutout_kwargs, pedestrian_only, num_samples, **real_kwargs):
    if drow_io.has_drow_files(data_path):
        return DROWDataset(data_path, split=split, num_scans=num_scans, cutout_kwargs=cutout_kwargs,
                           pedestrian_only=pedestrian_only, **real_kwargs)
    if data_path and os.path.isdir(data_path) and os.listdir(data_path):
        raise FileNotFoundError("%s exists but holds no <split>/*.csv DROWv2 recordings; pass a DROWv2 directory, or a "
                                "path that does not exist to run on synthetic DROW-shaped sequences" % data_path)
    return SyntheticDROWDataset(split=split, num_scans=num_scans, num_samples=num_samples,
                                cutout_kwargs=cutout_kwargs, pedestrian_only=pedestrian_only)


def _distributed_sampler(dataset, shuffle):
    import torch.distributed as dist

    if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
        return DistributedSampler(dataset, num_replicas=dist.get_world_size(), rank=dist.get_rank(), shuffle=shuffle)
    return None


def create_dataloader(data_path, num_scans, batch_size, num_workers, network_type="cutout", train_with_val=False,
                      use_data_augumentation=False, cutout_kwargs=None, polar_grid_kwargs=None, pedestrian_only=False,
                      num_samples=256, sampler=None, device=None):
    """Returns (train_loader, eval_loader_or_None), as dataset_dr_spaam.py:12-45.

    Under an initialised multi-process group the samples are sharded with a DistributedSampler (each rank sees its own
    1 / world of every epoch; `Trainer.train` re-seeds it per epoch).  `device`: wrap the loaders in `DeviceBatches`.
    The worker processes only read and label ranges (no CUDA there), so `num_workers` applies to real data as in the
    reference; the synthetic set is generated in-process."""
    train = _dataset(data_path, "train", num_scans, cutout_kwargs, pedestrian_only, num_samples, network_type=network_type,
                     train_with_val=train_with_val, use_data_augumentation=use_data_augumentation,
                     polar_grid_kwargs=polar_grid_kwargs)
    real = isinstance(train, DROWDataset)
    if sampler is None:
        sampler = _distributed_sampler(train, shuffle=True)
    train_loader = DataLoader(train, batch_size=batch_size, pin_memory=False, num_workers=num_workers if real else 0,
                              shuffle=sampler is None, sampler=sampler, collate_fn=collate_batch, drop_last=True)
    eval_loader = None
    if train_with_val:
        val = _dataset(data_path, "val", num_scans, cutout_kwargs, pedestrian_only, max(num_samples // 8, batch_size),
                       network_type=network_type, polar_grid_kwargs=polar_grid_kwargs)
        eval_loader = DataLoader(val, batch_size=batch_size, num_workers=0, shuffle=False, collate_fn=collate_batch)
    if device is not None:
        train_loader = DeviceBatches(train_loader, device)
        eval_loader = DeviceBatches(eval_loader, device) if eval_loader is not None else None
    return train_loader, eval_loader


def create_test_dataloader(data_path, num_scans, network_type="cutout", cutout_kwargs=None, polar_grid_kwargs=None,
                           pedestrian_only=False, split="test", scan_stride=1, pt_stride=1, num_samples=32):
    test = _dataset(data_path, split, num_scans, cutout_kwargs, pedestrian_only, num_samples, network_type=network_type,
                    polar_grid_kwargs=polar_grid_kwargs, scan_stride=scan_stride, pt_stride=pt_stride)
    return DataLoader(test, batch_size=1, num_workers=0, shuffle=False, collate_fn=collate_batch)
