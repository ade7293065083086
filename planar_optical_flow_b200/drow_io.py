"""Readers for the DROWv2 file family (reference: src/utils/dataset_dr_spaam.py:473-508, src/utils/dataset.py:109-132).

One recording `<stem>` is a set of sibling text files:

    <stem>.csv      one scan per line:  seq, time, r_0 ... r_{N-1}                 (:473-478)
    <stem>.wc/.wa/.wp   one annotated scan per line:  seq, [[r, phi], ...]   (JSON tail; wheelchairs / walkers / persons, :480-494)
    <stem>.odom2    seq, time, x, y, phi   - odometry sampled at the scans          (:503-508)
    <stem>.difodom  time, dx, dy, dphi     - odometry increments (flow prototype)   (dataset.py:121-126)
    <stem>.flow     N * 2 values per scan  - flow targets written by bin/data_prepare.py (dataset.py:128-131)

The values and dtypes returned are the reference's (uint32 sequence numbers, float32 times, ranges and poses); parsing
goes through pandas' C reader when available (a DROW recording is ~10^4 lines of 452 numbers: `np.genfromtxt` needs
minutes per file) and falls back to NumPy's text readers.  Pure host-side I/O: nothing here touches the GPU.
"""
import json
import os

import numpy as np

try:                                    # optional fast path
    import pandas as _pd
except Exception:                       # noqa: BLE001
    _pd = None


def _table(path):
    """A numeric comma-separated file as a 2-D float64 array."""
    if os.path.getsize(path) == 0:
        return np.zeros((0, 0), dtype=np.float64)
    if _pd is not None:
        arr = _pd.read_csv(path, header=None, dtype=np.float64).to_numpy()
    else:
        arr = np.loadtxt(path, delimiter=",", dtype=np.float64, ndmin=2)
    return np.ascontiguousarray(arr)


def load_scan_file(stem):
    """`<stem>.csv` -> (seq [T] uint32, time [T] float32, scans [T, N] float32)."""
    data = _table(stem + ".csv")
    return data[:, 0].astype(np.uint32), data[:, 1].astype(np.float32), data[:, 2:].astype(np.float32)


def _load_annotations(path):
    seqs, dets = [], []
    with open(path) as f:
        for line in f:
            if not line.strip():
                continue
            seq, tail = line.split(",", 1)
            seqs.append(int(seq))
            dets.append(json.loads(tail))
    return seqs, dets


def load_det_file(stem):
    """`<stem>.wc/.wa/.wp` -> (seq [K] int, wcs, was, wps): per annotated scan a list of (r, phi) detections per class.
    The three files must annotate the same scans in the same order (the reference asserts it, :491-492)."""
    s1, wcs = _load_annotations(stem + ".wc")
    s2, was = _load_annotations(stem + ".wa")
    s3, wps = _load_annotations(stem + ".wp")
    if not (s1 == s2 == s3):
        raise ValueError("%s: the .wc / .wa / .wp files annotate different scans" % stem)
    return np.array(s1), wcs, was, wps


def load_odom2(stem):
    """`<stem>.odom2` -> (seq [T] uint32, time [T] float32, pose [T, 3] float32 = x, y, phi)."""
    data = _table(stem + ".odom2")
    if data.size == 0:
        return np.zeros(0, np.uint32), np.zeros(0, np.float32), np.zeros((0, 3), np.float32)
    return data[:, 0].astype(np.uint32), data[:, 1].astype(np.float32), data[:, 2:5].astype(np.float32)


def load_difodom(stem):
    """`<stem>.difodom` -> (time [T] float64, increments [T, 3] float64)."""
    data = _table(stem + ".difodom")
    return data[:, 0], data[:, 1:]


def load_flow_file(stem, n_pts):
    """`<stem>.flow` -> [T, n_pts, 2] float64."""
    return _table(stem + ".flow").reshape(-1, n_pts, 2)


def sequence_stems(data_path, split, limit=None):
    """Recording stems of a split, sorted by name (the reference takes `glob` order, which is the file system's;
    sorting makes runs reproducible).  `limit`: the reference trains on the first five (:274)."""
    d = os.path.join(data_path, split)
    stems = sorted(os.path.join(d, f[:-4]) for f in os.listdir(d) if f.endswith(".csv")) if os.path.isdir(d) else []
    return stems if limit is None else stems[:limit]


def has_drow_files(data_path):
    """True if `data_path` looks like a DROWv2 directory (at least one split with a .csv recording)."""
    if not data_path or not os.path.isdir(data_path):
        return False
    return any(sequence_stems(data_path, s) for s in ("train", "val", "test"))
