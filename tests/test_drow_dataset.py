"""The DROWv2 readers and `DROWDataset` (row N2) on a tiny recording written in the DROW file formats, against the
reference's own `DROWDataset2` when the reference tree is mounted, and on their own everywhere else."""
import json
import os

import numpy as np
import pytest

from planar_optical_flow_b200 import dataset_dr_spaam as dds
from planar_optical_flow_b200 import drow_io, synth
from planar_optical_flow_b200 import utils as u

CUTOUT = dict(fixed=True, centered=True, window_width=1.0, window_depth=0.5, num_cutout_pts=56, padding_val=29.99,
              area_mode=True)


def write_recording(stem, seed, n_scans=40, moving=True, annotate_every=4, first_annotated=12):
    """One synthetic recording in the DROWv2 formats (.csv / .wc / .wa / .wp / .odom2)."""
    rs = np.random.RandomState(seed)
    phi = u.get_laser_phi()
    scans = synth.structured_sequence(n_scans, len(phi), seed=seed, phi=phi)
    seqs = 1000 + 2 * np.arange(n_scans)
    times = 10.0 + 0.08 * np.arange(n_scans)
    with open(stem + ".csv", "w") as f:
        for s, t, r in zip(seqs, times, scans):
            f.write("%d,%.4f,%s\n" % (s, t, ",".join("%.3f" % v for v in r)))
    pose = np.zeros((n_scans, 3))
    if moving:
        step = np.concatenate([np.zeros((5, 3)), np.tile([0.03, 0.01, 0.004], (n_scans - 5, 1))])      # still for 5 scans, then moving
        pose = np.cumsum(step, axis=0)
    with open(stem + ".odom2", "w") as f:
        for s, t, p in zip(seqs, times, pose):
            f.write("%d,%.4f,%.5f,%.5f,%.5f\n" % (s, t, p[0], p[1], p[2]))
    files = {ext: open(stem + ext, "w") for ext in (".wc", ".wa", ".wp")}
    for k in range(first_annotated, n_scans - 1, annotate_every):
        for ext, count in ((".wc", k % 2), (".wa", 1 if k % 3 == 0 else 0), (".wp", 2)):
            dets = []
            for _ in range(count):
                i = int(rs.randint(30, len(phi) - 30))
                dets.append([float(scans[k, i]) + 0.1, float(phi[i])])
            files[ext].write("%d,%s\n" % (seqs[k], json.dumps(dets)))
    for f in files.values():
        f.close()


@pytest.fixture()
def drow_dir(tmp_path):
    d = tmp_path / "DROWv2-data"
    (d / "train").mkdir(parents=True)
    write_recording(str(d / "train" / "rec_a"), seed=1)
    write_recording(str(d / "train" / "rec_b"), seed=2, n_scans=30)
    write_recording(str(d / "train" / "rec_static"), seed=3, moving=False)
    return str(d)


def test_readers_return_the_reference_dtypes(drow_dir):
    stem = os.path.join(drow_dir, "train", "rec_a")
    ns, ts, scans = drow_io.load_scan_file(stem)
    assert ns.dtype == np.uint32 and ts.dtype == np.float32 and scans.dtype == np.float32 and scans.shape == (40, 450)
    d_ns, wc, wa, wp = drow_io.load_det_file(stem)
    assert len(d_ns) == len(wc) == len(wa) == len(wp) and all(len(p) == 2 for p in wp)
    o_ns, o_t, pose = drow_io.load_odom2(stem)
    assert o_ns.dtype == np.uint32 and pose.dtype == np.float32 and pose.shape == (40, 3)
    assert drow_io.has_drow_files(drow_dir) and not drow_io.has_drow_files(os.path.join(drow_dir, "nope"))
    assert [os.path.basename(s) for s in drow_io.sequence_stems(drow_dir, "train")] == ["rec_a", "rec_b", "rec_static"]


def test_dataset_samples_and_loader(drow_dir):
    ds = dds.DROWDataset(drow_dir, split="train", num_scans=10, cutout_kwargs=CUTOUT, pedestrian_only=True)
    assert [os.path.basename(s) for s in ds.seq_names] == ["rec_a", "rec_b"]          # the static recording is dropped
    assert len(ds) > 0
    s = ds[0]
    assert s["scans"].shape == (11, 450) and s["scans"].dtype == np.float32
    assert s["target_cls"].shape == (450,) and s["target_cls"].dtype == np.int64 and set(np.unique(s["target_cls"])) <= {0, 1}
    assert s["target_reg"].shape == (450, 2) and s["target_reg"].dtype == np.float32
    assert s["target_flow"].shape == (450, 2) and s["exclude_mask"].shape == (450,)
    assert np.array_equal(s["scans"][-1], ds.scans[0][ds.idet2iscan[0][0][1]])         # the annotated scan comes last
    assert s["target_cls"].sum() > 0 and np.abs(s["target_reg"][s["target_cls"] > 0]).max() < 0.5
    assert np.all(s["target_reg"][s["target_cls"] == 0] == 0)
    # every scan of a sample comes from the moving stretch (the first 5 scans of a recording stand still)
    kept = set(int(n) for n in ds.scans_ns[0])
    assert all(int(n) in kept for n in s["scans_ns"]) and 1000 not in kept
    train, val = dds.create_dataloader(drow_dir, num_scans=10, batch_size=4, num_workers=0, cutout_kwargs=CUTOUT,
                                       pedestrian_only=True)
    batch = next(iter(train))
    assert val is None and batch["scans"].shape == (4, 11, 450) and batch["target_reg"].shape == (4, 450, 2)
    assert batch["scan_phi"].shape == (450,) and len(batch["phi_grid"]) == 4 and len(batch["seq_name"]) == 4


def test_missing_directory_serves_synthetic_and_foreign_directory_is_refused(tmp_path):
    train, _ = dds.create_dataloader(str(tmp_path / "absent"), num_scans=10, batch_size=2, num_workers=0, cutout_kwargs=CUTOUT,
                                     num_samples=4)
    assert isinstance(train.dataset, dds.SyntheticDROWDataset)
    (tmp_path / "other").mkdir()
    (tmp_path / "other" / "readme.txt").write_text("not DROW")
    with pytest.raises(FileNotFoundError):
        dds.create_dataloader(str(tmp_path / "other"), num_scans=10, batch_size=2, num_workers=0)


def test_host_helpers_small_cases():
    phi = u.get_laser_phi()
    scan = np.full(450, 5.0, dtype=np.float32)
    assert np.array_equal(u.closest_detection(scan, phi, [], []), np.zeros(450, dtype=int))
    which = u.closest_detection(scan, phi, [(5.0, phi[100]), (5.0, phi[101])], [0.3, 0.3])
    assert which[100] == 1 and which[101] == 2 and which[0] == 0
    flow = np.random.RandomState(0).randn(450, 2)
    assert np.allclose(u.canonical_to_global_flow(u.global_to_canonical_flow(flow, phi), phi), flow, atol=1e-12)
    xy = np.stack(u.rphi_to_xy(scan, phi), axis=1)
    same = u.get_displacement_from_odometry(xy, np.array([1.0, 2.0, 0.3], np.float32), np.array([1.0, 2.0, 0.3], np.float32))
    assert np.abs(same).max() < 1e-6


@pytest.mark.reference
def test_dataset_matches_reference_dataset(drow_dir, capsys):
    """Same files -> the reference's DROWDataset2 and ours yield the same samples (every key but the CPU cutout `input`)."""
    from oracle import ref_shim

    ref = ref_shim.load_dataset_module()
    ref_utils, _ = ref_shim.load()
    theirs = ref.DROWDataset2(drow_dir, split="train", num_scans=10, cutout_kwargs=CUTOUT, pedestrian_only=False)
    ours = dds.DROWDataset(drow_dir, split="train", num_scans=10, cutout_kwargs=CUTOUT, pedestrian_only=False)
    assert len(theirs) == len(ours) > 0
    by_name = {}
    for i in range(len(theirs)):
        t = theirs[i]
        by_name[(os.path.basename(t["seq_name"]), int(t["dets_ns"]))] = t
    for i in range(len(ours)):
        o = ours[i]
        t = by_name[(os.path.basename(o["seq_name"]), int(o["dets_ns"]))]
        for k in ("scans", "target_cls", "target_reg", "target_flow", "exclude_mask", "odom1", "phi_grid"):
            assert np.array_equal(np.asarray(o[k]), np.asarray(t[k])), k
        assert o["scans"].dtype == t["scans"].dtype and o["target_reg"].dtype == t["target_reg"].dtype
        assert list(o["scans_ns"]) == list(t["scans_ns"]) and o["dets_wp"] == t["dets_wp"]
        # the reference's `input` is its CPU cutout of the same scans: what the device cutout replaces
        assert t["input"].shape == (450, 11, 56)
    # the helpers themselves, on random geometry
    rs = np.random.RandomState(3)
    phi = u.get_laser_phi()
    for _ in range(5):
        scan = rs.uniform(0.5, 12, 450).astype(np.float32)
        dets = [(float(rs.uniform(1, 10)), float(rs.uniform(phi[0], phi[-1]))) for _ in range(6)]
        a = u.get_regression_target(scan, phi, dets[:2], dets[2:3], dets[3:])
        b = ref_utils.get_regression_target(scan, phi, dets[:2], dets[2:3], dets[3:])
        assert np.array_equal(a[0], b[0]) and np.array_equal(a[1], b[1])
        xy = np.array(u.rphi_to_xy(scan, phi)).T
        o0, o1 = rs.randn(3).astype(np.float32), rs.randn(3).astype(np.float32)
        assert np.array_equal(u.get_displacement_from_odometry(xy, o0, o1), ref_utils.get_displacement_from_odometry(xy, o0, o1))
        f = rs.randn(450, 2)
        assert np.array_equal(u.global_to_canonical_flow(f, phi), ref_utils.global_to_canonical_flow(f, phi))


def test_flow_dataset_reads_difodom_and_flow_files(drow_dir):
    from planar_optical_flow_b200 import dataset as fds

    rs = np.random.RandomState(9)
    for name, n in (("rec_a", 40), ("rec_b", 30), ("rec_static", 40)):
        stem = os.path.join(drow_dir, "train", name)
        inc = np.column_stack([10.0 + 0.08 * np.arange(n), rs.normal(0, 0.02, (n, 3))])
        np.savetxt(stem + ".difodom", inc, delimiter=",", fmt="%.6f")
        np.savetxt(stem + ".flow", rs.normal(0, 0.05, (n, 900)), delimiter=",", fmt="%.5f")
    ds = fds.FlowDataset(drow_dir, split="train")
    assert isinstance(ds, fds.DROWFlowDataset) and len(ds) == 110
    s = ds[39]                                                         # last scan of rec_a: paired with itself
    assert s["scan_pair"][0].shape == (450, 2) and s["flow_target"].shape == (450, 2)
    rot_back = s["scan_pair"][1] - np.matmul(s["odom"][:-1], np.array([[np.cos(ds.scan_dir[39]), -np.sin(ds.scan_dir[39])],
                                                                      [np.sin(ds.scan_dir[39]), np.cos(ds.scan_dir[39])]], np.float32).T)
    assert np.allclose(np.linalg.norm(rot_back, axis=1), np.linalg.norm(s["scan_pair"][0], axis=1), atol=1e-4)
    loader = fds.create_flow_dataloader(8, split="train", data_path=drow_dir)
    b = next(iter(loader))
    assert b["scan_pair"].shape == (8, 2, 450, 2) and b["flow_target"].shape == (8, 450, 2)
