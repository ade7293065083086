"""The north_star entry points run end to end on the GPU with the reference's CLI."""
import os
import subprocess
import sys

import pytest
import torch

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _run(args, tmp_path):
    env = dict(os.environ, PYTHONPATH=ROOT)
    p = subprocess.run([sys.executable] + args, cwd=ROOT, env=env, capture_output=True, text=True, timeout=900)
    assert p.returncode == 0, p.stdout[-2000:] + p.stderr[-4000:]
    return p.stdout


def test_train_then_eval_scripts(tmp_path):
    out = str(tmp_path)
    log = _run(["bin/train_dr_spaam.py", "--cfg", "config/dr_spaam.yaml", "--out", out, "--max-iters", "3",
                "--num-samples", "32", "--data", "/nonexistent"], tmp_path)
    assert "final loss" in log
    # a checkpoint in the reference's format, written by hand here (the 3-iteration run stops before epoch end)
    from planar_optical_flow_b200 import train_utils as tu
    from planar_optical_flow_b200.model import SpatialDROW

    m = SpatialDROW(num_scans=10, num_pts=56, alpha=0.5, window_size=11, pedestrian_only=True)
    tu.save_checkpoint(tu.checkpoint_state(m, None, 1, 3), filename=os.path.join(out, "ckpt_e1"))
    log = _run(["bin/eval_dr_spaam.py", "--cfg", "config/dr_spaam.yaml", "--out", out, "--ckpt",
                os.path.join(out, "ckpt_e1.pth"), "--num-samples", "2", "--data", "/nonexistent"], tmp_path)
    assert "detections_per_scan" in log


def test_prototype_train_then_eval_scripts(tmp_path):
    out = str(tmp_path)
    log = _run(["bin/train_prototype.py", "--out", out, "--epochs", "1", "--batch-size", "20", "--num-samples", "60",
                "--ckpt-save-interval", "1", "--data", "/nonexistent"], tmp_path)
    assert "Analysis finished" in log and os.path.isfile(os.path.join(out, "ckpts", "ckpt_e1.pth"))
    log = _run(["bin/eval_prototype.py", "--ckpt", os.path.join(out, "ckpts", "ckpt_e1.pth"), "--num-samples", "40",
                "--batch-size", "20", "--data", "/nonexistent"], tmp_path)
    assert "epe" in log


def test_training_step_decreases_loss_on_fixed_batch():
    from planar_optical_flow_b200.dataset_dr_spaam import create_dataloader
    from planar_optical_flow_b200.eval_utils import make_model_fn_obj_det
    from planar_optical_flow_b200.model import SpatialDROW

    cfg = dict(fixed=True, centered=True, window_width=1.0, window_depth=0.5, num_cutout_pts=56, padding_val=29.99,
               area_mode=True)
    torch.manual_seed(0)
    loader, _ = create_dataloader("/nonexistent", 4, 4, 0, cutout_kwargs=cfg, pedestrian_only=True, num_samples=4)
    batch = next(iter(loader))
    model = SpatialDROW(num_scans=4, num_pts=56, alpha=0.5, window_size=11, pedestrian_only=True).cuda().train()
    fn = make_model_fn_obj_det(cfg)
    opt = torch.optim.Adam(model.parameters(), lr=1e-3)
    losses = []
    for _ in range(6):
        opt.zero_grad()
        loss, tb, _ = fn(model, batch)
        loss.backward()
        opt.step()
        losses.append(loss.item())
    assert all(torch.isfinite(torch.tensor(losses))) and losses[-1] < losses[0], losses


def test_train_and_eval_scripts_on_drow_format_files(tmp_path):
    """The same entry points on a DROWv2-format directory (recordings written in the .csv/.wc/.wa/.wp/.odom2 formats): the real
    reader, targets, pinned staging and device cutouts end to end."""
    from tests.test_drow_dataset import write_recording

    data = tmp_path / "DROWv2-data"
    for split, seed in (("train", 1), ("val", 3), ("test", 5)):          # config/dr_spaam.yaml trains with validation
        (data / split).mkdir(parents=True)
        write_recording(str(data / split / "rec_a"), seed=seed, n_scans=48)
        write_recording(str(data / split / "rec_b"), seed=seed + 1, n_scans=40)
    out = str(tmp_path / "out")
    log = _run(["bin/train_dr_spaam.py", "--cfg", "config/dr_spaam.yaml", "--out", out, "--max-iters", "2", "--data", str(data)], tmp_path)
    assert "valid files found" in log and "final loss" in log
    from planar_optical_flow_b200 import train_utils as tu
    from planar_optical_flow_b200.model import SpatialDROW

    m = SpatialDROW(num_scans=10, num_pts=56, alpha=0.5, window_size=11, pedestrian_only=True)
    tu.save_checkpoint(tu.checkpoint_state(m, None, 1, 2), filename=os.path.join(out, "ckpt_e1"))
    log = _run(["bin/eval_dr_spaam.py", "--cfg", "config/dr_spaam.yaml", "--out", out, "--ckpt", os.path.join(out, "ckpt_e1.pth"),
                "--data", str(data)], tmp_path)
    assert "detections_per_scan" in log
