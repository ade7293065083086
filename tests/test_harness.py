"""Host-side harness: reference import paths, config schema, dataset batches, checkpoints, schedulers."""
import os

import numpy as np
import torch
import yaml

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_reference_import_paths_resolve():
    import src.utils.train_utils as tu
    import src.utils.utils as u
    from src.depracted.model import DROW, SpatialDROW  # noqa: F401
    from src.depracted.model.dr_spaam import _SpatialAttention  # noqa: F401
    from src.utils.dataset_dr_spaam import create_dataloader, create_test_dataloader  # noqa: F401
    from src.utils.train_utils import Trainer, load_checkpoint  # noqa: F401

    assert callable(u.scans_to_cutout) and callable(u.nms_predicted_center) and callable(u.scans_to_cutout_torch)
    assert tu.lr_scheduler() == 0.01
    phi = u.get_laser_phi()
    assert phi.shape == (450,) and phi.dtype == np.float64 and abs(phi[1] - phi[0] - np.radians(0.5)) < 1e-12


def test_prototype_import_paths_and_batches():
    """Row N3: `from src.depracted.model import Prototype`, FlowDataset batches with the reference's keys."""
    import numpy as np

    from src.depracted.model import Prototype  # noqa: F401
    from src.utils.dataset import FlowDataset
    from src.utils.eval_utils import model_fn, model_fn_eval  # noqa: F401

    ds = FlowDataset(None, split="train", num_samples=6)
    batch = ds.collate_batch([ds[0], ds[1], ds[2]])
    assert batch["scan_pair"].shape == (3, 2, 450, 2) and batch["flow_target"].shape == (3, 450, 2)
    assert batch["scan_pair"].dtype == np.float32
    m = Prototype(in_channel=2)
    ref_keys = {"encoder_0", "encoder_1", "encoder_2", "decoder_1", "decoder_0", "flow_reg"}
    assert {k.split(".")[0] for k in m.state_dict()} == ref_keys
    assert m.decoder_1[0].weight.shape == (128, 11 + 128, 3) and m.flow_reg[0].weight.shape == (2, 130, 1)


def test_config_schema_matches_reference_keys():
    cfg = yaml.safe_load(open(os.path.join(ROOT, "config", "dr_spaam.yaml")))
    for k in ("tag", "epochs", "batch_size", "grad_norm_clip", "num_workers", "num_scans", "use_data_augumentation",
              "train_with_val", "use_polar_grid", "focal_loss_gamma", "pedestrian_only", "network", "similarity_kwargs",
              "cutout_kwargs", "polar_grid_kwargs"):
        assert k in cfg, k
    assert cfg["cutout_kwargs"] == dict(fixed=True, centered=True, window_width=1.0, window_depth=0.5,
                                        num_cutout_pts=56, padding_val=29.99, area_mode=True)
    assert cfg["similarity_kwargs"] == dict(alpha=0.5, window_size=11)
    assert cfg["num_scans"] == 10 and cfg["batch_size"] == 8 and cfg["network"] == "cutout_spatial"


def test_dataset_batches_carry_raw_scans():
    from planar_optical_flow_b200.dataset_dr_spaam import create_dataloader, create_test_dataloader

    train, val = create_dataloader("/nonexistent", num_scans=10, batch_size=4, num_workers=8, train_with_val=True,
                                   cutout_kwargs={}, pedestrian_only=True, num_samples=16)
    b = next(iter(train))
    assert b["scans"].shape == (4, 11, 450) and b["scans"].dtype == np.float32       # num_scans + current
    assert b["target_cls"].shape == (4, 450) and b["target_reg"].shape == (4, 450, 2)
    assert val is not None and len(train) == 4
    single, none = create_dataloader("/nonexistent", 10, 4, 0, train_with_val=False, num_samples=8)
    assert none is None
    test = create_test_dataloader("/nonexistent", num_scans=10, num_samples=3)
    assert len(test) == 3 and next(iter(test))["scans"].shape == (1, 11, 450)
    # determinism: same index -> same sample
    assert np.array_equal(train.dataset[3]["scans"], train.dataset[3]["scans"])


def test_checkpoint_round_trip_uses_reference_keys(tmp_path):
    from planar_optical_flow_b200 import train_utils as tu
    from planar_optical_flow_b200.model import SpatialDROW

    m = SpatialDROW(num_scans=10, num_pts=56, alpha=0.5, window_size=11, pedestrian_only=True)
    opt = torch.optim.Adam(m.parameters(), lr=tu.lr_scheduler())
    state = tu.checkpoint_state(m, opt, epoch=3, it=77)
    assert sorted(state) == ["epoch", "it", "model_state", "optimizer_state"]
    assert "gate.conv.0.weight" in state["model_state"] and "conv_block_1.0.0.weight" in state["model_state"]
    tu.save_checkpoint(state, filename=str(tmp_path / "ckpt_e3"))
    m2 = SpatialDROW(num_scans=10, num_pts=56, alpha=0.5, window_size=11, pedestrian_only=True)
    it, epoch = tu.load_checkpoint(model=m2, optimizer=None, filename=str(tmp_path / "ckpt_e3.pth"))
    assert (it, epoch) == (77, 3)
    for (k, a), (_, b) in zip(m.state_dict().items(), m2.state_dict().items()):
        assert torch.equal(a, b), k
    try:
        tu.load_checkpoint(model=m2, filename=str(tmp_path / "missing.pth"))
        raise AssertionError("expected FileNotFoundError")
    except FileNotFoundError:
        pass


def test_schedulers_and_trainer_on_a_cpu_toy():
    from planar_optical_flow_b200 import train_utils as tu

    net = torch.nn.Linear(3, 1)
    opt = torch.optim.SGD(net.parameters(), lr=0.1)
    sch = tu.LucasScheduler(opt, e0=1, v0=0.1, e1=3, v1=0.001)
    sch.step(0.5)
    assert sch.get_lr() == 0.1
    sch.step(2)
    assert abs(sch.get_lr() - 0.01) < 1e-12
    sch.step(10)
    assert sch.get_lr() == 0.001
    data = [{"x": torch.randn(8, 3), "y": torch.randn(8, 1)} for _ in range(4)]
    trainer = tu.Trainer(net, lambda m, b: torch.nn.functional.mse_loss(m(b["x"]), b["y"]), opt, ckpt_dir="/tmp",
                         lr_scheduler=None, grad_norm_clip=0.0)
    last = trainer.train(num_epochs=100, train_loader=data, max_iters=3)
    assert trainer._it == 3 and np.isfinite(last)
