"""Training / evaluation closures for the DR-SPAAM entry points (reference: src/utils/eval_utils.py).

`model_fn_obj_det` is the loss that matches SpatialDROW's outputs in the reference
(eval_utils.py:31-88: BCE on sigmoid scores + masked vote regression); the only change is where the
cutouts come from: the batch carries raw `scans` and the cutouts are generated on the device by
the cutout kernel (one launch per batch), not in DataLoader workers.
`eval_dr_spaam` streams test sequences through the model (memory carried, as the reference's
`testing=True` branch) and runs the NMS kernel on every scan.
"""
import numpy as np
import torch
import torch.nn.functional as F

from . import ops


def batch_cutouts(batch, cutout_kwargs, device, fast=False):
    """Raw ranges [B, S, N] (NumPy) -> device cutouts [B, N, S, P] (dataset_dr_spaam.py:445 on the GPU)."""
    scans = batch["scans"]
    if isinstance(scans, torch.Tensor):                 # pinned staging tensor of the loader: asynchronous copy
        scans = scans.to(device, dtype=torch.float32, non_blocking=True).contiguous()
    else:
        scans = torch.from_numpy(np.ascontiguousarray(scans, dtype=np.float32)).to(device, non_blocking=True)
    if "area_mode" not in cutout_kwargs:          # legacy configs (dataset_dr_spaam.py:440-443): integer window + cv2.resize
        phi_h = np.asarray(batch["scan_phi"])
        incre = phi_h[1] - phi_h[0]
        return ops.cutout_original(scans, float(incre), angle_incre_is_f32=phi_h.dtype == np.float32, **cutout_kwargs)
    phi = torch.from_numpy(np.ascontiguousarray(batch["scan_phi"])).to(device)
    return ops.cutout(scans, phi, fast=fast, **cutout_kwargs)


def _to_device(a, device):
    """NumPy array or (possibly pinned) CPU tensor -> device tensor, asynchronously when the source is pinned."""
    t = a if isinstance(a, torch.Tensor) else torch.from_numpy(np.ascontiguousarray(a))
    return t.to(device, non_blocking=True)


def model_fn_obj_det(model, batch, rtn_result=False, cutout_kwargs=None):
    """The reference's `model_fn_obj_det(model, batch, rtn_result=False)` (eval_utils.py:31-88).

    A batch in the reference's own format (`input` = cutouts made by the dataset, `target_flow_cls/_reg`) is taken as
    it is; a batch of this package's loaders carries raw `scans` and the cutouts are made on the device with
    `cutout_kwargs` (or the batch's own `cutout_kwargs` entry)."""
    device = next(model.parameters()).device
    if "input" in batch:
        net_input = _to_device(batch["input"], device).float()
    else:
        kw = cutout_kwargs if cutout_kwargs is not None else batch.get("cutout_kwargs")
        if kw is None:
            raise KeyError("the batch has neither `input` (cutouts) nor `cutout_kwargs` to make them from `scans`")
        net_input = batch_cutouts(batch, kw, device)
    model_rtn = model(net_input)
    pred_cls, pred_reg = model_rtn[0], model_rtn[1]
    cls_key, reg_key = ("target_flow_cls", "target_flow_reg") if "target_flow_cls" in batch else ("target_cls", "target_reg")
    target_cls = _to_device(batch[cls_key], device).long()
    target_reg = _to_device(batch[reg_key], device).float()
    n_batch, n_pts = target_cls.shape[:2]
    target_cls = target_cls.view(n_batch * n_pts)
    pred_cls = pred_cls.view(n_batch * n_pts, -1)
    core = model.module if hasattr(model, "module") else model
    if pred_cls.shape[1] == 1:                                   # eval_utils.py:55-58
        cls_loss = core.cls_loss(torch.sigmoid(pred_cls.squeeze(-1)), target_cls.float(), reduction="mean")
    else:
        cls_loss = core.cls_loss(pred_cls, target_cls, reduction="mean")
    total, tb = cls_loss, {"cls_loss": cls_loss.item()}
    fg = target_cls.ne(0)
    tb["fg_ratio"] = float(fg.sum().item()) / (n_batch * n_pts)
    if tb["fg_ratio"] > 0.0:                                     # eval_utils.py:69-76
        reg = F.mse_loss(pred_reg.view(n_batch * n_pts, -1)[fg], target_reg.view(n_batch * n_pts, -1)[fg],
                         reduction="none")
        reg_loss = torch.sqrt(torch.sum(reg, dim=1)).mean()
        total = total + reg_loss
        tb["reg_loss"] = reg_loss.item()
    rtn = {}
    if rtn_result:
        rtn = {"pred_reg": pred_reg.view(n_batch, n_pts, -1), "pred_cls": pred_cls.view(n_batch, n_pts, -1)}
    return total, tb, rtn


def make_model_fn_obj_det(cutout_kwargs):
    """`model_fn_obj_det` bound to the config's `cutout_kwargs` (what the Trainer calls as model_fn(model, batch))."""
    def bound(model, batch, rtn_result=False):
        return model_fn_obj_det(model, batch, rtn_result=rtn_result, cutout_kwargs=cutout_kwargs)

    return bound


# ------------------------------------------------------------------ scan-pair flow prototype (row N3)
def model_fn(model, batch):
    """Training closure of the prototype (eval_utils.py:10-29): scan pair -> flow, mean end-point error.

    The reference reads `batch["flow_target_flow"]` (its datasets emit `flow_target`) and unpacks three values
    from a loss that returns two; this is the version that runs."""
    device = next(model.parameters()).device
    pair = torch.from_numpy(np.ascontiguousarray(batch["scan_pair"], dtype=np.float32)).to(device, non_blocking=True)
    target = torch.from_numpy(np.ascontiguousarray(batch["flow_target"], dtype=np.float32)).to(device, non_blocking=True)
    core = model.module if hasattr(model, "module") else model
    loss, _ = core.loss_fn(model(pair[:, 0], pair[:, 1]), target)
    return loss


def loss_fn_eval(pred_flow, target_flow):
    """End-point error and average angular error per sample (eval_utils.py:129-134)."""
    epe = torch.mean(torch.norm(pred_flow - target_flow, dim=-1), dim=1)
    aae = torch.mean(torch.abs(torch.atan2(pred_flow[..., 0], pred_flow[..., 1]) -
                               torch.atan2(target_flow[..., 0], target_flow[..., 1])), dim=1) * 180 / np.pi
    return epe, aae


@torch.no_grad()
def model_fn_eval(model, eval_loader):
    """(mean EPE, mean AAE) over a loader of scan pairs (eval_utils.py:136-155, for the prototype's batches)."""
    device = next(model.parameters()).device
    model.eval()
    epe_sum = aae_sum = 0.0
    for batch in eval_loader:
        pair = torch.from_numpy(np.ascontiguousarray(batch["scan_pair"], dtype=np.float32)).to(device)
        target = torch.from_numpy(np.ascontiguousarray(batch["flow_target"], dtype=np.float32)).to(device)
        epe, aae = loss_fn_eval(model(pair[:, 0], pair[:, 1]), target)
        epe_sum += torch.mean(epe).item()
        aae_sum += torch.mean(aae).item()
    n = max(len(eval_loader), 1)
    return epe_sum / n, aae_sum / n


@torch.no_grad()
def eval_dr_spaam(model, test_loader, cfg, output_dir=None, max_sequences=None):
    """Stream every test sample's scans through the detector (memory carried) and run NMS on each."""
    device = next(model.parameters()).device
    model.eval()
    n_scans = n_det = 0
    results = []
    for i, batch in enumerate(test_loader):
        if max_sequences is not None and i >= max_sequences:
            break
        cutouts = batch_cutouts(batch, cfg["cutout_kwargs"], device)              # [1, N, S, P]
        scans = torch.from_numpy(np.ascontiguousarray(batch["scans"], dtype=np.float32)).to(device)
        phi = torch.from_numpy(np.ascontiguousarray(batch["scan_phi"])).to(device)
        tmpl = None
        for s in range(cutouts.shape[2]):
            pred_cls, pred_reg, tmpl, _ = model(cutouts[:, :, s:s + 1, :].contiguous(), testing=True, fea_template=tmpl)
            conf = torch.sigmoid(pred_cls[..., 0]).contiguous()
            res = ops.nms_centers(scans[:, s].contiguous(), phi, conf, pred_reg.contiguous())
            n_scans += 1
            n_det += int(res["n_keep"][0])
        k = int(res["n_keep"][0])
        results.append({"det_xy": res["det_xy"][0, :k].cpu().numpy(), "det_cls": res["det_cls"][0, :k].cpu().numpy(),
                        "instance_mask": res["instance_mask"][0].cpu().numpy()})
    summary = {"scans": n_scans, "detections": n_det, "detections_per_scan": n_det / max(n_scans, 1)}
    if output_dir is not None:
        import json
        import os

        with open(os.path.join(output_dir, "eval_summary.json"), "w") as f:
            json.dump(summary, f)
    return summary, results
