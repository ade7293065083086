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
    scans = torch.from_numpy(np.ascontiguousarray(batch["scans"], dtype=np.float32)).to(device, non_blocking=True)
    if "area_mode" not in cutout_kwargs:          # legacy configs (dataset_dr_spaam.py:440-443): integer window + cv2.resize
        phi_h = np.asarray(batch["scan_phi"])
        incre = phi_h[1] - phi_h[0]
        return ops.cutout_original(scans, float(incre), angle_incre_is_f32=phi_h.dtype == np.float32, **cutout_kwargs)
    phi = torch.from_numpy(np.ascontiguousarray(batch["scan_phi"])).to(device)
    return ops.cutout(scans, phi, fast=fast, **cutout_kwargs)


def make_model_fn_obj_det(cutout_kwargs):
    def model_fn_obj_det(model, batch, rtn_result=False):
        device = next(model.parameters()).device
        net_input = batch_cutouts(batch, cutout_kwargs, device)
        model_rtn = model(net_input)
        pred_cls, pred_reg = model_rtn[0], model_rtn[1]
        target_cls = torch.from_numpy(batch["target_cls"]).to(device, non_blocking=True).long()
        target_reg = torch.from_numpy(batch["target_reg"]).to(device, non_blocking=True).float()
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

    return model_fn_obj_det


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
