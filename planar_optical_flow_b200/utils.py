"""Drop-in replacements for the hot-path functions of the reference's
`src/utils/utils.py`, same names, signatures, defaults, return types and error
behaviour — executed by the sm_100a kernels of libpof.so.

    scans_to_cutout         utils.py:259-334   NumPy in / NumPy out (host buffers; H2D + D2H inside)
    scans_to_cutout_torch   utils.py:337-420   torch in / torch out (stays on the tensor's device)
    nms_predicted_center    utils.py:535-571   NumPy in / NumPy out
    get_laser_phi, rphi_to_xy, canonical_to_global, ...   small host helpers the callers use

There is no CPU path: without a CUDA device these raise RuntimeError.
"""
import numpy as np
import torch

from . import ops


def _device():
    if not torch.cuda.is_available():
        raise RuntimeError("planar_optical_flow_b200 needs a CUDA device (B200, sm_100a); there is no CPU fallback")
    return torch.device("cuda", torch.cuda.current_device())


# ------------------------------------------------------------------ small host helpers
def get_laser_phi(angle_inc=np.radians(0.5), num_pts=450):
    """Beam angles of the DROW laser (utils.py:25-29)."""
    fov = (num_pts - 1) * angle_inc
    return np.linspace(-fov * 0.5, fov * 0.5, num_pts)


def rphi_to_xy(r, phi):
    return r * np.cos(phi), r * np.sin(phi)


def xy_to_rphi(x, y):
    return np.hypot(x, y), np.arctan2(y, x)


def scan_to_xy(scan, phi=None):
    return rphi_to_xy(scan, get_laser_phi() if phi is None else phi)


def canonical_to_global(scan_r, scan_phi, dx, dy):
    """Canonical vote -> global polar (utils.py:109-116)."""
    tmp_y = scan_r + dy
    tmp_phi = np.arctan2(dx, tmp_y)
    return tmp_y / np.cos(tmp_phi), tmp_phi + scan_phi


def global_to_canonical(scan_r, scan_phi, dets_r, dets_phi):
    """utils.py:55-59."""
    dx = np.sin(dets_phi - scan_phi) * dets_r
    dy = np.cos(dets_phi - scan_phi) * dets_r - scan_r
    return dx, dy


def global_to_canonical_flow(flow, scan_phi):
    """Per-point rotation of xy flow vectors into each beam's canonical frame (utils.py:62-76):
    rows (cos, -sin) and (sin, cos) of the beam angle applied to (fx, fy)."""
    s, c = np.sin(scan_phi), np.cos(scan_phi)
    rot = np.stack([np.stack([c, -s], axis=1), np.stack([s, c], axis=1)], axis=1)      # [N, 2, 2]
    return np.einsum("ijk,ik->ij", rot, flow)


def canonical_to_global_flow(flow_canonical, scan_phi):
    """Inverse of `global_to_canonical_flow` (utils.py:79-91)."""
    s, c = np.sin(scan_phi), np.cos(scan_phi)
    rot = np.stack([np.stack([c, s], axis=1), np.stack([-s, c], axis=1)], axis=1)
    return np.einsum("ijk,ik->ij", rot, flow_canonical)


def closest_detection(scan, scan_phi, dets, radii):
    """1-based index of the closest detection whose radius contains each scan point, 0 = none (utils.py:230-256).
    `dets`: (r, phi) pairs; distances are Euclidean in the scanner's xy frame, the radius is subtracted before the
    arg-min over (0 = outside everything, d_1 - radius_1, ...)."""
    if len(dets) == 0:
        return np.zeros_like(scan, dtype=int)
    if len(dets) != len(radii):
        raise AssertionError("Need to give a radius for each detection!")
    sx, sy = rphi_to_xy(np.asarray(scan), np.asarray(scan_phi))
    pts = np.stack([sx, sy], axis=1).astype(np.float64)
    d = np.asarray(dets, dtype=np.float64).reshape(-1, 2)
    centres = np.stack(rphi_to_xy(d[:, 0], d[:, 1]), axis=1)
    diff = pts[:, None, :] - centres[None, :, :]
    dists = np.sqrt((diff * diff).sum(axis=2)) - np.asarray(radii, dtype=np.float64)[None, :]
    return np.argmin(np.hstack([np.zeros((len(pts), 1)), dists]), axis=1)


def get_regression_target(scan, scan_phi, wcs, was, wps, radius_wc=0.6, radius_wa=0.4, radius_wp=0.35,
                          label_wc=1, label_wa=2, label_wp=3, pedestrian_only=False):
    """Per-point class label and canonical (dx, dy) vote to the centre of the closest annotated object
    (utils.py:143-181): `target_cls [N] int64`, `target_reg [N, 2] float32`."""
    n = len(scan)
    target_cls = np.zeros(n, dtype=np.int64)
    target_reg = np.zeros((n, 2), dtype=np.float32)
    if pedestrian_only:
        all_dets, radii, labels = list(wps), [radius_wp] * len(wps), [0] + [1] * len(wps)
    else:
        all_dets = list(wcs) + list(was) + list(wps)
        radii = [radius_wc] * len(wcs) + [radius_wa] * len(was) + [radius_wp] * len(wps)
        labels = [0] + [label_wc] * len(wcs) + [label_wa] * len(was) + [label_wp] * len(wps)
    which = closest_detection(scan, scan_phi, all_dets, radii)
    hit = which > 0
    if hit.any():
        d = np.asarray(all_dets, dtype=np.float64).reshape(-1, 2)[which[hit] - 1]
        target_cls[hit] = np.asarray(labels, dtype=np.int64)[which[hit]]
        dx, dy = global_to_canonical(np.asarray(scan)[hit], np.asarray(scan_phi)[hit], d[:, 0], d[:, 1])
        target_reg[hit, 0], target_reg[hit, 1] = dx, dy
    return target_cls, target_reg


def get_displacement_from_odometry(scan1_xy, odom0, odom1):
    """Apparent displacement of stationary points between two scanner poses (x, y, phi), expressed in the current
    scanner frame (utils.py:639-662): p - R0^T (R1 p + T1 - T0)."""
    def rot(phi):
        c, s = np.cos(phi), np.sin(phi)
        return np.array([[c, -s], [s, c]], dtype=np.float32)

    r0, r1 = rot(odom0[2]), rot(odom1[2])
    m = np.eye(2) - np.matmul(r0.T, r1)
    shift = (odom1[:2] - odom0[:2]).reshape(2, 1)
    return np.matmul(scan1_xy, m.T) - np.matmul(r0.T, shift).reshape(1, 2)


def data_augmentation(sample_dict, rng=np.random):
    """Random left-right flip of all scans of a sample and of the x component of its votes (utils.py:126-140)."""
    scans, target_reg = sample_dict["scans"], sample_dict["target_reg"]
    if rng.rand() < 0.5:
        scans = scans[:, ::-1]
        target_reg[:, 0] = -target_reg[:, 0]
    sample_dict.update({"target_reg": target_reg, "scans": scans})
    return sample_dict


# ------------------------------------------------------------------ cutout
def _phi_tensor(scan_phi, device):
    phi = np.ascontiguousarray(scan_phi)
    if phi.dtype not in (np.float32, np.float64):
        phi = phi.astype(np.float64)
    return torch.from_numpy(phi).to(device)


def scans_to_cutout(scans, scan_phi, stride=1, centered=True, fixed=False, window_width=1.66,
                    window_depth=1.0, num_cutout_pts=48, padding_val=29.99, area_mode=False):
    """`scans [S, N]`, `scan_phi [N]` (NumPy)  ->  cutouts `[ceil(N/stride), S, P]` float32 (NumPy).

    Same contract as utils.py:259-334.  float32 scans follow the reference's
    arithmetic exactly (see csrc/pof_cutout.cu); other scan dtypes are converted
    to float32 first (every reference caller on the DR-SPAAM path feeds float32,
    src/utils/dataset_dr_spaam.py:477).
    """
    scans = np.asarray(scans)
    if scans.ndim != 2:
        raise ValueError("scans must be [num_scans, num_pts]")
    dev = _device()
    s = torch.from_numpy(np.ascontiguousarray(scans, dtype=np.float32)).to(dev).unsqueeze(0)
    out = ops.cutout(s, _phi_tensor(scan_phi, dev), stride=stride, centered=centered, fixed=fixed,
                     window_width=window_width, window_depth=window_depth, num_cutout_pts=num_cutout_pts,
                     padding_val=padding_val, area_mode=area_mode)
    return out[0].cpu().numpy()


def scans_to_cutout_torch(scans, scan_phi, stride=1, centered=True, fixed=False, window_width=1.66,
                          window_depth=1.0, num_cutout_pts=48, padding_val=29.99, area_mode=False):
    """torch in / torch out variant (utils.py:337-420 signature).

    The reference's torch port is all-float32 and disagrees with its own NumPy
    function by up to 0.75 on discontinuous scans; results here
    follow the NumPy function, which is the one every caller uses.
    """
    if not scans.is_cuda:
        raise RuntimeError("scans_to_cutout_torch needs CUDA tensors (no CPU path)")
    phi = scan_phi if scan_phi.dtype in (torch.float32, torch.float64) else scan_phi.double()
    out = ops.cutout(scans.float().contiguous().unsqueeze(0), phi.to(scans.device).contiguous(), stride=stride,
                     centered=centered, fixed=fixed, window_width=window_width, window_depth=window_depth,
                     num_cutout_pts=num_cutout_pts, padding_val=padding_val, area_mode=area_mode)
    return out[0]


def scans_to_cutout_batch(scans, scan_phi, **cutout_kwargs):
    """`[B, S, N]` NumPy or CUDA tensor -> `[B, N, S, P]` CUDA tensor: one launch for a whole
    batch of samples (what dataset_dr_spaam.py:445 + collate :464-468 produce one by one)."""
    dev = _device()
    if isinstance(scans, np.ndarray):
        scans = torch.from_numpy(np.ascontiguousarray(scans, dtype=np.float32)).to(dev)
    phi = scan_phi if isinstance(scan_phi, torch.Tensor) else _phi_tensor(scan_phi, dev)
    return ops.cutout(scans.float().contiguous(), phi.to(scans.device), **cutout_kwargs)


def scans_to_cutout_original(scans, angle_incre, fixed=True, centered=True, pt_inds=None, window_width=1.66, window_depth=1.0,
                             num_cutout_pts=48, padding_val=29.99):
    """The legacy cutout (utils.py:423-489; configs without `area_mode`): `scans [S, N]` -> `[N, S, P]` float32.

    `pt_inds` restricts the reference's loop to some points and leaves the other rows of its np.empty
    output uninitialised; here every row is computed."""
    scans = np.asarray(scans)
    if scans.ndim != 2:
        raise ValueError("scans must be [num_scans, num_pts]")
    dev = _device()
    s = torch.from_numpy(np.ascontiguousarray(scans, dtype=np.float32)).to(dev).unsqueeze(0)
    is_f32 = isinstance(angle_incre, np.floating) and np.dtype(type(angle_incre)) == np.float32
    out = ops.cutout_original(s, float(angle_incre), angle_incre_is_f32=is_f32, fixed=fixed, centered=centered,
                              window_width=window_width, window_depth=window_depth, num_cutout_pts=num_cutout_pts,
                              padding_val=padding_val)
    return out[0].cpu().numpy()


def scans_to_polar_grid(scans, min_range=0.0, max_range=30.0, range_bin_size=1.0, tsdf_clip=1.0, normalize=True):
    """utils.py:492-531: `scans [S, N]` -> polar grid `[S, R, N]` float32."""
    scans = np.asarray(scans)
    dev = _device()
    s = torch.from_numpy(np.ascontiguousarray(scans, dtype=np.float32)).to(dev)
    return ops.polar_grid(s, min_range, max_range, range_bin_size, tsdf_clip, normalize).cpu().numpy()


# ------------------------------------------------------------------ NMS
def nms_predicted_center(scan_grid, phi_grid, pred_cls, pred_reg, min_dist=0.5):
    """Returns `(det_xys [K,2], det_cls [K,1], instance_mask [N] int32)` as utils.py:535-571.

    dtypes follow NumPy promotion of the inputs exactly as the reference's
    expressions do (float32 scans with a float64 angle grid give float64 xy).
    """
    assert pred_cls.shape[1] == 1                                   # utils.py:536
    dev = _device()
    scan = np.ascontiguousarray(scan_grid)
    phi = np.ascontiguousarray(phi_grid)
    if scan.dtype not in (np.float32, np.float64):
        scan = scan.astype(np.float64)
    if phi.dtype not in (np.float32, np.float64):
        phi = phi.astype(np.float64)
    n = scan.shape[0]
    cls = torch.from_numpy(np.ascontiguousarray(pred_cls[:, 0], dtype=np.float32)).to(dev).view(1, n)
    reg = torch.from_numpy(np.ascontiguousarray(pred_reg, dtype=np.float32)).to(dev).view(1, n, 2)
    res = ops.nms_centers(torch.from_numpy(scan).to(dev).view(1, n), torch.from_numpy(phi).to(dev), cls, reg,
                          min_dist=min_dist)
    k = int(res["n_keep"][0].item())
    xy_dtype = np.result_type(scan.dtype, phi.dtype)
    det_xys = res["det_xy"][0, :k].cpu().numpy().astype(xy_dtype, copy=False)
    det_cls = res["det_cls"][0, :k].cpu().numpy().reshape(k, 1).astype(pred_cls.dtype, copy=False)
    instance_mask = res["instance_mask"][0].cpu().numpy()
    return det_xys, det_cls, instance_mask
