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
