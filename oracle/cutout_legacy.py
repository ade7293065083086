"""CPU oracle for the reference's legacy preprocessing.  TEST INFRASTRUCTURE ONLY.

  * `scans_to_cutout_original`  <- /root/reference/src/utils/utils.py:423-489 (integer window, cv2.resize
                                   with INTER_AREA when shrinking / INTER_LINEAR otherwise)
  * `scans_to_polar_grid`       <- utils.py:492-531
  * `resize_column`             <- what `cv2.resize(column, (1, P), interpolation=...)` computes for a
                                   float32 column (OpenCV imgproc/resize.cpp: `resizeGeneric_` with
                                   VResizeLinear / `ResizeAreaFast_` / `computeResizeAreaTab` + `ResizeArea_`)

OpenCV is a third-party dependency of the reference (opencv-python, unpinned in requirements.txt; 4.13.0
in this image).  Its SIMD paths may or may not contract multiply-adds, so `resize_column` is pinned to
cv2 itself to 1e-6 relative, not bit for bit (tests/test_oracle_vs_reference.py); everything around the
resize follows NumPy's arithmetic exactly.
"""
import math

import numpy as np

F32 = np.float32


def _area_tab(ssize, dsize, scale):
    """computeResizeAreaTab (resize.cpp): list of (dst, src, alpha float32)."""
    tab = []
    for dx in range(dsize):
        fsx1 = dx * scale
        fsx2 = fsx1 + scale
        cell = min(scale, ssize - fsx1)
        sx1, sx2 = math.ceil(fsx1), math.floor(fsx2)
        sx2 = min(sx2, ssize - 1)
        sx1 = min(sx1, sx2)
        if sx1 - fsx1 > 1e-3:
            tab.append((dx, sx1 - 1, F32((sx1 - fsx1) / cell)))
        for sx in range(sx1, sx2):
            tab.append((dx, sx, F32(1.0 / cell)))
        if fsx2 - sx2 > 1e-3:
            tab.append((dx, sx2, F32(min(min(fsx2 - sx2, 1.0), cell) / cell)))
    return tab


def resize_column(col, P, area):
    """float32 column [len] -> [P]; `area` selects INTER_AREA, else INTER_LINEAR."""
    col = np.asarray(col, dtype=F32)
    n = len(col)
    out = np.empty(P, dtype=F32)
    scale = 1.0 / (float(P) / float(n))                 # resize(): inv_scale = dsize / ssize, scale = 1 / inv_scale
    if area and n > P:
        iscale = int(round(scale)) if abs(scale - round(scale)) < 2.220446049250313e-16 else None
        if iscale is not None and abs(scale - iscale) < 2.220446049250313e-16:
            inv = F32(1.0 / iscale)                      # ResizeAreaFast_: sum of the cell times 1/area
            for dy in range(P):
                acc = F32(0)
                for k in range(iscale):
                    acc = F32(acc + col[dy * iscale + k])
                out[dy] = F32(acc * inv)
            return out
        acc, prev = F32(0), -1                           # ResizeArea_: sum = beta*v, then sum += beta*v, source order
        for dy, sy, beta in _area_tab(n, P, scale):
            term = F32(beta * col[sy])
            if dy != prev:
                if prev >= 0:
                    out[prev] = acc
                acc, prev = term, dy
            else:
                acc = F32(acc + term)
        out[prev] = acc
        return out
    for dy in range(P):                                  # resizeGeneric_, linear: fy = (dy + .5) * scale - .5
        fy = F32((dy + 0.5) * scale - 0.5)
        sy = math.floor(fy)
        fy = F32(fy - F32(sy))
        v0 = col[min(max(sy, 0), n - 1)]
        v1 = col[min(max(sy + 1, 0), n - 1)]
        out[dy] = F32(F32(v0 * F32(F32(1.0) - fy)) + F32(v1 * fy))
    return out


def window_bounds(pt_r, pt_idx, angle_incre, window_width):
    """(start_idx, end_idx) of the integer window (utils.py:449-453), with NumPy's scalar arithmetic."""
    half_alpha = float(np.arctan(0.5 * window_width / max(pt_r, 0.01)))
    start = int(round(pt_idx - half_alpha / angle_incre))
    end = int(round(pt_idx + half_alpha / angle_incre))
    return start, end


def window_margins(scans, angle_incre, fixed, window_width):
    """[S, N] distance of the two window ends from a rounding boundary (x.5): a device arctangent that is
    1-2 ulp from NumPy's can only move a window where this is ~1e-6."""
    num_scans, num_pts = scans.shape
    out = np.empty((num_scans, num_pts))
    for s in range(num_scans):
        for i in range(num_pts):
            pt_r = scans[s, i] if fixed else scans[-1, i]
            q = float(np.arctan(0.5 * window_width / max(pt_r, 0.01))) / angle_incre
            out[s, i] = min(abs(((i - q) % 1.0) - 0.5), abs(((i + q) % 1.0) - 0.5))
    return out


def scans_to_cutout_original(scans, angle_incre, fixed=True, centered=True, pt_inds=None, window_width=1.66, window_depth=1.0,
                             num_cutout_pts=48, padding_val=29.99):
    num_scans, num_pts = scans.shape
    if pt_inds is None:
        pt_inds = range(num_pts)
    padded = np.pad(scans, ((0, 0), (0, 1)), mode="constant", constant_values=padding_val)          # :440-442
    out = np.empty((num_pts, num_scans, num_cutout_pts), dtype=F32)
    for s in range(num_scans):
        for i in pt_inds:
            pt_r = scans[s, i] if fixed else scans[-1, i]                                            # :448
            start, end = window_bounds(pt_r, i, angle_incre, window_width)
            inds = np.clip(np.arange(start, end + 1), -1, num_pts)                                   # :455-456
            col = padded[s, inds]                                                                    # :461
            v = resize_column(col, num_cutout_pts, area=num_cutout_pts < len(inds))                  # :464-471
            v = np.clip(v, pt_r - window_depth, pt_r + window_depth)                                 # :474-476
            if centered:
                v = v - pt_r                                                                         # :485
                v = v / window_depth                                                                 # :486
            out[i, s, :] = v
    return out


def scans_to_polar_grid(scans, min_range=0.0, max_range=30.0, range_bin_size=1.0, tsdf_clip=1.0, normalize=True):
    """utils.py:492-531."""
    num_scans, num_pts = scans.shape
    num_range = int((max_range - min_range) / range_bin_size) + 1
    mag, mid = max_range - min_range, 0.5 * (max_range - min_range)
    grid = np.empty((num_scans, num_range, num_pts), dtype=F32)
    scans = np.clip(scans, min_range, max_range)
    inds = ((scans - min_range) / range_bin_size).astype(np.int32)
    for s in range(num_scans):
        for i in range(num_pts):
            ind, val = inds[s, i], scans[s, i]
            if tsdf_clip > 0.0:
                tsdf = np.arange(0 - ind, num_range - ind, step=1).astype(F32) * range_bin_size
                tsdf = np.clip(tsdf, -tsdf_clip, tsdf_clip)
            else:
                tsdf = np.zeros(num_range, dtype=F32)
            if normalize:
                val = (val - mid) / mag * 2.0
                tsdf = tsdf / mag * 2.0
            tsdf[ind] = val
            grid[s, :, i] = tsdf
    return grid
