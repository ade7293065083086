"""CPU oracle for the predicted-centre vote / group / NMS.  TEST INFRASTRUCTURE ONLY.

Restates `nms_predicted_center` (/root/reference/src/utils/utils.py:535-571)
with its helpers `canonical_to_global` (:109-116) and `rphi_to_xy` (:47-48).
Only `tests/`, `__graft_entry__.smoke()` and `bench.py`'s CPU-baseline legs may
import it.  Pinned bit-for-bit against the imported reference by
`tests/test_oracle_vs_reference.py` and the `tests/golden/nms_*.npz` fixtures.

Two forms are given:
  * `nms_predicted_center`  — the reference's algorithm as it runs on the CPU
    (dense N x N distance matrix, Python sweep); this is the one timed as the
    CPU baseline;
  * `nms_sweep_spec`        — the equivalent "one serial sweep over a boolean
    adjacency" statement the CUDA kernel implements (SURVEY.md §8 a-bis); the
    tests prove both forms agree, so the kernel can be checked against either.

dtype rule (NumPy promotion, probed on the reference): with T1 = dtype(scan)
and T2 = promote(T1, dtype(phi)), `tmp_y`, `atan2` and `r'` are T1, while
`phi'`, x, y and every distance are T2.  The dataset path feeds float32 scans
with a float64 angle grid (src/utils/dataset_dr_spaam.py:477, utils.py:25-29);
the streaming script feeds float64 scans (depracted_scripts/infer_person_flow.py:54-56).
"""
import numpy as np


def votes_to_xy(scan_grid, phi_grid, pred_reg):
    """Canonical (dx, dy) votes -> global x, y.   utils.py:109-116, :47-48."""
    dx, dy = pred_reg[:, 0], pred_reg[:, 1]
    fwd = scan_grid + dy                       # :110
    bearing = np.arctan2(dx, fwd)              # :111  (dx first, by geometry)
    phi_v = bearing + phi_grid                 # :114
    r_v = fwd / np.cos(bearing)                # :115
    return r_v * np.cos(phi_v), r_v * np.sin(phi_v)   # :48


def descending_order(scores):
    """`argsort(...)[::-1]` of utils.py:544.

    NumPy's default argsort is unstable, so the reference's order on TIED
    scores is unspecified; parity is defined on tie-free scores only
    (SURVEY.md §7 hard part 6).  The product kernel breaks ties by descending
    point index, i.e. `argsort(kind="stable")[::-1]`; the oracle uses the same
    rule so tied inputs still compare deterministically.
    """
    return np.argsort(scores, kind="stable")[::-1]


def nms_predicted_center(scan_grid, phi_grid, pred_cls, pred_reg, min_dist=0.5):
    """Reference algorithm, reference cost.  Returns (det_xys, det_cls, instance_mask)."""
    assert pred_cls.shape[1] == 1                                    # :536
    xs, ys = votes_to_xy(scan_grid, phi_grid, pred_reg)              # :538-541
    order = descending_order(pred_cls[:, 0])                         # :544
    xs, ys = xs[order], ys[order]
    conf = pred_cls[order]

    n = len(scan_grid)
    ddx = xs[:, None] - xs[None, :]                                  # :550-552
    ddy = ys[:, None] - ys[None, :]
    dist = np.sqrt(np.square(ddx) + np.square(ddy))

    alive = np.ones(n, dtype=np.bool_)                               # :555-566
    instance_mask = np.zeros(n, dtype=np.int32)
    next_id = 1
    for a in range(n):
        if not alive[a]:
            continue
        near = dist[a] < min_dist
        alive[near] = False
        alive[a] = True
        instance_mask[order[near]] = next_id      # later centres overwrite earlier ones
        next_id += 1

    det_xys = np.stack((xs, ys), axis=1)[alive]                      # :568-569
    det_cls = conf[alive]
    return det_xys, det_cls, instance_mask


def nms_sweep_spec(scan_grid, phi_grid, pred_cls, pred_reg, min_dist=0.5):
    """Adjacency + single-sweep form (what the CUDA kernel does).

    Returns dict(order, keep_sorted, keep_idx, instance_mask, det_xys, det_cls,
    margin) where `keep_idx` are ORIGINAL point indices of the kept centres in
    descending-confidence order and `margin` is min |dist - min_dist| over all
    pairs (how close the input is to a threshold flip).
    """
    xs, ys = votes_to_xy(scan_grid, phi_grid, pred_reg)
    order = descending_order(pred_cls[:, 0])
    xs, ys = xs[order], ys[order]
    n = len(order)
    dist = np.sqrt(np.square(xs[:, None] - xs[None, :]) + np.square(ys[:, None] - ys[None, :]))
    adj = dist < min_dist
    suppressed = np.zeros(n, dtype=bool)
    keep = np.zeros(n, dtype=bool)
    for a in range(n):
        if suppressed[a]:
            continue
        keep[a] = True
        suppressed |= adj[a]
    ids = np.cumsum(keep)                       # 1-based id of each kept centre
    # id of the LAST kept centre adjacent to b  (reference: last writer wins)
    tagged = np.where(adj & keep[None, :], ids[None, :], 0)
    inst_sorted = tagged.max(axis=1).astype(np.int32)
    instance_mask = np.zeros(n, dtype=np.int32)
    instance_mask[order] = inst_sorted
    off = np.abs(dist - min_dist)
    return {
        "order": order.astype(np.int32),
        "keep_sorted": keep,
        "keep_idx": order[keep].astype(np.int32),
        "instance_mask": instance_mask,
        "det_xys": np.stack((xs, ys), axis=1)[keep],
        "det_cls": pred_cls[order][keep],
        "margin": float(off.min()) if n else float("inf"),
    }
