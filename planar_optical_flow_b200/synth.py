"""Seeded synthetic laser scans and network outputs.

No dataset ships with the reference and there is no network, so every test and
benchmark runs on these generators.  Shapes follow the two conventions found in
the reference:
  * DROW-shaped : 450 points, 0.5 deg pitch, 224.5 deg fov, float64 angle grid
                  (`get_laser_phi`, /root/reference/src/utils/utils.py:25-29)
  * JRDB-shaped : 1091 points over [-pi, pi], float32 angle grid
                  (/root/reference/src/depracted/data_handle/jrdb_handle.py:119-121)
NumPy only; nothing here touches the GPU.
"""
import numpy as np

DROW_POINTS = 450
JRDB_POINTS = 1091
MAX_RANGE = 29.99


def drow_phi(num_pts=DROW_POINTS, angle_inc=np.radians(0.5)):
    fov = (num_pts - 1) * angle_inc
    return np.linspace(-fov * 0.5, fov * 0.5, num_pts)


def jrdb_phi(num_pts=JRDB_POINTS):
    return np.linspace(-np.pi, np.pi, num_pts, dtype=np.float32)


def phi_for(shape):
    if shape == "drow":
        return drow_phi()
    if shape == "jrdb":
        return jrdb_phi()
    raise ValueError("unknown scan shape %r" % (shape,))


def adversarial_scans(n_scans, n_pts, seed, lo=0.3, hi=25.0):
    """i.i.d. uniform ranges: maximises discontinuities and area-mode rows."""
    rs = np.random.RandomState(seed)
    return rs.uniform(lo, hi, size=(n_scans, n_pts)).astype(np.float32)


def structured_sequence(n_steps, n_pts, seed, phi=None):
    """A walk through a piecewise-linear room with moving leg-like arcs.

    Returns [n_steps, n_pts] float32.  Walls in [2, 15] m, up to 8 legs of
    radius 0.05-0.15 m at 0.5-10 m, 3 % max-range returns, N(0, 0.01) noise,
    clipped to [0.05, 29.99].
    """
    rs = np.random.RandomState(seed)
    if phi is None:
        phi = np.linspace(-np.pi, np.pi, n_pts)
    phi = np.asarray(phi, dtype=np.float64)
    n_knots = int(rs.randint(6, 14))
    knots = np.sort(rs.uniform(phi[0], phi[-1], n_knots))
    knot_r = rs.uniform(2.0, 15.0, n_knots)
    walls = np.interp(phi, knots, knot_r)
    n_legs = int(rs.randint(0, 9))
    leg_r = rs.uniform(0.5, 10.0, n_legs)
    leg_phi = rs.uniform(phi[0], phi[-1], n_legs)
    leg_rad = rs.uniform(0.05, 0.15, n_legs)
    leg_vel = rs.normal(0.0, 0.01, size=(n_legs, 2))
    out = np.empty((n_steps, n_pts), dtype=np.float32)
    shift = 0.0
    for t in range(n_steps):
        shift += rs.uniform(-0.05, 0.05)
        r = walls + shift
        for k in range(n_legs):
            lr = leg_r[k] + leg_vel[k, 0] * t
            lp = leg_phi[k] + leg_vel[k, 1] * t
            half = np.arctan2(leg_rad[k], max(lr, 0.2))
            hit = np.abs(phi - lp) < half
            chord = lr - leg_rad[k] * np.cos((phi - lp) / max(half, 1e-6) * (np.pi / 2))
            r = np.where(hit & (chord < r), chord, r)
        r = r + rs.normal(0.0, 0.01, n_pts)
        far = rs.rand(n_pts) < 0.03
        r = np.where(far, 29.96, r)
        out[t] = np.clip(r, 0.05, MAX_RANGE).astype(np.float32)
    return out


def edge_scans(n_pts, seed):
    """Degenerate ranges: <= 1e-2 (clamped half-angle), exactly padding_val, tiny steps."""
    rs = np.random.RandomState(seed)
    base = rs.uniform(0.3, 25.0, size=(4, n_pts)).astype(np.float32)
    base[0, :: 7] = 0.0
    base[0, 3:: 11] = 0.01
    base[1, :] = MAX_RANGE
    base[2, : n_pts // 2] = 0.005
    base[3, :: 2] = 0.02
    return base


def distinct_scores(n, seed):
    """Sigmoid-shaped confidences in (0,1) with pairwise-distinct float32 values."""
    rs = np.random.RandomState(seed)
    s = (1.0 / (1.0 + np.exp(-rs.normal(0.0, 2.0, 4 * n)))).astype(np.float32)
    s = np.unique(s)
    rs.shuffle(s)
    assert len(s) >= n
    return s[:n].reshape(n, 1).copy()


def clustered_votes(scan, phi, seed, n_people=12, spread=0.08):
    """Offsets (dx, dy) in each point's canonical frame that vote for a few centres.

    Roughly half of the points vote for one of `n_people` centres (with
    `spread` metres of noise), the rest vote for themselves plus noise, which
    yields both large groups and many singletons after NMS.
    """
    rs = np.random.RandomState(seed)
    n = len(scan)
    phi64 = np.asarray(phi, dtype=np.float64)
    px, py = scan * np.cos(phi64), scan * np.sin(phi64)
    centres = rs.randint(0, n, n_people)
    cx = px[centres] + rs.normal(0, 0.3, n_people)
    cy = py[centres] + rs.normal(0, 0.3, n_people)
    who = rs.randint(0, n_people, n)
    voter = rs.rand(n) < 0.5
    tx = np.where(voter, cx[who], px) + rs.normal(0, spread, n)
    ty = np.where(voter, cy[who], py) + rs.normal(0, spread, n)
    # global (tx, ty) -> canonical offsets about each point (inverse of utils.py:109-116)
    tr, tphi = np.hypot(tx, ty), np.arctan2(ty, tx)
    dx = np.sin(tphi - phi64) * tr
    dy = np.cos(tphi - phi64) * tr - scan
    return np.stack((dx, dy), axis=1).astype(np.float32)


def feature_like(shape, seed, negative_slope=0.1):
    """Post-LeakyReLU-like activations: N(0,1) with the negative side scaled."""
    rs = np.random.RandomState(seed)
    z = rs.standard_normal(shape).astype(np.float32)
    return np.where(z > 0, z, negative_slope * z).astype(np.float32)
