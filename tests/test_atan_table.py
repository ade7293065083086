"""The scan cutout kernel evaluates its arctangent from a table of degree-7 Taylor cells compiled into
csrc/pof_cutout.cu (kAtanTab).  This CPU test re-reads the table from the source and checks that the polynomial a
lane evaluates rounds to the same float32 as float32(atan(float64(x))) - the arithmetic the EXACT kernels use - so an
edit of the table cannot go unnoticed."""
import math
import os
import re

import numpy as np

SRC = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "planar_optical_flow_b200", "csrc", "pof_cutout.cu")


def _table():
    text = open(SRC).read()
    body = text[text.index("kAtanTab[kAtanDeg][kAtanCells] = {"):]
    body = body[body.index("{") + 1:body.index("};")]
    rows = re.findall(r"\{([^{}]*)\}", body)
    tab = np.array([[float(v) for v in r.split(",") if v.strip()] for r in rows])
    assert tab.shape == (8, 33)
    return tab


def _atan_unit(x, tab):
    j = np.rint(x.astype(np.float32) * np.float32(32.0)).astype(np.int64)
    h = x - j * 0.03125
    r = tab[7][j]
    for k in range(6, -1, -1):
        r = r * h + tab[k][j]
    return r


def test_table_cells_are_the_taylor_coefficients_of_atan():
    tab = _table()
    for j in range(33):
        t = math.atan(j / 32.0)
        assert abs(tab[0][j] - t) <= 1e-16 + 1e-15 * t
        for k in range(1, 8):
            want = math.cos(t) ** k * math.sin(k * (t + math.pi / 2)) / k
            assert abs(tab[k][j] - want) <= 1e-15, (k, j)


def test_polynomial_rounds_like_the_float64_arctangent():
    tab = _table()
    rng = np.random.RandomState(0)
    x = np.concatenate([rng.uniform(0, 1, 400000), np.linspace(0, 1, 4097), 10.0 ** rng.uniform(-7, 0, 100000)]).astype(np.float32)
    x = np.clip(x, 0, 1).astype(np.float64)
    got = _atan_unit(x, tab)
    want = np.arctan(x)
    assert np.abs(got - want).max() <= 2e-15
    assert np.array_equal(got.astype(np.float32), want.astype(np.float32))
    # ratios above one go through the reciprocal: pi/2 - atan(1/x)
    big = (1.0 / np.clip(rng.uniform(0.01, 1, 200000), 1e-3, 1)).astype(np.float32).astype(np.float64)
    inv = 1.0 / big
    j = np.rint(inv.astype(np.float32) * np.float32(32.0)).astype(np.int64)
    h = inv - j * 0.03125
    r = tab[7][j]
    for k in range(6, -1, -1):
        r = r * h + tab[k][j]
    got_big = (1.5707963267948966 - r).astype(np.float32)
    ulps = np.abs(got_big.view(np.int32).astype(np.int64) - np.arctan(big).astype(np.float32).view(np.int32).astype(np.int64))
    assert ulps.max() <= 1 and (ulps > 0).mean() < 1e-3
