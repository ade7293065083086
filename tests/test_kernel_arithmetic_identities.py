"""Arithmetic identities the CUDA kernels rely on, restated in NumPy and checked on the CPU.

  * csrc/pof_cutout.cu::div_taps - the area-mode mean divides by s_area with Markstein's three operations
    (q0 = a * y, r = fma(-q0, s, a), q = fma(r, y, q0), y = RN(1/s)) instead of a float32 division;
  * csrc/pof_nms.cu::largest_square_below - `sqrt(dx^2 + dy^2) < min_dist` (utils.py:562) is decided on the sum of
    squares against the largest value whose ROUNDED square root is still below the threshold.
"""
import numpy as np
import pytest


@pytest.mark.parametrize("s", list(range(2, 33)))
def test_markstein_division_by_small_integers_is_correctly_rounded(s):
    rng = np.random.RandomState(s)
    a = np.concatenate([rng.uniform(0, 300, 200000), rng.randint(1, 2 ** 24, 100000) * 2.0 ** rng.randint(-12, 8, 100000)]).astype(np.float32)
    y = np.float32(1.0) / np.float32(s)
    q0 = (a * y).astype(np.float32)
    r = a.astype(np.float64) - q0.astype(np.float64) * s                 # exact, and representable in float32
    assert np.array_equal(r.astype(np.float32).astype(np.float64), r)
    q = (q0.astype(np.float64) + r * np.float64(y)).astype(np.float32)
    assert np.array_equal(q, (a / np.float32(s)).astype(np.float32))


def _largest_square_below(thr, dtype):
    thr = dtype(thr)
    if not thr > 0:
        return dtype(-1)
    s = dtype(thr * thr)
    for _ in range(64):
        if not np.sqrt(s) < thr:
            break
        s = np.nextafter(s, dtype(np.inf))
    for _ in range(128):
        if np.sqrt(s) < thr:
            break
        s = np.nextafter(s, dtype(-np.inf))
    return s


@pytest.mark.parametrize("dtype", [np.float32, np.float64])
@pytest.mark.parametrize("thr", [0.5, 0.3, 1.0, 100.0, 1e-3, 0.7071067811865476, 2.0 ** -60, 3.3333333])
def test_threshold_on_the_sum_of_squares_is_the_same_predicate(dtype, thr):
    s_max = _largest_square_below(thr, dtype)
    t = dtype(thr)
    assert np.sqrt(s_max) < t and not np.sqrt(np.nextafter(s_max, dtype(np.inf))) < t
    rng = np.random.RandomState(1)
    q = np.concatenate([rng.uniform(0, 4 * thr * thr, 200000),
                        float(s_max) * (1 + rng.uniform(-1e-6, 1e-6, 200000)),
                        [float(s_max), float(np.nextafter(s_max, dtype(np.inf))), float(np.nextafter(s_max, dtype(-np.inf))), 0.0]]).astype(dtype)
    assert np.array_equal(np.sqrt(q) < t, q <= s_max)


def test_non_positive_threshold_keeps_nothing():
    assert _largest_square_below(0.0, np.float32) == -1 and _largest_square_below(-1.0, np.float64) == -1
