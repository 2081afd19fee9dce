"""TEST INFRASTRUCTURE ONLY -- NumPy restatement of the reference's per-cycle cutoff fit:
find_eqodds_point() (scripts/calculate-optimal-parameters.py:57-71) on top of the weighted
gaussian_kde of tailseeker/stats.py:411-466,540-558 and scipy.optimize.bisect's algorithm
(scipy/optimize/Zeros/bisect.c; xtol=2e-12, rtol=4*eps, maxiter=100).

Pinned by tests/test_estimator.py against golden outputs of the reference script itself
(tests/golden/make_golden.py ran it in the build container)."""
from __future__ import annotations

import math

import numpy as np

XTOL = 2e-12
RTOL = 8.881784197001252e-16
VERY_SMALL_PROBABILITY = 0.001


class WeightedKde:
    def __init__(self, xs: np.ndarray, counts: np.ndarray, bw: float):
        self.xs = xs
        self.w = counts / np.sum(counts)
        mean = np.sum(self.w * xs)
        resid = xs - mean
        cov = np.dot(resid * self.w, resid) / (1 - np.sum(self.w ** 2))
        self.inv_cov = (1.0 / cov) / bw ** 2
        self.norm = math.sqrt(2 * math.pi * (cov * bw ** 2))

    def __call__(self, x: float) -> float:
        d = x - self.xs
        chi2 = np.sqrt(d * self.inv_cov * d) ** 2
        return float(np.sum(np.exp(-.5 * chi2) * self.w) / self.norm)


def _bisect(f, a, b):
    fa, fb = f(a), f(b)
    if fa * fb > 0:
        raise ValueError("f(a) and f(b) must have different signs")
    if fa == 0:
        return a
    if fb == 0:
        return b
    dm = b - a
    xm = a
    for _ in range(100):
        dm *= .5
        xm = a + dm
        fm = f(xm)
        if fm * fa >= 0:
            a = xm
        if fm == 0 or abs(dm) < XTOL + RTOL * abs(xm):
            return xm
    return xm


def fit_cutoffs(pos: np.ndarray, neg: np.ndarray, min_spots: int = 300, bw: float = 0.1,
                low: float = 0.0, highs=(0.4, 0.6)) -> np.ndarray:
    """Per-cycle equal-odds point, NaN where the cycle has too few spots or no root."""
    pos = np.asarray(pos, dtype=np.uint64)
    neg = np.asarray(neg, dtype=np.uint64)
    cycles, bins = pos.shape
    xs = np.arange(0, 1, 1 / bins)
    out = np.full(cycles, np.nan)
    for c in range(cycles):
        if pos[c].sum() < min_spots or neg[c].sum() < min_spots:
            continue
        kp = WeightedKde(xs, pos[c], bw)
        kn = WeightedKde(xs, neg[c], bw)
        f = lambda x: math.log(max(VERY_SMALL_PROBABILITY, kp(x)) / max(VERY_SMALL_PROBABILITY, kn(x)))
        for hi in highs:
            try:
                r = _bisect(f, low, hi)
            except ValueError:
                continue
            if low < r < hi:
                out[c] = r
                break
    return out
