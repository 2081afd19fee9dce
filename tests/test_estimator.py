"""Cutoff estimation: host bookkeeping (tailseeker_b200/estimator.py) and the per-cycle fit against
golden output of the reference's scripts/calculate-optimal-parameters.py."""
import numpy as np
import pytest

import golden_util as gu


def _sigdists(g):
    shape = tuple(int(v) for v in g["shape"])
    out = {}
    for tid in ("a1101", "a1102"):
        for kind in ("pos", "neg"):
            arr = np.zeros(shape, dtype=np.uint32)
            idx = g["%s_%s_idx" % (kind, tid)]
            arr[idx[0], idx[1]] = g["%s_%s_val" % (kind, tid)]
            out[tid, kind] = arr
    return out


def _run(fit):
    from tailseeker_b200 import estimator
    g = gu.load("estimator_two_tiles")
    sig = _sigdists(g)
    T = int(g["shape"][0])
    cut, reports = estimator.calculate_optimal_cutoffs(fit, sig, ["a1101", "a1102"], "run", T, 300)
    assert estimator.format_cutoffs(cut, T) == bytes(g["text"]).decode()
    assert estimator.format_bases(reports) == bytes(g["bases"]).decode()


def test_estimator_with_oracle_fit():
    from oracle import estimator_oracle
    _run(lambda p, n: estimator_oracle.fit_cutoffs(p, n, 300, 0.1, 0.0, (0.4, 0.6)))


@pytest.mark.gpu
def test_estimator_with_gpu_fit(stock_cfg):
    from tailseeker_b200.importer import TileProcessor
    from oracle import estimator_oracle
    proc = TileProcessor(stock_cfg)
    try:
        _run(proc.fit_cutoffs)
        # and value-level agreement with the NumPy restatement (tolerance: 1e-9 absolute; the
        # printed precision is 1e-6)
        g = gu.load("estimator_two_tiles")
        sig = _sigdists(g)
        a = proc.fit_cutoffs(sig["a1101", "pos"], sig["a1101", "neg"])
        b = estimator_oracle.fit_cutoffs(sig["a1101", "pos"], sig["a1101", "neg"])
        assert (np.isnan(a) == np.isnan(b)).all()
        assert np.nanmax(np.abs(a - b)) < 1e-9
    finally:
        proc.close()


def test_infer_missing_cutoffs_shapes():
    from tailseeker_b200.estimator import infer_missing_cutoffs
    cut = {10: 0.2, 11: 0.25, 12: 0.3, 20: 0.5, 21: None}
    out = infer_missing_cutoffs(cut, range(5, 25))
    assert out[21] is None                              # tried without a root: stays None
    assert out[5] == out[9] == 0.275                    # left edge: median of the (up to 5) nearest on the right
    assert abs(out[16] - np.interp(16, [10, 11, 12, 20], [0.2, 0.25, 0.3, 0.5])) < 1e-15
    assert out[24] == np.median([0.2, 0.25, 0.3, 0.5])  # right edge: median of the nearest on the left
