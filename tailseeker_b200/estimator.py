"""Signal-cutoff estimation: host-side mirror of scripts/calculate-optimal-parameters.py.

The heavy part -- per cycle, two weighted Gaussian KDEs over 1000 score bins and a bisection on
their log-likelihood ratio (find_eqodds_point, :57-71) -- runs on the GPU through
tsk_fit_cutoffs().  What stays on the host is bookkeeping on <= a few hundred numbers per tile:
which cycles are tried (:73-78), interpolation / extrapolation of the cycles without enough
spots (infer_missing_cutoffs, :80-124), the per-tile fall-back to the run-wide cutoffs
(:157-169) and the text outputs (:173-186).
"""
from __future__ import annotations

from typing import Callable, Dict, Iterable, List, Optional, Tuple

import numpy as np

NUM_CLOSEST_CYCLES_INTERPOLATION = 3     # calculate-optimal-parameters.py:35
NUM_CLOSEST_CYCLES_EXTRAPOLATION = 5     # :36

Cutoffs = Dict[int, Optional[float]]


def tried_cycles(pos: np.ndarray, neg: np.ndarray, min_spots: int) -> List[int]:
    """find_all_eqodds_point(), :73-75: cycles with enough positive AND negative samples."""
    ok = (pos.sum(axis=1) >= min_spots) & (neg.sum(axis=1) >= min_spots)
    return np.where(ok)[0].tolist()


def eqodds_points(fit: Callable[[np.ndarray, np.ndarray], np.ndarray], pos: np.ndarray,
                  neg: np.ndarray, min_spots: int) -> Cutoffs:
    """{cycle: root or None} for the tried cycles (:73-78); `fit` is TileProcessor.fit_cutoffs."""
    roots = fit(pos, neg)
    return {c: (None if np.isnan(roots[c]) else float(roots[c]))
            for c in tried_cycles(pos, neg, min_spots)}


def infer_missing_cutoffs(cutoffs: Cutoffs, allcycles: Iterable[int]) -> Cutoffs:
    """:80-124.  Cycles absent from `cutoffs` are filled per run of consecutive missing cycles:
    linear interpolation between the 3 nearest determined cycles on each side, or the median of
    the 5 nearest on the only side that has any.  Cycles that were tried without a root stay None."""
    have = np.array(sorted(c for c, v in cutoffs.items() if v is not None), dtype=np.int64)
    missing = sorted(set(allcycles) - set(cutoffs))
    filled: Cutoffs = {}
    runs: List[Tuple[int, int]] = []
    for c in missing:
        if runs and c == runs[-1][1] + 1:
            runs[-1] = (runs[-1][0], c)
        else:
            runs.append((c, c))
    for first, last in runs:
        span = list(range(first, last + 1))
        left = have[have < first].tolist()
        right = have[have > last].tolist()
        if left and right:
            lc = left[-NUM_CLOSEST_CYCLES_INTERPOLATION:]
            rc = right[:NUM_CLOSEST_CYCLES_INTERPOLATION]
            xs = lc + rc
            ys = [cutoffs[c] for c in xs]
            for c, v in zip(span, np.interp(span, xs, ys)):
                filled[c] = float(v)
        else:
            near = left[-NUM_CLOSEST_CYCLES_EXTRAPOLATION:] if left else \
                right[:NUM_CLOSEST_CYCLES_EXTRAPOLATION]
            v = float(np.median([cutoffs[c] for c in near])) if near else float("nan")
            for c in span:
                filled[c] = v
    out = dict(cutoffs)
    out.update(filled)
    return out


def report_marks(direct: Cutoffs, inferred: Cutoffs, total_cycles: int) -> str:
    """get_report_marks(), :126-135: d = determined, i = inferred, x = none."""
    return "".join("d" if direct.get(i) is not None else ("i" if inferred.get(i) is not None else "x")
                   for i in range(total_cycles))


def calculate_optimal_cutoffs(fit, sigdists: Dict[Tuple[str, str], np.ndarray], tile_ids: List[str],
                              run_id: str, total_cycles: int, min_spots: int,
                              postotal: Optional[np.ndarray] = None,
                              negtotal: Optional[np.ndarray] = None):
    """calculate_optimal_cutoffs(), :137-171.  `postotal` / `negtotal` may be supplied when the
    cross-tile sums were already formed (e.g. all-reduced across GPUs); otherwise they are
    summed here from `sigdists` like :139-140."""
    if postotal is None:
        postotal = np.sum([sigdists[t, "pos"].astype(np.uint64) for t in tile_ids], axis=0)
    if negtotal is None:
        negtotal = np.sum([sigdists[t, "neg"].astype(np.uint64) for t in tile_ids], axis=0)
    all_cycles = np.where(negtotal.sum(axis=1) > 0)[0].tolist()
    runwide = eqodds_points(fit, postotal, negtotal, min_spots)
    runwide_inferred = infer_missing_cutoffs(runwide, all_cycles)
    reports = {run_id: report_marks(runwide, runwide_inferred, total_cycles)}
    result = {}
    for t in sorted(tile_ids):
        own = eqodds_points(fit, sigdists[t, "pos"].astype(np.uint64),
                            sigdists[t, "neg"].astype(np.uint64), min_spots)
        merged = dict(own)
        determined = {c for c, v in own.items() if v is not None}
        for c in set(all_cycles) - determined:
            merged[c] = runwide_inferred[c]
        result[t] = merged
        reports[t] = report_marks(own, merged, total_cycles)
    return result, reports


def calculate_optimal_cutoffs_batched(fit, sigdists: Dict[Tuple[str, str], np.ndarray], tile_ids: List[str],
                                      run_id: str, total_cycles: int, min_spots: int):
    """calculate_optimal_cutoffs() with ONE call of `fit`: the run-wide histograms and every tile's
    are stacked row-wise ([(1 + tiles) * cycles][bins]; the fit treats each row on its own), so a run
    of any size costs one launch and two copies instead of one launch and two copies per tile
    (the reference spreads the tiles over a process pool, :149-155)."""
    order = sorted(tile_ids)
    postotal = np.sum([sigdists[t, "pos"].astype(np.uint64) for t in tile_ids], axis=0)
    negtotal = np.sum([sigdists[t, "neg"].astype(np.uint64) for t in tile_ids], axis=0)
    cycles = postotal.shape[0]
    pos = np.concatenate([postotal] + [sigdists[t, "pos"].astype(np.uint64) for t in order])
    neg = np.concatenate([negtotal] + [sigdists[t, "neg"].astype(np.uint64) for t in order])
    roots = np.asarray(fit(pos, neg)).reshape(1 + len(order), cycles)

    def points(k, p, n):
        return {c: (None if np.isnan(roots[k][c]) else float(roots[k][c])) for c in tried_cycles(p, n, min_spots)}

    all_cycles = np.where(negtotal.sum(axis=1) > 0)[0].tolist()
    runwide = points(0, postotal, negtotal)
    runwide_inferred = infer_missing_cutoffs(runwide, all_cycles)
    reports = {run_id: report_marks(runwide, runwide_inferred, total_cycles)}
    result = {}
    for k, t in enumerate(order):
        own = points(1 + k, sigdists[t, "pos"], sigdists[t, "neg"])
        merged = dict(own)
        determined = {c for c, v in own.items() if v is not None}
        for c in set(all_cycles) - determined:
            merged[c] = runwide_inferred[c]
        result[t] = merged
        reports[t] = report_marks(own, merged, total_cycles)
    return result, reports


def format_cutoffs(cutoffs: Dict[str, Cutoffs], total_cycles: int) -> str:
    """write_outputs(), :173-182: signal-cutoffs.txt."""
    lines = []
    for tile, cut in sorted(cutoffs.items()):
        cells = [(format(cut[i], ".6f") if cut.get(i) is not None else "nan") + "\t"
                 for i in range(total_cycles)]
        lines.append(tile + "\t" + "".join(cells) + "\n")
    return "".join(lines)


def format_bases(reports: Dict[str, str]) -> str:
    """write_outputs(), :184-186."""
    return "".join("%s\t%s\n" % (k, v) for k, v in sorted(reports.items()))
