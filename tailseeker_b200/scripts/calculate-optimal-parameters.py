#!/usr/bin/env python3
"""Drop-in for the reference's scripts/calculate-optimal-parameters.py (rule
calculate_optimal_parameters, tailseeker/main.py:212-222): same SNAKEMAKE_PARAMS in, same
signal-cutoffs.txt / cutoff-bases text out.  The weighted-KDE + bisection of
find_all_eqodds_point() (:57-78) runs on the GPU (tsk_fit_cutoffs): the run-wide histograms of
every source and all their tiles go down in ONE call, one CTA per (histogram, cycle).  The
bookkeeping around it (which cycles are tried, interpolation of the missing ones, the d/i/x
report) is tailseeker_b200/estimator.py."""
import gzip
import os
import struct
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))

from tailseeker_b200 import estimator                       # noqa: E402
from tailseeker_b200.scripts import _params                 # noqa: E402


def load_sig_dists(filename):
    with gzip.open(filename, "rb") as f:
        elemsize, cycles, bins = struct.unpack("<III", f.read(12))
        if elemsize != 4:
            raise SystemExit("%s: element size %d, expected 4" % (filename, elemsize))
        return np.frombuffer(f.read(), np.uint32).reshape((cycles, bins)).astype(np.uint64)


def main():
    v = _params.load()
    inp, out, params = v["input"], v["output"], v["params"]
    conf = params["conf"]["polyA_seeder"]
    total_cycles = params.total_cycles
    sigdists = {}
    for f in inp:                      # .../{pos,neg}_{tile}.sigdists (the reference's own parsing, :48-53)
        tilename = f.rsplit("_", 1)[1].split(".", 1)[0]
        signaltype = f.rsplit("/", 1)[1].split("_", 1)[0]
        sigdists[tilename, signaltype] = load_sig_dists(f)
    sources = {}
    for tileinfo in params.tileinfo.values():
        sources.setdefault(tileinfo["source"], []).append(tileinfo["id"])

    from tailseeker_b200.importer import TileProcessor
    from tailseeker_b200.config import stock_config
    cfg = stock_config("miseq").replace(kde_bandwidth_for_dist=float(conf["kde_bandwidth_for_dist"]),
                                        cutoff_score_search_low=float(conf["cutoff_score_search_low"]),
                                        cutoff_score_search_high=tuple(float(x) for x in conf["cutoff_score_search_high"]),
                                        minimum_spots_for_dist_sampling=int(conf["minimum_spots_for_dist_sampling"]))
    proc = TileProcessor(cfg, device=int(os.environ.get("TAILSEEKER_DEVICE", "0")))
    try:
        cutoffs, bases = {}, {}
        for source, tiles in sources.items():
            c, b = estimator.calculate_optimal_cutoffs_batched(
                proc.fit_cutoffs, sigdists, tiles, source, total_cycles,
                int(conf["minimum_spots_for_dist_sampling"]))
            cutoffs.update(c)
            bases.update(b)
    finally:
        proc.close()
    with open(out.cutoff_values, "wt") as f:
        f.write(estimator.format_cutoffs(cutoffs, total_cycles))
    with open(out.cutoff_bases, "wt") as f:
        f.write(estimator.format_bases(bases))


if __name__ == "__main__":
    main()
