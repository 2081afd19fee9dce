#!/usr/bin/env python3
"""Drop-in for the reference's scripts/measure-polya-lengths.py (rule
measure_polya_lengths_from_fluorescence, tailseeker/main.py:228-240): same SNAKEMAKE_PARAMS in, same
per-sample corrected taginfo (bgzip-compressed) and summed positive .sigdists out.  The reference
starts one tailseq-polya-ruler process per sample (:53-62); here every sample of the tile goes to
ONE `tailseq-polya-ruler --batch` process, i.e. one CUDA context, and the per-sample histograms
are summed like :64-70."""
import gzip
import os
import shutil
import struct
import subprocess
import sys
import tempfile

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))

from tailseeker_b200.scripts import _params                 # noqa: E402


def load_sig_dists(filename):
    with gzip.open(filename, "rb") as f:
        elemsize, cycles, bins = struct.unpack("<III", f.read(12))
        assert elemsize == 4
        return np.frombuffer(f.read(), np.uint32).reshape((cycles, bins)).astype(np.uint64)


def write_sig_dists(counts, filename):
    with gzip.open(filename, "wb") as f:
        f.write(struct.pack("<III", 4, counts.shape[0], counts.shape[1]))
        f.write(counts.astype(np.uint32).tobytes())


def main():
    v = _params.load()
    inp, out, params, wildcards = v["input"], v["output"], v["params"], v["wildcards"]
    conf = params.CONF
    bindir = params.BINDIR
    bgzip = params.BGZIP_CMD
    signals = inp.signals if isinstance(inp.signals, list) else [inp.signals]
    taginfo = inp.taginfo if isinstance(inp.taginfo, list) else [inp.taginfo]
    outs = out.taginfo if isinstance(out.taginfo, list) else [out.taginfo]
    work = tempfile.mkdtemp(prefix="tskruler_")
    try:
        jobs, parts = [], []
        for i, (sig, tag, dst) in enumerate(zip(signals, taginfo, outs)):
            hist = os.path.join(work, "pos_%d.sigdists" % i)
            text = os.path.join(work, "taginfo_%d.txt" % i)
            parts.append((hist, text, dst))
            jobs.append("\t".join(str(x) for x in (
                wildcards.tile, sig, inp.score_cutoffs, conf["polyA_finder"]["signal_analysis_trigger"],
                conf["polyA_ruler"]["downhill_extension_weight"], tag, conf["polyA_seeder"]["dist_sampling_bins"],
                conf["polyA_ruler"]["signal_resampling_gap"], hist, text)))
        joblist = os.path.join(work, "jobs.tsv")
        with open(joblist, "w") as f:
            f.write("\n".join(jobs) + "\n")
        subprocess.check_call([os.path.join(bindir, "tailseq-polya-ruler"), "--batch", joblist])
        counts = None
        for hist, text, dst in parts:
            with open(dst, "wb") as g:                       # `... | bgzip -c > out` of the reference
                subprocess.check_call("%s -c < %s" % (bgzip, text), shell=True, stdout=g)
            c = load_sig_dists(hist)
            counts = c if counts is None else counts + c
        write_sig_dists(counts, out.sigdists)
    finally:
        shutil.rmtree(work, ignore_errors=True)


if __name__ == "__main__":
    main()
