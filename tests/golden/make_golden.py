#!/usr/bin/env python
"""Generates the golden fixtures of tests/golden/ by running the UNMODIFIED reference in the build
container (oracle/_ref binaries built from /root/reference/src by oracle/Makefile, and the
reference's scripts/calculate-optimal-parameters.py imported from /root/reference).  The
reference ships no test vectors of its own, so these are the pins.

    python tests/golden/make_golden.py        # rewrites tests/golden/*.npz

Fixtures hold the (small) synthetic inputs as well as the reference outputs, so the tests do not
depend on the random stream of the generator.
"""
import os
import sys
import tempfile

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)

from oracle import refrun                                   # noqa: E402
from tailseeker_b200.config import stock_config             # noqa: E402
from tailseeker_b200.synth import make_tile                 # noqa: E402

SEEDER = dict(minimum_spots_for_dist_sampling=300, favored_spots_for_dist_sampling=500,
              kde_bandwidth_for_dist=0.1, cutoff_score_search_low=0.0,
              cutoff_score_search_high=[0.4, 0.6])


def _bytes(b):
    return np.frombuffer(b, dtype=np.uint8)


def importer_fixture(name, cfg, n, seed, **tilekw):
    tile = make_tile(cfg, n, seed=seed, **tilekw)
    work = tempfile.mkdtemp(prefix="tskgold_")
    r = refrun.run_import(cfg, tile, workdir=work)
    tid = cfg.tile_id
    rng = np.random.default_rng(seed)
    cut = np.clip(0.25 + 0.1 * rng.standard_normal(cfg.total_cycles), 0.01, 0.9)
    cut[rng.random(cfg.total_cycles) < 0.05] = np.nan
    cpath = os.path.join(work, "cutoffs.txt")
    refrun.write_cutoffs_file(cpath, tid, [float(c) for c in cut])
    out = dict(bcl=tile.bcl, cif=tile.cif, matrix_text=_bytes(tile.matrix_text.encode()),
               pos=r["pos"], neg=r["neg"], stats_csv=_bytes(r["stats_csv"].encode()),
               cutoffs=cut, cutoffs_text=_bytes(open(cpath, "rb").read()))
    for s in cfg.samples:
        out["taginfo_" + s.name] = _bytes(r["taginfo"][s.name])
        out["seqqual_" + s.name] = _bytes(r["seqqual"][s.name])
        total, dump, recs = r["sigpack"][s.name]
        out["sigpack_" + s.name] = recs
        out["sigpack_meta_" + s.name] = np.array([total, dump], dtype=np.int64)
        rr = refrun.run_ruler(tid, os.path.join(work, "signals", "%s_%s.sigpack" % (s.name, tid)), cpath,
                              os.path.join(work, "taginfo", "%s_%s.txt.gz" % (s.name, tid)),
                              os.path.join(work, "rulerpos_%s.gz" % s.name))
        out["ruler_taginfo_" + s.name] = _bytes(rr["taginfo"])
        out["ruler_pos_" + s.name] = rr["pos"]
    np.savez_compressed(os.path.join(HERE, name + ".npz"), **out)
    print(name, "clusters", n, "records", sum(out["sigpack_" + s.name].shape[0] for s in cfg.samples))


def estimator_fixture(name):
    """Histograms of two larger oracle-independent reference runs -> reference estimator output."""
    cfg = stock_config("miseq")
    sig = {}
    for k, (tid, seed) in enumerate((("a1101", 71), ("a1102", 72))):
        tile = make_tile(cfg.replace(tile=1101 + k), 40000 if k == 0 else 12000, seed=seed)
        r = refrun.run_import(cfg.replace(tile=1101 + k), tile)
        sig[tid, "pos"], sig[tid, "neg"] = r["pos"], r["neg"]
    est = refrun.run_reference_estimator(sig, cfg.total_cycles, SEEDER)
    out = dict(text=_bytes(est["text"].encode()), bases=_bytes(est["bases"].encode()))
    for (tid, kind), arr in sig.items():
        nz = np.nonzero(arr)
        out["%s_%s_idx" % (kind, tid)] = np.stack(nz).astype(np.int32)
        out["%s_%s_val" % (kind, tid)] = arr[nz]
    out["shape"] = np.array(sig["a1101", "pos"].shape)
    np.savez_compressed(os.path.join(HERE, name + ".npz"), **out)
    print(name, est["bases"])


def control_fixture(name, first=5, length=40, tile_n=2000, tile_seed=303):
    """Per-read scores / decisions of the reference control aligner (controlaligner.c + ssw.c via
    oracle/_ref/libtskref_control.so) and the demultiplexing statistics of the reference importer
    run with a [control] section on a tile regenerated from (tile_n, tile_seed)."""
    import ctypes
    sys.path.insert(0, os.path.dirname(HERE))
    from test_control import control_reads, REFLIB
    ref = ctypes.CDLL(REFLIB)
    assert ref.tskref_control_setup(first, length) == int(length * 0.65)
    reads = control_reads(4242, 3000)
    score = np.array([ref.tskref_control_score(bytes(r)) for r in reads], dtype=np.int32)
    decision = np.array([ref.tskref_control_try(bytes(r) + b"\0") for r in reads], dtype=np.int8)
    cfg = stock_config("miseq", phix_match_name="PhiX")
    tile = make_tile(cfg, tile_n, seed=tile_seed, unknown_frac=0.2, phix_frac=0.6)
    r = refrun.run_import(cfg, tile)
    np.savez_compressed(os.path.join(HERE, name + ".npz"), reads=reads, score=score, decision=decision,
                        first=first, length=length, tile_n=tile_n, tile_seed=tile_seed,
                        stats_csv=r["stats_csv"])
    print(name, "reads", len(reads), "aligned", int(decision.sum()), r["stats_csv"].splitlines()[:3])


def dedup_fixture(name):
    """Inputs for tailseq-dedup-perfect and the reference binary's stdout / duplicate trace."""
    import tempfile
    from oracle import dedup_oracle
    assert dedup_oracle.ref_available(), "build oracle/_ref first (make -C oracle ref)"
    cases = {"mixed": dict(n=2500, seed=11, n_umis=700, big_groups=(300, 70, 40)),
             "sparse": dict(n=1500, seed=12),
             "one_group": dict(n=600, seed=13, n_umis=1)}
    out = {}
    for key, kw in cases.items():
        data = dedup_oracle.synth_taginfo(**kw)
        with tempfile.TemporaryDirectory() as d:
            so, tr = dedup_oracle.run_reference(data, d)
        out[key + "_in"], out[key + "_out"], out[key + "_trace"] = _bytes(data), _bytes(so), _bytes(tr)
        print(name, key, len(data), len(so), len(tr))
    np.savez_compressed(os.path.join(HERE, name + ".npz"), **out)


if __name__ == "__main__":
    assert refrun.available(), "build oracle/_ref first (make -C oracle ref)"
    if sys.argv[1:] == ["dedup"]:
        dedup_fixture("dedup_perfect")
        sys.exit(0)
    if sys.argv[1:] == ["control"]:
        control_fixture("control_reads")
        sys.exit(0)
    importer_fixture("importer_miseq", stock_config("miseq"), 640, 101, dark_cluster_frac=0.02, lowdiv_frac=0.08)
    importer_fixture("importer_hiseq_keep", stock_config("hiseq", keep_no_delimiter=1, keep_low_quality_balancer=1),
                     384, 202, spikein_weight=1.0)
    estimator_fixture("estimator_two_tiles")
    control_fixture("control_reads")
    dedup_fixture("dedup_perfect")
