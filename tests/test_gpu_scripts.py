"""Script surfaces of the path (GPU part): the drop-in scripts run the way the pipeline runs them
(SNAKEMAKE_PARAMS), and the importer takes the INI rendered by the reference's own generator."""
import gzip
import json
import os
import subprocess
import sys

import numpy as np
import pytest

import golden_util as gu
from test_estimator import _sigdists

pytestmark = pytest.mark.gpu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SCRIPTS = os.path.join(ROOT, "tailseeker_b200", "scripts")
BIN = os.path.join(ROOT, "tailseeker_b200", "bin")


def _run_script(name, params, cwd):
    pj = os.path.join(cwd, "params_%s.json" % name.replace(".py", ""))
    with open(pj, "w") as f:
        json.dump(params, f)
    r = subprocess.run([sys.executable, os.path.join(SCRIPTS, name)], cwd=cwd, capture_output=True,
                       env=dict(os.environ, SNAKEMAKE_PARAMS=pj))
    assert r.returncode == 0, r.stderr.decode()[-3000:]


def test_calculate_optimal_parameters_script(tmp_path):
    from oracle import refrun
    g = gu.load("estimator_two_tiles")
    sig = _sigdists(g)
    T = int(g["shape"][0])
    inputs = []
    os.makedirs(str(tmp_path / "sigdists"))
    for (tile, kind), arr in sig.items():
        path = str(tmp_path / "sigdists" / ("%s_%s.sigdists" % (kind, tile)))
        refrun.write_sigdists(path, arr)
        inputs.append([None, path])
    params = {"input": inputs, "output": [["cutoff_values", str(tmp_path / "cut.txt")], ["cutoff_bases", str(tmp_path / "bases.txt")]],
              "threads": 2, "wildcards": [],
              "params": [["conf", {"polyA_seeder": dict(minimum_spots_for_dist_sampling=300, favored_spots_for_dist_sampling=1000,
                                                        kde_bandwidth_for_dist=0.1, cutoff_score_search_low=0.0,
                                                        cutoff_score_search_high=[0.4, 0.6])}],
                         ["tileinfo", {t: {"id": t, "source": "run"} for t in ("a1101", "a1102")}], ["total_cycles", T]]}
    _run_script("calculate-optimal-parameters.py", params, str(tmp_path))
    assert open(str(tmp_path / "cut.txt")).read() == bytes(g["text"]).decode()          # the reference script's own output
    assert open(str(tmp_path / "bases.txt")).read() == bytes(g["bases"]).decode()


@pytest.mark.skipif("not __import__('oracle.refrun').refrun.available()")
def test_importer_takes_the_generated_ini_and_ruler_script(tmp_path):
    from oracle import refrun
    from tailseeker_b200.config import stock_config
    from tailseeker_b200.synth import make_tile, write_tile_files
    cfg = stock_config("miseq", phix_match_name="PhiX")
    tile = make_tile(cfg, 4000, seed=91, unknown_frac=0.08, phix_frac=0.5)
    data = str(tmp_path / "data")
    write_tile_files(tile, cfg, data)
    text = open(os.path.join(ROOT, "tests", "golden", "generated_signalproc.ini")).read().replace("@DATADIR@", data)
    text = "\n".join(ln for ln in text.splitlines() if "aybcalls" not in ln) + "\n"      # no third-party base caller here
    outs = {}
    for arm, exe in (("ref", refrun.REF_IMPORT), ("our", os.path.join(BIN, "tailseq-import"))):
        w = tmp_path / arm
        for sub in ("seqqual", "taginfo", "signals", "sigdists-r00", "stats"):
            os.makedirs(str(w / "scratch" / sub))
        (w / "tile.ini").write_text(text.replace("threads = 4", "threads = 1"))
        r = subprocess.run([exe, "tile.ini"], cwd=str(w), capture_output=True, env=dict(os.environ, TSK_SHIM_GZ_LEVEL="1"))
        assert r.returncode == 0, r.stderr.decode()[-2000:]
        files = {}
        for dp, _, fs in os.walk(str(w / "scratch")):
            for f in fs:
                p = os.path.join(dp, f)
                raw = open(p, "rb").read()
                files[os.path.relpath(p, str(w))] = gzip.decompress(raw) if raw[:2] == b"\x1f\x8b" else raw
        outs[arm] = files
    assert set(outs["ref"]) == set(outs["our"]) and len(outs["ref"]) == 3 * 11 + 3
    for k in outs["ref"]:
        if k.endswith(".sigpack"):               # record order is the reference's with threads = 1
            assert outs["ref"][k] == outs["our"][k], k
        else:
            assert outs["ref"][k] == outs["our"][k], k

    # measure-polya-lengths.py: every sample of the tile through one --batch process
    tid = "a1101"
    cpath = str(tmp_path / "cutoffs.txt")
    refrun.write_cutoffs_file(cpath, tid, [0.25] * cfg.total_cycles)
    names = sorted(s.name for s in cfg.samples)
    w = tmp_path / "our"
    os.makedirs(str(w / "scratch" / "taginfo-fl-r01"))
    os.makedirs(str(w / "scratch" / "sigdists-r01"))
    conf = {"polyA_finder": {"signal_analysis_trigger": 8}, "polyA_ruler": {"downhill_extension_weight": 0.49, "signal_resampling_gap": 0.1},
            "polyA_seeder": {"dist_sampling_bins": 1000}}
    params = {"input": [["signals", "scratch/signals/%s_%s.sigpack" % (n, tid)] for n in names] +
                       [["taginfo", "scratch/taginfo/%s_%s.txt.gz" % (n, tid)] for n in names] + [["score_cutoffs", cpath]],
              "output": [["taginfo", "scratch/taginfo-fl-r01/%s_%s.txt.gz" % (n, tid)] for n in names] +
                        [["sigdists", "scratch/sigdists-r01/pos_%s.sigdists" % tid]],
              "threads": 1, "wildcards": [["tile", tid], ["round", "01"]],
              "params": [["CONF", conf], ["BINDIR", BIN], ["BGZIP_CMD", "gzip"]]}
    _run_script("measure-polya-lengths.py", params, str(w))
    total = None
    for n in names:
        a = refrun.run_ruler(tid, str(tmp_path / "ref" / "scratch" / "signals" / ("%s_%s.sigpack" % (n, tid))), cpath,
                             str(tmp_path / "ref" / "scratch" / "taginfo" / ("%s_%s.txt.gz" % (n, tid))), str(tmp_path / "rp.gz"))
        got = gzip.decompress(open(str(w / "scratch" / "taginfo-fl-r01" / ("%s_%s.txt.gz" % (n, tid))), "rb").read())
        assert got == a["taginfo"], n
        total = a["pos"].astype(np.uint64) if total is None else total + a["pos"]
    assert (refrun.read_sigdists(str(w / "scratch" / "sigdists-r01" / ("pos_%s.sigdists" % tid))) == total).all()
