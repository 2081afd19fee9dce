"""GPU: the drop-in tailseq-import with [alternative_calls] FASTQ overrides and a .bcl.gz cycle
against the reference binary on the same files (src/importer/altcalls.c, bclreader.c:52-60,185-191).
The override is host-side input decoding (csrc/host/tileio.c, pinned on the CPU by
tests/test_altcalls.py); this test closes the loop through the kernels and the writers."""
import gzip
import os
import shutil

import numpy as np
import pytest

from oracle import refrun

pytestmark = [pytest.mark.gpu,
              pytest.mark.skipif(not refrun.available(), reason="oracle/_ref not built")]

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
OUR_IMPORT = os.path.join(ROOT, "tailseeker_b200", "bin", "tailseq-import")


@pytest.mark.parametrize("bufsize", [None, 4 * 1024 * 1024])
def test_importer_with_overrides_matches_reference(tmp_path, bufsize):
    from tailseeker_b200 import altcalls
    from tailseeker_b200.config import stock_config
    from tailseeker_b200.synth import make_tile
    from test_altcalls import _fastq, _same
    cfg = stock_config("miseq")
    n = 3500
    tile = make_tile(cfg, n, seed=91, lowdiv_frac=0.05, dark_cluster_frac=0.01, spikein_weight=1.0,
                     unknown_frac=0.05)
    rng = np.random.default_rng(8)
    fq = [str(tmp_path / "r5.fastq.gz"), str(tmp_path / "mid.fastq"), str(tmp_path / "tail.fastq.gz")]
    _fastq(fq[0], rng, n, 51)
    _fastq(fq[1], rng, n, 30)
    t0 = cfg.threep_start - 1 + 20
    keep = tile.bcl[t0:t0 + 60].copy()
    keep[keep == 0] = 1 << 2
    altcalls.write_fastq(fq[2], (keep & 3) | (np.uint8(30) << 2))
    acfg = cfg.replace(alternative_calls=((1, fq[0]), (40, fq[1]), (t0 + 1, fq[2])))
    wref, wour = str(tmp_path / "ref"), str(tmp_path / "our")
    os.makedirs(wref); os.makedirs(wour)
    ref = refrun.run_import(acfg, tile, workdir=wref, read_buffer_size=bufsize)
    our = refrun.run_import(acfg, tile, workdir=wour, read_buffer_size=bufsize, binary=OUR_IMPORT)
    _same(ref, our)
    for line in ref["stdout"].splitlines():
        if "Using FASTQ" in line or "Using BCL" in line:
            assert line in our["stdout"]


def test_several_tiles_in_one_process(tmp_path):
    """`tailseq-import a.ini b.ini`: the drop-in's way of paying for the CUDA start-up once per lane
    instead of once per tile.  Every tile's files must equal the reference's one-process-per-tile run
    (src/importer/tailseq-import.c:717-757)."""
    import subprocess
    from tailseeker_b200.config import stock_config
    from tailseeker_b200.synth import make_tile
    wref, wour = str(tmp_path / "ref"), str(tmp_path / "our")
    os.makedirs(wref); os.makedirs(wour)
    base = stock_config("miseq", phix_match_name="PhiX")
    inis, cfgs = [], []
    for k, (tileno, n) in enumerate(((1101, 2500), (1102, 300), (2114, 3100))):
        cfg = base.replace(tile=tileno)
        tile = make_tile(cfg, n, seed=100 + k, lowdiv_frac=0.05, dark_cluster_frac=0.01, spikein_weight=1.0,
                         unknown_frac=0.1, phix_frac=0.5)
        ref = refrun.run_import(cfg, tile, workdir=wref)
        ini = str(tmp_path / ("tile%d.ini" % tileno))
        shutil.copy(ref["ini"], ini)
        inis.append(ini); cfgs.append(cfg)
    for sub in ("seqqual", "taginfo", "signals", "sigdists-r00", "stats"):
        os.makedirs(os.path.join(wour, sub), exist_ok=True)
    r = subprocess.run([OUR_IMPORT] + inis, cwd=wour, capture_output=True)
    assert r.returncode == 0, r.stderr.decode()[-2000:]
    assert r.stdout.decode().count("Finished.") == 3
    for cfg in cfgs:
        tid = cfg.tile_id
        for s in cfg.samples:
            for sub in ("taginfo", "seqqual"):
                name = os.path.join(sub, "%s_%s.txt.gz" % (s.name, tid))
                with gzip.open(os.path.join(wref, name), "rb") as a, gzip.open(os.path.join(wour, name), "rb") as b:
                    assert a.read() == b.read(), name
            name = os.path.join("signals", "%s_%s.sigpack" % (s.name, tid))
            ta, da, ra = refrun.read_sigpack(os.path.join(wref, name))
            tb, db, rb = refrun.read_sigpack(os.path.join(wour, name))
            assert (ta, da) == (tb, db) and ra.shape == rb.shape and (ra == rb).all(), name
        for kind in ("pos", "neg"):
            name = os.path.join("sigdists-r00", "%s_%s.sigdists" % (kind, tid))
            assert (refrun.read_sigdists(os.path.join(wref, name)) ==
                    refrun.read_sigdists(os.path.join(wour, name))).all(), name
        name = os.path.join("stats", "signal-proc-%s.csv" % tid)
        assert open(os.path.join(wref, name)).read() == open(os.path.join(wour, name)).read()
    # a failing tile stops the chain with the reference's exit status
    r = subprocess.run([OUR_IMPORT, str(tmp_path / "missing.ini"), inis[0]], cwd=wour, capture_output=True)
    assert r.returncode == 255 and b"Finished." not in r.stdout
