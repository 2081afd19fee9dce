"""GPU: the drop-in programs (tailseeker_b200/bin/tailseq-import, tailseq-polya-ruler: C host +
CUDA library) against the reference binaries on the same files, same INI, same argv.
Every output file must decompress to identical bytes."""
import os

import numpy as np
import pytest

from oracle import refrun

pytestmark = [pytest.mark.gpu,
              pytest.mark.skipif(not refrun.available(), reason="oracle/_ref not built")]

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
OUR_IMPORT = os.path.join(ROOT, "tailseeker_b200", "bin", "tailseq-import")
OUR_RULER = os.path.join(ROOT, "tailseeker_b200", "bin", "tailseq-polya-ruler")


def _same(a, b):
    assert a["taginfo"] == b["taginfo"]
    assert a["seqqual"] == b["seqqual"]
    for k in a["sigpack"]:
        ta, da, ra = a["sigpack"][k]
        tb, db, rb = b["sigpack"][k]
        assert (ta, da) == (tb, db) and ra.shape == rb.shape and (ra == rb).all(), k
    assert (a["pos"] == b["pos"]).all() and (a["neg"] == b["neg"]).all()
    assert a["stats_csv"] == b["stats_csv"]


@pytest.mark.parametrize("control,bufsize", [("", None), ("PhiX", None), ("", 4 * 1024 * 1024)])
def test_programs_match_reference(tmp_path, control, bufsize):
    from tailseeker_b200.config import stock_config
    from tailseeker_b200.synth import make_tile
    assert os.access(OUR_IMPORT, os.X_OK), "host programs not built (__graft_entry__.build())"
    cfg = stock_config("miseq", phix_match_name=control)
    n = 3000 if bufsize is None else 3500          # 4 MiB buffer -> several read blocks
    tile = make_tile(cfg, n, seed=55, lowdiv_frac=0.05, dark_cluster_frac=0.01, spikein_weight=1.0,
                     unknown_frac=0.15 if control else 0.03, phix_frac=0.6 if control else 0.0)
    wref, wour = str(tmp_path / "ref"), str(tmp_path / "our")
    os.makedirs(wref); os.makedirs(wour)
    ref = refrun.run_import(cfg, tile, workdir=wref, read_buffer_size=bufsize)
    our = refrun.run_import(cfg, tile, workdir=wour, read_buffer_size=bufsize, binary=OUR_IMPORT)
    _same(ref, our)
    if bufsize is not None:
        import re
        nref = re.findall(r"#(\d+)/(\d+)\] Analyzing", ref["stdout"])
        nour = re.findall(r"#(\d+)/(\d+)\] Analyzing", our["stdout"])
        assert nref == nour and len(nref) >= 3          # same multi-block split as the reference

    # ruler: argv and outputs of polyaruler.c
    tid = cfg.tile_id
    rng = np.random.default_rng(1)
    cut = np.clip(0.25 + 0.1 * rng.standard_normal(cfg.total_cycles), 0.01, 0.9)
    cut[rng.random(cfg.total_cycles) < 0.05] = np.nan
    cpath = str(tmp_path / "cutoffs.txt")
    refrun.write_cutoffs_file(cpath, tid, [float(c) for c in cut])
    for s in cfg.samples[:3] + cfg.samples[-2:]:
        args = (tid, os.path.join(wref, "signals", "%s_%s.sigpack" % (s.name, tid)), cpath,
                os.path.join(wref, "taginfo", "%s_%s.txt.gz" % (s.name, tid)))
        a = refrun.run_ruler(*args, str(tmp_path / "ra.gz"))
        b = refrun.run_ruler(*args, str(tmp_path / "rb.gz"), binary=OUR_RULER)
        assert a["taginfo"] == b["taginfo"], s.name
        assert (a["pos"] == b["pos"]).all(), s.name

    # --batch: every sample of the tile on one CUDA context (not in the reference)
    import subprocess
    jobs, want = [], {}
    for s in cfg.samples:
        args = (tid, os.path.join(wref, "signals", "%s_%s.sigpack" % (s.name, tid)), cpath, "8", "0.49",
                os.path.join(wref, "taginfo", "%s_%s.txt.gz" % (s.name, tid)), "1000", "0.1",
                str(tmp_path / ("bpos_%s.gz" % s.name)), str(tmp_path / ("btag_%s.txt" % s.name)))
        jobs.append("\t".join(args))
        want[s.name] = refrun.run_ruler(*args[:3], args[5], str(tmp_path / "rc.gz"))
    (tmp_path / "jobs.tsv").write_text("\n".join(jobs) + "\n")
    r = subprocess.run([OUR_RULER, "--batch", str(tmp_path / "jobs.tsv")], capture_output=True)
    assert r.returncode == 0, r.stderr.decode()
    for s in cfg.samples:
        assert (tmp_path / ("btag_%s.txt" % s.name)).read_bytes() == want[s.name]["taginfo"], s.name
        assert (refrun.read_sigdists(str(tmp_path / ("bpos_%s.gz" % s.name))) == want[s.name]["pos"]).all(), s.name


@pytest.mark.parametrize("bufsize", [None, 4 * 1024 * 1024])
def test_host_array_path_matches_reference(tmp_path, monkeypatch, bufsize):
    """TSK_DIRECT_INGEST=0: the blocks are read into whole-block host arrays and uploaded with one
    tsk_tile_upload() each (the default sends every cycle file from a page-locked staging slot as it is
    read); same outputs."""
    from tailseeker_b200.config import stock_config
    from tailseeker_b200.synth import make_tile
    cfg = stock_config("miseq", phix_match_name="PhiX")
    tile = make_tile(cfg, 3300, seed=56, lowdiv_frac=0.05, dark_cluster_frac=0.01, spikein_weight=1.0,
                     unknown_frac=0.15, phix_frac=0.6)
    wref, wour = str(tmp_path / "ref"), str(tmp_path / "our")
    os.makedirs(wref); os.makedirs(wour)
    ref = refrun.run_import(cfg, tile, workdir=wref, read_buffer_size=bufsize)
    monkeypatch.setenv("TSK_DIRECT_INGEST", "0")
    monkeypatch.setenv("TSK_HOST_TIMING", "1")
    our = refrun.run_import(cfg, tile, workdir=wour, read_buffer_size=bufsize, binary=OUR_IMPORT)
    _same(ref, our)
    assert "read CIF/BCL -> device" not in our["stderr"] and "read CIF/BCL" in our["stderr"]


def test_ruler_exit_codes(tmp_path):
    import subprocess
    assert subprocess.run([OUR_RULER], capture_output=True).returncode == 1
    r = subprocess.run([OUR_RULER, "a1101", "/nonexistent.sigpack", "/nonexistent.txt", "8", "0.49",
                        "/nonexistent.gz", "1000", "0.1", str(tmp_path / "o.gz")], capture_output=True)
    assert r.returncode == 1          # cutoffs unreadable (polyaruler.c:421-423)
