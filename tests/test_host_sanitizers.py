"""AddressSanitizer + UBSan builds (make -C tailseeker_b200/csrc/host asan) of the host code that runs
without a GPU -- INI reader, configuration parser, CIF / BCL / .bcl.gz / FASTQ-override readers,
block split -- driven over the awkward INI file and a small tile.  Skipped where the toolchain has no
libasan (the round-2 build image: `ld: cannot find -lasan`)."""
import os
import subprocess

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HOST = os.path.join(ROOT, "tailseeker_b200", "csrc", "host")
ASAN = os.path.join(ROOT, "tailseeker_b200", "bin", "asan")


@pytest.fixture(scope="module")
def asan_tools():
    r = subprocess.run(["make", "-s", "-C", HOST, "asan"], capture_output=True)
    if r.returncode != 0:
        pytest.skip("no sanitizer runtime in this toolchain: " + r.stderr.decode().strip().splitlines()[0])
    return ASAN


def test_ini_reader_under_sanitizers(asan_tools, tmp_path):
    from test_ini import AWKWARD
    p = tmp_path / "a.ini"
    p.write_bytes(AWKWARD)
    r = subprocess.run([os.path.join(asan_tools, "tsk-ini-dump"), str(p)], capture_output=True)
    assert r.returncode == 0 and b"ERROR" not in r.stderr, r.stderr.decode()[-2000:]


def test_tile_readers_under_sanitizers(asan_tools, tmp_path):
    from tailseeker_b200.config import stock_config
    from tailseeker_b200.synth import make_tile, write_tile_files
    cfg = stock_config("miseq", phix_match_name="PhiX")
    tile = make_tile(cfg, 700, seed=5)
    data = str(tmp_path / "data")
    mpath = write_tile_files(tile, cfg, data)
    ini = tmp_path / "t.ini"
    ini.write_text(cfg.replace(data_dir=data, threep_colormatrix=mpath, read_buffer_size=600000).to_ini())
    r = subprocess.run([os.path.join(asan_tools, "tsk-tile-dump"), str(ini), str(tmp_path / "bcl.out"), str(tmp_path / "cif.out")],
                       capture_output=True, cwd=str(tmp_path))
    assert r.returncode == 0 and b"ERROR" not in r.stderr, r.stderr.decode()[-2000:]
    assert os.path.getsize(str(tmp_path / "bcl.out")) > 700 * cfg.total_cycles
