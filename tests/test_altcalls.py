"""CPU: the input side of tailseq-import -- .bcl / .bcl.gz files and [alternative_calls] FASTQ
overrides (reference src/importer/altcalls.c, bclreader.c:185-191, tailseq-import.c:203-230).

Three links, none of which needs a GPU:
  1. the REFERENCE binary run with [alternative_calls] equals the reference binary run without them
     on BCL files that tailseeker_b200.altcalls.overlay() rewrote  -> the Python mirror is pinned;
  2. the C host reader of the drop-in importer (csrc/host/tileio.c, through bin/tsk-tile-dump) hands
     the GPU exactly the mirror's bytes, block after block;
  3. the error behaviour of the reader (short / long / ragged / non-FASTQ files).
"""
import gzip
import os
import shutil
import struct
import subprocess

import numpy as np
import pytest

from oracle import refrun
from tailseeker_b200 import altcalls
from tailseeker_b200.config import stock_config
from tailseeker_b200.synth import make_tile, write_tile_files

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
DUMP = os.path.join(ROOT, "tailseeker_b200", "bin", "tsk-tile-dump")


def _fastq(path, rng, n, length, n_frac=0.02):
    """Random FASTQ with N calls of both kinds ('!' quality and real qualities)."""
    seq = rng.integers(0, 4, (n, length))
    letters = np.frombuffer(b"ACGT", np.uint8)[seq].copy()
    qual = rng.integers(2, 41, (n, length)).astype(np.uint8) + 33
    isn = rng.random((n, length)) < n_frac
    letters[isn] = ord("N")
    qual[isn & (rng.random((n, length)) < 0.5)] = ord("!")
    lower = rng.random((n, length)) < 0.01
    letters[lower & ~isn] |= 0x20
    body = b"".join(b"@x:%d\n%s\n+x:%d\n%s\n" % (i, letters[i].tobytes(), i, qual[i].tobytes())
                    for i in range(n))
    with (gzip.open(path, "wb", compresslevel=1) if path.endswith(".gz") else open(path, "wb")) as f:
        f.write(body)


def _bcl_path(data, cfg, cycle):
    return os.path.join(data, "BaseCalls", "L%03d" % cfg.lane, "C%d.1" % cycle,
                        "s_%d_%04d.bcl" % (cfg.lane, cfg.tile))


def _same(a, b):
    assert a["taginfo"] == b["taginfo"]
    assert a["seqqual"] == b["seqqual"]
    for k in a["sigpack"]:
        ta, da, ra = a["sigpack"][k]
        tb, db, rb = b["sigpack"][k]
        assert (ta, da) == (tb, db) and ra.shape == rb.shape and (ra == rb).all(), k
    assert (a["pos"] == b["pos"]).all() and (a["neg"] == b["neg"]).all()
    assert a["stats_csv"] == b["stats_csv"]


def test_mirror_byte_rule(tmp_path):
    p = str(tmp_path / "a.fastq")
    with open(p, "wb") as f:
        f.write(b"@r0\nACGTNNUacgtu\n+\n!!!!!I#IIIII\n@r1\nTTTTTTTTTTTT\n+\nIIIIIIIIIII~\n")
    c = altcalls.read_fastq_calls(p)
    assert c.shape == (12, 2)
    q = ord("I") - 33
    assert list(c[:, 0]) == [0, 1, 2, 3, 0, q << 2, 3 | (2 << 2), q << 2, 1 | q << 2, 2 | q << 2,
                             3 | q << 2, 3 | q << 2]
    assert c[11, 1] == (3 | ((ord("~") - 33) << 2)) & 0xff          # quality bits are truncated
    with pytest.raises(altcalls.AltCallError, match="Not enough"):
        altcalls.read_fastq_calls(p, 3)
    with pytest.raises(altcalls.AltCallError, match="Extra"):
        altcalls.read_fastq_calls(p, 1)


@pytest.mark.skipif(not refrun.available(), reason="oracle/_ref not built")
def test_reference_with_overrides_equals_reference_on_rewritten_bcl(tmp_path):
    cfg = stock_config("miseq")
    n = 1500
    tile = make_tile(cfg, n, seed=77, lowdiv_frac=0.05, dark_cluster_frac=0.01, spikein_weight=1.0,
                     unknown_frac=0.05)
    rng = np.random.default_rng(5)
    fq = [str(tmp_path / "r5.fastq.gz"), str(tmp_path / "mid.fastq"), str(tmp_path / "tail.fastq.gz")]
    _fastq(fq[0], rng, n, 51)                        # the AYB use case: all of read 5
    _fastq(fq[1], rng, n, 30)                        # overlaps the first entry, the index and read 3
    # read 3 from the delimiter on, unchanged calls with new qualities (keeps poly(A) tails alive)
    t0 = cfg.threep_start - 1 + 20
    keep = tile.bcl[t0:t0 + 60].copy()
    keep[keep == 0] = 1 << 2
    altcalls.write_fastq(fq[2], (keep & 3) | (np.uint8(30) << 2))
    entries = ((1, fq[0]), (40, fq[1]), (t0 + 1, fq[2]))
    want_bcl = altcalls.overlay(tile.bcl, entries)
    assert (want_bcl[39:51] == altcalls.read_fastq_calls(fq[0])[39:51]).all()      # earlier INI entry wins
    assert (want_bcl[51:69] == altcalls.read_fastq_calls(fq[1])[12:30]).all()

    wa, wb = str(tmp_path / "a"), str(tmp_path / "b")
    os.makedirs(wa); os.makedirs(wb)
    a = refrun.run_import(cfg.replace(alternative_calls=entries), tile, workdir=wa)
    assert "Using FASTQ %s for cycles 1-51." % fq[0] in a["stdout"]
    import copy
    tile2 = copy.copy(tile)
    tile2.bcl = want_bcl
    b = refrun.run_import(cfg, tile2, workdir=wb)
    _same(a, b)
    assert a["taginfo"] != refrun.run_import(cfg, tile)["taginfo"]        # the overrides do matter
    # the drop-in's reader on the very same INI: same source messages, in the same order
    r = subprocess.run([DUMP, a["ini"], str(tmp_path / "d.bcl")], capture_output=True, cwd=wa)
    assert r.returncode == 0, r.stderr.decode()
    using = lambda text: [l for l in text.splitlines() if "Using " in l]
    assert using(r.stdout.decode()) == using(a["stdout"]) and len(using(a["stdout"])) == 6


def _dump(cfg, data, mpath, workdir, read_buffer_size=None, threads=1):
    rcfg = cfg.replace(data_dir=data, threep_colormatrix=mpath, threads=threads,
                       read_buffer_size=read_buffer_size or 2147483648)
    ini = os.path.join(workdir, "dump.ini")
    with open(ini, "w") as f:
        f.write(rcfg.to_ini())
    out = os.path.join(workdir, "bcl.out")
    r = subprocess.run([DUMP, ini, out, out + ".cif"], capture_output=True, cwd=workdir)
    blocks, cifs = [], []
    if r.returncode == 0:
        raw = open(out, "rb").read()
        craw = np.fromfile(out + ".cif", "<i2")
        at = cat = 0
        while at < len(raw):
            (count,) = struct.unpack_from("<I", raw, at)
            at += 4
            blocks.append(np.frombuffer(raw, np.uint8, cfg.total_cycles * count, at).reshape(cfg.total_cycles, count))
            at += cfg.total_cycles * count
            k = cfg.threep_length * 4 * count
            cifs.append(craw[cat:cat + k].reshape(cfg.threep_length, 4, count))
            cat += k
    return r, blocks, cifs


def test_host_reader_hands_over_the_mirror_bytes(tmp_path):
    assert os.access(DUMP, os.X_OK), "host programs not built (__graft_entry__.build())"
    cfg = stock_config("miseq")
    n = 3500
    tile = make_tile(cfg, n, seed=78, spikein_weight=1.0)
    data = str(tmp_path / "data")
    mpath = write_tile_files(tile, cfg, data)
    rng = np.random.default_rng(6)
    fq = [str(tmp_path / "r5.fastq.gz"), str(tmp_path / "mid.fastq")]
    _fastq(fq[0], rng, n, 51)
    _fastq(fq[1], rng, n, 30)
    entries = ((1, fq[0]), (40, fq[1]))
    # overridden cycles need no BCL file; others may be gzip'ed (bclreader.c:52-60)
    for c in (1, 2, 45, 69):
        os.remove(_bcl_path(data, cfg, c))
    for c in (70, 308):
        p = _bcl_path(data, cfg, c)
        with open(p, "rb") as f, gzip.open(p + ".gz", "wb") as g:
            shutil.copyfileobj(f, g)
        os.remove(p)
    want = altcalls.overlay(tile.bcl, entries)
    # [options] threads: the cycle files of a block are read side by side
    for bufsize, minblocks, threads in ((None, 1, 1), (4 * 1024 * 1024, 3, 1), (None, 1, 6), (4 * 1024 * 1024, 3, 16)):
        r, blocks, cifs = _dump(cfg.replace(alternative_calls=entries), data, mpath, str(tmp_path), bufsize,
                                threads)
        assert r.returncode == 0, r.stderr.decode()
        assert len(blocks) >= minblocks
        assert (np.concatenate(blocks, axis=1) == want).all()
        assert (np.concatenate(cifs, axis=2) == tile.cif).all()
        out = r.stdout.decode()
        # newest entry first, then the stretches that still come from BCL files
        assert out.index("Using FASTQ %s for cycles 40-69." % fq[1]) < out.index("Using FASTQ %s for cycles 1-51." % fq[0])
        assert "Using BCL base calls for cycle 70-308." in out and out.count("Using BCL base calls") == 1
    # without overrides the removed files are missed
    r, _, _ = _dump(cfg, data, mpath, str(tmp_path))
    assert r.returncode != 0 and b"Can't open file" in r.stderr


@pytest.mark.parametrize("case,message", [
    ("short", b"Not enough sequences"), ("long", b"Extra sequences found"),
    ("ragged", b"Sequence length in the FASTQ is inconsistent"), ("notfastq", b"is not in FASTQ format"),
    ("past_end", b"run past the last cycle"), ("zero", b"Alternative call position is out of range"),
    ("missing", b"Can't open file"), ("empty", b"No sequence available")])
def test_host_reader_errors(tmp_path, case, message):
    cfg = stock_config("miseq")
    n = 40
    tile = make_tile(cfg, n, seed=79)
    data = str(tmp_path / "data")
    mpath = write_tile_files(tile, cfg, data)
    rng = np.random.default_rng(7)
    p = str(tmp_path / "x.fastq")
    first, length, records = 10, 20, n
    if case == "short":
        records = n - 1
    elif case == "long":
        records = n + 1
    elif case == "past_end":
        first = cfg.total_cycles - 5
    elif case == "zero":
        first = 0
    _fastq(p, rng, records, length, n_frac=0)
    if case == "ragged":
        lines = open(p, "rb").read().split(b"\n")
        lines[4 * 7 + 3] = lines[4 * 7 + 3][:-1]
        open(p, "wb").write(b"\n".join(lines))
    elif case == "notfastq":
        lines = open(p, "rb").read().split(b"\n")
        lines[4 * 3 + 2] = b"-"
        open(p, "wb").write(b"\n".join(lines))
    elif case == "missing":
        os.remove(p)
    elif case == "empty":
        open(p, "wb").close()
    r, blocks, _ = _dump(cfg.replace(alternative_calls=((first, p),)), data, mpath, str(tmp_path))
    assert message in r.stderr, r.stderr.decode()
    if case == "zero":
        # the reference's INI parser reports a refused entry and carries on (parseconfig.c:160-164,
        # 695: only an unreadable file aborts), so the entry is dropped
        assert r.returncode == 0 and (np.concatenate(blocks, axis=1) == tile.bcl).all()
        return
    assert r.returncode != 0
    if case not in ("past_end", "zero", "missing", "empty"):
        with pytest.raises(altcalls.AltCallError, match=message.decode()):
            altcalls.overlay(tile.bcl, ((first, p),))
