"""Row-wise ingestion and the device-side inflate of gzip-compressed base-call files
(tsk_ingest.cu): a tile handed over cycle by cycle -- some base-call rows as whole
`.bcl.gz` files -- must give exactly the results of the same tile uploaded as arrays, whatever
deflate flavour the files use (stored / fixed / dynamic blocks, several blocks, header with a file
name), also for a read block in the middle of the file; corrupt and short streams must fail loudly."""
import gzip
import io
import struct
import zlib

import numpy as np
import pytest

from helpers import compare_tile

pytestmark = pytest.mark.gpu


def _gz(data: bytes, how: str) -> bytes:
    if how == "gzip6":
        return gzip.compress(data, 6)
    if how == "gzip1":
        return gzip.compress(data, 1)
    if how == "gzip9":
        return gzip.compress(data, 9)
    if how == "stored":
        return gzip.compress(data, 0)
    if how == "named":                                   # FNAME field, like `gzip file`
        buf = io.BytesIO()
        with gzip.GzipFile(filename="s_1_1101.bcl", mode="wb", fileobj=buf, compresslevel=6) as f:
            f.write(data)
        return buf.getvalue()
    if how in ("fixed", "huffman", "rle"):
        strat = {"fixed": zlib.Z_FIXED, "huffman": zlib.Z_HUFFMAN_ONLY, "rle": zlib.Z_RLE}[how]
        c = zlib.compressobj(6, zlib.DEFLATED, 31, 8, strat)
        return c.compress(data) + c.flush()
    if how == "flushed":                                 # many deflate blocks, sync flushes (empty stored blocks) between
        c = zlib.compressobj(6, zlib.DEFLATED, 31)
        out = b""
        for i in range(0, len(data), 7001):
            out += c.compress(data[i:i + 7001]) + c.flush(zlib.Z_SYNC_FLUSH)
        return out + c.flush()
    raise KeyError(how)


HOWS = ["gzip6", "gzip1", "gzip9", "stored", "named", "fixed", "huffman", "rle", "flushed"]


def test_gz_rows_equal_array_upload(stock_cfg, oracle):
    from tailseeker_b200.synth import make_tile
    from tailseeker_b200.importer import TileProcessor, invert_colormatrix
    cfg = stock_cfg
    tile = make_tile(cfg, 60000, seed=17, lowdiv_frac=0.3)            # long repeats: real LZ matches in the streams
    inv = invert_colormatrix(tile.matrix)
    n = tile.nclusters
    gz_rows = {}
    for c in range(cfg.total_cycles):
        if c % 3 != 1:
            raw = _gz(struct.pack("<I", n) + tile.bcl[c].tobytes(), HOWS[c % len(HOWS)])
            assert gzip.decompress(raw)[4:] == tile.bcl[c].tobytes()
            gz_rows[c] = (raw, 4 + n, 4)
    p = TileProcessor(cfg)
    try:
        p.upload_rows(tile.bcl, tile.cif, inv, gz_rows=gz_rows)
        p.process()
        want = oracle.process(cfg, tile.bcl, tile.cif, oracle.invert_matrix(tile.matrix))
        compare_tile(p, cfg, None, want)
        # a read block in the middle of the files: every block inflates the stream from its start
        lo, hi = 20000, 45000
        sub = {c: (raw, size, 4 + lo) for c, (raw, size, skip) in gz_rows.items()}
        p.reset_counters()
        p.upload_rows(tile.bcl[:, lo:hi], tile.cif[:, :, lo:hi], inv, firstclusterno=lo, gz_rows=sub)
        p.process()
        want = oracle.process(cfg, np.ascontiguousarray(tile.bcl[:, lo:hi]), np.ascontiguousarray(tile.cif[:, :, lo:hi]),
                              oracle.invert_matrix(tile.matrix), firstclusterno=lo)
        compare_tile(p, cfg, None, want, firstclusterno=lo)
        # rows that really were taken from the compressed bytes: poison the array rows and compare again
        poisoned = tile.bcl.copy()
        for c in gz_rows:
            poisoned[c] = 0
        p.reset_counters()
        p.upload_rows(poisoned, tile.cif, inv, gz_rows=gz_rows)
        p.process()
        compare_tile(p, cfg, None, oracle.process(cfg, tile.bcl, tile.cif, oracle.invert_matrix(tile.matrix)))
    finally:
        p.close()


def test_staged_rows_equal_array_upload(stock_cfg, oracle):
    """Every row through the library's page-locked staging slots (553 rows over 24 slots: each slot is
    reused ~23 times while earlier copies may still be in flight), twice through the same context, with
    a row length that changes in between (slots grow)."""
    from tailseeker_b200.synth import make_tile
    from tailseeker_b200.importer import TileProcessor, invert_colormatrix
    cfg = stock_cfg
    p = TileProcessor(cfg)
    try:
        for n, seed in ((9000, 5), (31000, 6), (1, 7)):
            tile = make_tile(cfg, n, seed=seed)
            inv = invert_colormatrix(tile.matrix)
            p.reset_counters()
            p.upload_rows(tile.bcl, tile.cif, inv, staged=True)
            # the arrays may change as soon as upload_rows() has returned: the slots hold the data
            want = oracle.process(cfg, tile.bcl.copy(), tile.cif.copy(), oracle.invert_matrix(tile.matrix))
            tile.bcl[:] = 0
            tile.cif[:] = 0
            p.process()
            compare_tile(p, cfg, None, want)
    finally:
        p.close()


def test_bad_streams_fail_loudly(stock_cfg):
    from tailseeker_b200 import _lib
    from tailseeker_b200.synth import make_tile
    from tailseeker_b200.importer import TileProcessor, invert_colormatrix
    cfg = stock_cfg
    tile = make_tile(cfg, 3000, seed=18)
    inv = invert_colormatrix(tile.matrix)
    n = tile.nclusters
    good = gzip.compress(struct.pack("<I", n) + tile.bcl[5].tobytes(), 6)
    cases = {"truncated": good[: len(good) // 2], "not gzip": b"BCL!" + good[4:],
             "flipped": good[:60] + bytes([good[60] ^ 0x5a]) + good[61:],
             "other size": gzip.compress(struct.pack("<I", n) + tile.bcl[5].tobytes()[:-7], 6)}
    p = TileProcessor(cfg)
    try:
        for name, raw in cases.items():
            with pytest.raises(_lib.TskError) as e:
                p.upload_rows(tile.bcl, tile.cif, inv, gz_rows={5: (raw, 4 + n, 4)})
            assert "cycle 6" in str(e.value), name
        # and the context is usable afterwards
        p.upload_rows(tile.bcl, tile.cif, inv, gz_rows={5: (good, 4 + n, 4)})
        p.process()
        assert p.calls().shape[0] == n
    finally:
        p.close()


@pytest.mark.parametrize("bufsize", [None, 3 * 1024 * 1024])
def test_program_inflates_bcl_gz_on_the_device(tmp_path, bufsize):
    """tailseq-import with TSK_DIRECT_INGEST=1 on a run folder whose base-call files are .bcl.gz (two
    thirds of them, assorted compression levels): cycle files go to the device row by row, the
    compressed ones are inflated there -- for one read block and for several -- and every output equals
    the reference binary's on the same files (which inflates with gzread, bclreader.c:60-72,110)."""
    import glob
    import os
    import subprocess
    from oracle import refrun
    from tailseeker_b200.config import stock_config
    from tailseeker_b200.synth import make_tile, write_tile_files
    if not refrun.available():
        pytest.skip("oracle/_ref not built")
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    exe = os.path.join(root, "tailseeker_b200", "bin", "tailseq-import")
    cfg = stock_config("miseq", phix_match_name="PhiX")
    tile = make_tile(cfg, 3700, seed=23, lowdiv_frac=0.2, unknown_frac=0.1, phix_frac=0.5)
    data = str(tmp_path / "data")
    mpath = write_tile_files(tile, cfg, data)
    files = sorted(glob.glob(os.path.join(data, "BaseCalls", "L001", "C*.1", "*.bcl")))
    assert len(files) == cfg.total_cycles
    for k, f in enumerate(files):
        if k % 3 == 2:
            continue
        raw = open(f, "rb").read()
        open(f + ".gz", "wb").write(gzip.compress(raw, (0, 1, 6, 9)[k % 4]))
        os.unlink(f)
    outs = {}
    for arm, binary, env in (("ref", refrun.REF_IMPORT, {}), ("our", exe, {"TSK_DIRECT_INGEST": "1"})):
        w = tmp_path / arm
        for sub in ("seqqual", "taginfo", "signals", "sigdists-r00", "stats"):
            os.makedirs(str(w / sub))
        # (the read-block size depends on [options] threads: parseconfig.c:662-669 -- same value for both
        #  programs when the tile is split, several reader threads for ours otherwise)
        rcfg = cfg.replace(data_dir=data, threep_colormatrix=mpath, threads=1 if (arm == "ref" or bufsize) else 4,
                           read_buffer_size=bufsize if bufsize else 2147483648)
        (w / "tile.ini").write_text(rcfg.to_ini())
        r = subprocess.run([binary, "tile.ini"], cwd=str(w), capture_output=True,
                           env=dict(os.environ, TSK_SHIM_GZ_LEVEL="1", TSK_HOST_TIMING="1", **env))
        assert r.returncode == 0, r.stderr.decode()[-2000:]
        if arm == "our":
            assert b"read CIF/BCL -> device" in r.stderr          # the direct path really ran
            if bufsize:
                assert r.stdout.count(b"Analyzing") >= 3
        got = {}
        for dp, _, fs in os.walk(str(w)):
            for f in fs:
                if f == "tile.ini":
                    continue
                p = os.path.join(dp, f)
                raw = open(p, "rb").read()
                got[os.path.relpath(p, str(w))] = gzip.decompress(raw) if raw[:2] == b"\x1f\x8b" else raw
        outs[arm] = got
    assert set(outs["ref"]) == set(outs["our"])
    differ = sorted(k for k in outs["ref"] if outs["ref"][k] != outs["our"][k])
    assert not differ, differ
    # a damaged file stops the program with the cycle named
    victim = sorted(glob.glob(os.path.join(data, "BaseCalls", "L001", "C*.1", "*.bcl.gz")))[3]
    raw = open(victim, "rb").read()
    open(victim, "wb").write(raw[: len(raw) // 2])
    r = subprocess.run([exe, "tile.ini"], cwd=str(tmp_path / "our"), capture_output=True, env=dict(os.environ, TSK_DIRECT_INGEST="1"))
    assert r.returncode != 0 and b"cycle" in r.stderr
