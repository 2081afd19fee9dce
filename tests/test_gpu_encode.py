"""Device-side output encoding (tsk_encode.cu): the text the GPU formats and the
BGZF blocks it deflates must decompress to exactly the bytes of the host writers' mirror
(tailseeker_b200/textio.py, pinned to the reference binary by test_gpu_dropin.py / test_oracle_*),
every block must be a well-formed BGZF member, and the drop-in program (which writes these
blocks) must still equal the reference binary's decompressed files."""
import gzip
import struct
import zlib

import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def bgzf_blocks(data: bytes):
    """Walks the BGZF members of `data` by their BSIZE fields; yields (isize, payload bytes).
    Checks header constants, the BC subfield, CRC-32 and ISIZE of every member."""
    at, out = 0, []
    while at < len(data):
        assert data[at:at + 4] == b"\x1f\x8b\x08\x04", "not a gzip member with an extra field at %d" % at
        xlen, = struct.unpack_from("<H", data, at + 10)
        assert xlen == 6 and data[at + 12:at + 16] == b"BC\x02\x00"
        bsize, = struct.unpack_from("<H", data, at + 16)
        member = data[at:at + bsize + 1]
        assert len(member) == bsize + 1 <= 65536
        raw = zlib.decompress(member[18:-8], wbits=-15)
        crc, isize = struct.unpack("<II", member[-8:])
        assert isize == len(raw) and crc == (zlib.crc32(raw) & 0xffffffff)
        out.append(raw)
        at += bsize + 1
    return out


def _expected(cfg, proc, tile, calls, firstclusterno=0):
    from tailseeker_b200 import textio
    seq, qual = textio.decode_basecalls(tile.bcl)
    want = {}
    for s in cfg.sample_list():
        i = s["numindex"]
        if not s["has_streams"]:
            want[(i, 0)] = want[(i, 1)] = want[(i, 2)] = b""
            continue
        want[(i, 0)] = textio.format_seqqual(cfg, s, calls, seq, qual, firstclusterno)
        want[(i, 1)] = textio.format_taginfo(cfg, s, calls, seq, firstclusterno)
        want[(i, 2)] = proc.records(i).astype("<u4").tobytes()          # withdrawn slots dropped
    return want


@pytest.mark.parametrize("n,seed,kw,first", [
    (5000, 3, {}, 0),
    (1, 4, {}, 7),
    (129, 5, dict(unknown_frac=0.5), 123456),
    (40000, 6, dict(dark_cluster_frac=0.05, dark_cycle_frac=0.02, lowdiv_frac=0.05), 4000000000),
])
def test_encoded_streams_match_host_writers(stock_cfg, n, seed, kw, first):
    from tailseeker_b200.synth import make_tile
    from tailseeker_b200.importer import TileProcessor, invert_colormatrix
    cfg = stock_cfg
    tile = make_tile(cfg, n, seed=seed, **kw)
    p = TileProcessor(cfg)
    try:
        p.encode_configure()
        p.upload(tile.bcl, tile.cif, invert_colormatrix(tile.matrix), firstclusterno=first)
        p.process()
        got = p.encode()
        calls = p.calls()
        want = _expected(cfg, p, tile, calls, first)
        nwithdrawn = 0
        for key, w in want.items():
            blocks = bgzf_blocks(got[key])
            assert b"".join(blocks) == w, "stream %s" % (key,)
            assert gzip.decompress(got[key]) == w if got[key] else w == b""
            assert all(len(b) == 32768 for b in blocks[:-1])
            if key[1] == 2:
                nwithdrawn += p.records(key[0], drop_withdrawn=False).shape[0] - p.records(key[0]).shape[0]
        if kw.get("dark_cluster_frac"):
            assert nwithdrawn > 0          # the squeeze of withdrawn slots was exercised
        # compression: Huffman-only deflate is no worse than a few percent above zlib level 1
        tot_in = sum(len(w) for w in want.values())
        tot_out = sum(len(g) for g in got.values())
        if n >= 5000:
            lvl1 = sum(len(zlib.compress(w, 1)) for w in want.values())
            assert tot_out < 1.08 * lvl1 and tot_out < 0.8 * tot_in
        assert p.bgzf_eof() == bytes.fromhex("1f8b08040000000000ff0600424302001b0003000000000000000000")
        # a second tile through the same context
        p.upload(tile.bcl[:, : n // 2 + 1], tile.cif[:, :, : n // 2 + 1], invert_colormatrix(tile.matrix))
        p.process()
        got2 = p.encode()
        sub = type(tile)(bcl=tile.bcl[:, : n // 2 + 1], cif=tile.cif[:, :, : n // 2 + 1], matrix_text=tile.matrix_text,
                         nclusters=n // 2 + 1, truth_sample=None)
        want2 = _expected(cfg, p, sub, p.calls(), 0)
        for key, w in want2.items():
            assert b"".join(bgzf_blocks(got2[key])) == w
    finally:
        p.close()


def test_deflate_kernel_on_arbitrary_bytes(stock_cfg):
    """The deflate / CRC / gather kernels on inputs the tile path rarely produces: empty, one byte, one
    symbol only, random bytes (stored fallback), Fibonacci-skewed frequencies (Huffman depths beyond
    the 15-bit limit), every length around the piece size, and a long mixed stream."""
    from tailseeker_b200.importer import TileProcessor
    rng = np.random.default_rng(11)
    fib = [1, 1]
    while len(fib) < 22:
        fib.append(fib[-1] + fib[-2])
    skew = np.concatenate([np.full(f, i, dtype=np.uint8) for i, f in enumerate(fib)])
    rng.shuffle(skew)
    cases = [b"", b"x", b"A" * 70000, rng.integers(0, 256, 100000, dtype=np.uint8).tobytes(), skew.tobytes()[:32768],
             bytes(range(256)) * 3, rng.integers(0, 4, 200001, dtype=np.uint8).tobytes()]
    for ln in (32767, 32768, 32769, 65535, 65536, 65537, 127, 128, 129):
        cases.append((b"ACGTTTTTTTTTTTTTNACGGT\t!#5:<\n" * (ln // 30 + 1))[:ln])
    cases.append(b"".join(cases))
    p = TileProcessor(stock_cfg)
    try:
        for i, data in enumerate(cases):
            z = p.selftest_deflate(data)
            blocks = bgzf_blocks(z)
            assert b"".join(blocks) == data, "case %d (%d bytes)" % (i, len(data))
            assert len(blocks) == (len(data) + 32767) // 32768
            if len(data):
                assert gzip.decompress(z) == data
        z = p.selftest_deflate(cases[3])                       # random bytes: stored, 31 bytes of framing per piece
        assert len(z) == len(cases[3]) + 31 * ((len(cases[3]) + 32767) // 32768)
    finally:
        p.close()
