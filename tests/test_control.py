"""PhiX control classification (controlaligner.c:107-145): the oracle's restatement against the
reference's own aligner (oracle/_ref/libtskref_control.so = controlaligner.c + ssw.c +
phix_control.c behind oracle/ref_control_shim.c), against committed golden reads, and the
whole importer with a [control] section against the reference binary."""
import ctypes
import os

import numpy as np
import pytest

from oracle import refrun

HERE = os.path.dirname(os.path.abspath(__file__))
GOLDEN = os.path.join(HERE, "golden", "control_reads.npz")
REFLIB = os.path.join(os.path.dirname(refrun.REF_IMPORT), "libtskref_control.so")
LETTERS = np.frombuffer(b"ACGTN", dtype=np.uint8)


def control_reads(seed: int, count: int, length: int = 64) -> np.ndarray:
    """`count` reads of `length` letters: random, verbatim PhiX, and PhiX with substitutions,
    indels and no-calls spread so that alignment scores cover both sides of the threshold."""
    from tailseeker_b200.synth import phix_genome
    rng = np.random.default_rng(seed)
    fwd = phix_genome()
    rev = (3 - fwd)[::-1]
    out = np.empty((count, length), dtype=np.uint8)
    for i in range(count):
        if rng.random() < 0.15:
            r = rng.integers(0, 4, length + 8).astype(np.uint8)
        else:
            st = fwd if rng.random() < 0.5 else rev
            s0 = int(rng.integers(0, st.size - length - 8))
            r = st[s0:s0 + length + 8].copy()
            if rng.random() < 0.4:
                at, ln = int(rng.integers(6, 44)), int(rng.integers(1, 6))
                r = np.concatenate([r[:at], rng.integers(0, 4, ln).astype(np.uint8), r[at:]]) \
                    if rng.random() < 0.5 else np.concatenate([r[:at], r[at + ln:]])
            ns = int(rng.integers(0, 9))
            w = rng.integers(0, 50, ns)
            r[w] = (r[w] + rng.integers(1, 4, ns)) & 3
        r = r[:length].copy()
        if rng.random() < 0.3:
            r[rng.integers(0, 50, int(rng.integers(1, 4)))] = 4
        out[i] = r
    return LETTERS[out]


def oracle_scores(oracle, reads, first, length):
    L = oracle.lib()
    L.tsk_oracle_control_score.restype = ctypes.c_int
    L.tsk_oracle_control_score.argtypes = [ctypes.c_char_p, ctypes.c_int]
    return np.array([L.tsk_oracle_control_score(bytes(r[first:first + length]), length) for r in reads])


@pytest.mark.skipif(not os.path.exists(REFLIB), reason="oracle/_ref/libtskref_control.so not built")
@pytest.mark.parametrize("first,length", [(5, 40), (0, 32), (3, 50)])
def test_oracle_score_equals_reference_ssw(oracle, first, length):
    ref = ctypes.CDLL(REFLIB)
    assert ref.tskref_control_setup(first, length) == int(length * 0.65)
    reads = control_reads(100 + length, 1500)
    want = np.array([ref.tskref_control_score(bytes(r)) for r in reads])
    got = oracle_scores(oracle, reads, first, length)
    assert (got == want).all()
    assert (want >= int(length * 0.65)).sum() > 200 and (want < int(length * 0.65)).sum() > 200


def test_oracle_score_golden(oracle):
    """Scores and decisions of the reference aligner on committed reads (tests/make_golden.py)."""
    g = np.load(GOLDEN)
    first, length = int(g["first"]), int(g["length"])
    got = oracle_scores(oracle, g["reads"], first, length)
    assert (got == g["score"]).all()
    L = oracle.lib()
    L.tsk_oracle_control_try.argtypes = [ctypes.c_char_p, ctypes.c_int, ctypes.c_int]
    decided = np.array([L.tsk_oracle_control_try(bytes(r), first, length) for r in g["reads"]])
    assert (decided == g["decision"]).all()
    assert 500 < int(g["decision"].sum()) < len(decided) - 500


def test_importer_with_control_golden(oracle):
    """Reference tailseq-import on a tile whose unassigned clusters are half PhiX: demultiplexing
    statistics (PhiX / Unknown rows) and per-cluster sample assignment of the oracle."""
    from tailseeker_b200 import textio
    from tailseeker_b200.config import stock_config
    from tailseeker_b200.synth import make_tile
    g = np.load(GOLDEN)
    cfg = stock_config("miseq", phix_match_name="PhiX")
    tile = make_tile(cfg, int(g["tile_n"]), seed=int(g["tile_seed"]), unknown_frac=0.2, phix_frac=0.6)
    o = oracle.process(cfg, tile.bcl, tile.cif, oracle.invert_matrix(tile.matrix))
    assert textio.format_stats_csv(cfg, o["counters"]) == str(g["stats_csv"])
    assert o["counters"][1, 0] > 100 and o["counters"][0, 0] > 100


@pytest.mark.skipif(not refrun.available(), reason="oracle/_ref not built")
def test_importer_with_control_live(oracle):
    from tailseeker_b200 import textio
    from tailseeker_b200.config import stock_config
    from tailseeker_b200.synth import make_tile
    cfg = stock_config("miseq", phix_match_name="PhiX", phix_match_start=4, phix_match_length=36)
    tile = make_tile(cfg, 2500, seed=91, unknown_frac=0.25, phix_frac=0.7)
    r = refrun.run_import(cfg, tile)
    o = oracle.process(cfg, tile.bcl, tile.cif, oracle.invert_matrix(tile.matrix))
    assert textio.format_stats_csv(cfg, o["counters"]) == r["stats_csv"]
    assert (o["pos"] == r["pos"]).all() and (o["neg"] == r["neg"]).all()
