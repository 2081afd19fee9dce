"""Full-size parity (BASELINE.json configs[1] and configs[2] tile sizes): the CUDA path against
the CPU oracle on a whole 1.07 M-cluster MiSeq tile and a 3.1 M-cluster HiSeq tile -- calls, every
record word, both histograms, counters, ruler lengths and the resampled positive histogram --
and the drop-in programs against the UNMODIFIED reference binaries on a 1.07 M-cluster tile.

Sizes matter here: a 3.1 M-cluster tile has tens of thousands of over-subscribed fair-sampling
buckets, record bases beyond 2^31 bytes, more than 24 000 CTAs per launch and full lists; the
oracle needs about 20 s per million clusters on one host core."""
import os
import shutil
import tempfile

import numpy as np
import pytest

from helpers import compare_tile

pytestmark = pytest.mark.gpu


def _device_tile(cfg, N, seed):
    from tailseeker_b200.importer import invert_colormatrix
    from tailseeker_b200.synth_torch import make_tile_device
    bcl, cif, mtext, pitch = make_tile_device(cfg, N, seed=seed, device="cuda", phix_frac=0.6)
    inv = invert_colormatrix(np.array([np.float32(x) for x in mtext.split()], dtype=np.float32))
    return bcl, cif, mtext, pitch, inv


@pytest.mark.parametrize("layout,N", [("miseq", 1_070_000), ("hiseq", 3_100_000)])
def test_full_tile_against_oracle(oracle, layout, N):
    from tailseeker_b200.config import stock_config
    from tailseeker_b200.importer import TileProcessor
    cfg = stock_config(layout, phix_match_name="PhiX")
    bcl, cif, mtext, pitch, inv = _device_tile(cfg, N, seed=20260)
    h_bcl = np.ascontiguousarray(bcl[:, :N].cpu().numpy())
    h_cif = np.ascontiguousarray(cif[:, :, :N].cpu().numpy())
    p = TileProcessor(cfg)
    try:
        p.upload(bcl.data_ptr(), cif.data_ptr(), inv, nclusters=N, on_device=True, bcl_pitch=pitch, cif_pitch=pitch)
        p.process()
        want = oracle.process(cfg, h_bcl, h_cif, inv)
        calls = compare_tile(p, cfg, None, want)
        # what makes the size matter is really there
        assert int((calls["emitted"] == 1).sum()) > N // 3
        assert want["neg"].sum() > 0 and want["pos"].sum() > 0

        # ruler over the resident records: every sample in one launch vs the oracle per sample
        T = cfg.total_cycles
        rng = np.random.default_rng(5)
        cut = np.clip(0.25 + 0.08 * rng.standard_normal(T), 0.02, 0.9)
        cut[rng.random(T) < 0.03] = np.nan
        packed = p.cutoffs_to_packed(cut)
        p.ruler_reset(T)
        lens, offs = p.ruler_tile_all(packed)
        opos = np.zeros((T, cfg.dist_sampling_bins), dtype=np.uint32)
        for s in range(p.params.num_samples):
            allslots = p.records(s, drop_withdrawn=False)
            assert allslots.shape[0] == offs[s + 1] - offs[s]
            if allslots.shape[0] == 0:
                continue
            live = allslots[:, 1] >> 16 != 0xffff                  # withdrawn slots are not measured
            exp, opos = oracle.ruler(np.ascontiguousarray(allslots[live]), packed, pos=opos)
            got = lens[offs[s]:offs[s + 1]]
            assert (got[live] == exp).all(), "ruler lengths of sample %d" % s
            assert (got[~live] == -1).all()
        assert (p.ruler_hist() == opos).all()
    finally:
        p.close()


def test_full_tile_programs_against_reference_binaries():
    """tailseq-import + tailseq-polya-ruler (drop-in, C host + CUDA) against oracle/_ref's binaries on
    the files of one 1.07 M-cluster tile, threads=1, one read block: identical bytes."""
    from oracle import refrun
    if not refrun.available():
        pytest.skip("oracle/_ref not built")
    from tailseeker_b200.config import stock_config
    from tailseeker_b200.synth import Tile
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    our_import = os.path.join(root, "tailseeker_b200", "bin", "tailseq-import")
    our_ruler = os.path.join(root, "tailseeker_b200", "bin", "tailseq-polya-ruler")
    N = 1_070_000
    cfg = stock_config("miseq", phix_match_name="PhiX")
    bcl, cif, mtext, pitch, inv = _device_tile(cfg, N, seed=777)
    tile = Tile(bcl=np.ascontiguousarray(bcl[:, :N].cpu().numpy()), cif=np.ascontiguousarray(cif[:, :, :N].cpu().numpy()),
                matrix_text=mtext, nclusters=N, truth_sample=None)
    del bcl, cif
    base = tempfile.mkdtemp(prefix="tskfull_", dir="/dev/shm" if os.path.isdir("/dev/shm") else None)
    try:
        wref, wour = os.path.join(base, "ref"), os.path.join(base, "our")
        os.makedirs(wref); os.makedirs(wour)
        big = 4 << 30                                # one read block: 2626 B per cluster
        ref = refrun.run_import(cfg, tile, workdir=wref, read_buffer_size=big)
        shutil.rmtree(os.path.join(wref, "data"))
        our = refrun.run_import(cfg, tile, workdir=wour, read_buffer_size=big, binary=our_import)
        shutil.rmtree(os.path.join(wour, "data"))
        assert ref["stdout"].count("Analyzing") == 1                 # really one block
        assert ref["taginfo"] == our["taginfo"]
        assert ref["seqqual"] == our["seqqual"]
        for k in ref["sigpack"]:
            ta, da, ra = ref["sigpack"][k]
            tb, db, rb = our["sigpack"][k]
            assert (ta, da) == (tb, db) and ra.shape == rb.shape and (ra == rb).all(), k
        assert (ref["pos"] == our["pos"]).all() and (ref["neg"] == our["neg"]).all()
        assert ref["stats_csv"] == our["stats_csv"]

        tid = cfg.tile_id
        cpath = os.path.join(base, "cutoffs.txt")
        refrun.write_cutoffs_file(cpath, tid, [0.25] * cfg.total_cycles)
        for s in (cfg.samples[0], cfg.samples[-1]):
            args = (tid, os.path.join(wref, "signals", "%s_%s.sigpack" % (s.name, tid)), cpath,
                    os.path.join(wref, "taginfo", "%s_%s.txt.gz" % (s.name, tid)))
            a = refrun.run_ruler(*args, os.path.join(base, "ra.gz"))
            b = refrun.run_ruler(*args, os.path.join(base, "rb.gz"), binary=our_ruler)
            assert a["taginfo"] == b["taginfo"], s.name
            assert (a["pos"] == b["pos"]).all(), s.name
    finally:
        shutil.rmtree(base, ignore_errors=True)
