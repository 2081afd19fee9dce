"""GPU parity tests proper: the CUDA path, called through the C ABI, against the CPU oracle
on the same seeded inputs.  Bar: bit-exact calls, flags, record packets, histograms, counters."""
import numpy as np
import pytest

from helpers import compare_tile

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def proc(stock_cfg):
    from tailseeker_b200.importer import TileProcessor
    p = TileProcessor(stock_cfg)
    yield p
    p.close()


def _run(proc, cfg, oracle, n, seed, **tilekw):
    from tailseeker_b200.synth import make_tile
    from tailseeker_b200.importer import invert_colormatrix
    tile = make_tile(cfg, n, seed=seed, **tilekw)
    inv = invert_colormatrix(tile.matrix)
    assert (inv == oracle.invert_matrix(tile.matrix)).all()
    proc.reset_counters()
    proc.upload(tile.bcl, tile.cif, inv)
    proc.process()
    o = oracle.process(cfg, tile.bcl, tile.cif, inv)
    return tile, o


def test_logf_exact(proc, oracle):
    """Device logf == host libm logf, bit for bit, on (0,1]: dense near 1, every binade,
    subnormals, and 4M pseudo-random floats (SURVEY appendix C re-checked on this box)."""
    rng = np.random.default_rng(7)
    bits = [rng.integers(1, 0x3f800001, size=1 << 22, dtype=np.uint32),
            np.arange(0x3f800000 - (1 << 18), 0x3f800001, dtype=np.uint32),
            np.arange(1, 1 << 14, dtype=np.uint32),
            (np.arange(1, 128, dtype=np.uint32) << 23)[:, None].repeat(64, 1).ravel()
            + rng.integers(0, 1 << 23, size=127 * 64, dtype=np.uint32)]
    x = np.concatenate(bits).astype(np.uint32).view(np.float32)
    got = proc.selftest_logf(x)
    import ctypes
    want = np.empty_like(x)
    lib = oracle.lib()
    # vectorised host logf through the oracle's exported libm wrapper
    f = lib.tsk_oracle_logf
    step = 1 << 16
    for i in range(0, x.size, step):   # keep the python loop cheap: sample every 8th value
        blk = x[i:i + step:8]
        want[i:i + step:8] = [f(ctypes.c_float(float(v))) for v in blk]
    sel = np.zeros(x.size, dtype=bool)
    sel[::8] = True
    assert (got[sel].view(np.uint32) == want[sel].view(np.uint32)).all()


def test_fast_arithmetic_sweeps(proc):
    """(1) FMA logf == separately-rounded logf on every normal float in (0,1];
    (2) reciprocal-multiply + one residual correction == IEEE division on 2^31 operand pairs
    (tools/div_onestep_proof.c holds the exhaustive hard-case search behind that step)."""
    logf_bad, div_bad = proc.selftest_sweeps(1 << 31)
    assert logf_bad == 0
    assert div_bad == 0


def test_generic_kernel_matches_oracle(stock_cfg, oracle):
    """The literal operation-by-operation scoring kernel (the fast kernel's fallback)."""
    from tailseeker_b200.importer import TileProcessor
    p = TileProcessor(stock_cfg)
    try:
        p.force_generic(True)
        _, o = _run(p, stock_cfg, oracle, 9000, 21, dark_cluster_frac=0.03, lowdiv_frac=0.05)
        compare_tile(p, stock_cfg, None, o)
    finally:
        p.close()


def test_row_form_of_fast_kernel(stock_cfg, oracle):
    """The fast kernel's row form (warp = 32 consecutive clusters, resident balancer cycles); the
    default is the compact form over the work list."""
    from tailseeker_b200.importer import TileProcessor
    p = TileProcessor(stock_cfg)
    try:
        p.force_generic(2)
        _, o = _run(p, stock_cfg, oracle, 9000, 22, dark_cluster_frac=0.03, lowdiv_frac=0.05)
        compare_tile(p, stock_cfg, None, o)
    finally:
        p.close()


@pytest.mark.parametrize("unknown_frac", [0.3, 0.5, 0.9])
def test_sparse_work_list(proc, stock_cfg, oracle, unknown_frac):
    """Many dropped clusters: 0.3 mixes work items of the narrow (64-cluster box) and the wide (128)
    instance of the scoring kernel, 0.5 is all wide, at 0.9 items span more clusters than the wide box
    and hand their far lanes to the literal kernel."""
    _, o = _run(proc, stock_cfg, oracle, 15000, 23, unknown_frac=unknown_frac)
    compare_tile(proc, stock_cfg, None, o)


def test_degenerate_bandwidth_handover(stock_cfg, oracle):
    """Clusters whose balancer cycles are constant (zero signal bandwidth -> division by zero,
    NaN/inf normalised signals) leave the fast kernel for the generic one; results still
    equal the oracle's."""
    from tailseeker_b200.synth import make_tile
    from tailseeker_b200.importer import TileProcessor, invert_colormatrix
    tile = make_tile(stock_cfg, 5000, seed=31)
    rng = np.random.default_rng(5)
    flat = rng.choice(5000, size=400, replace=False)
    tile.cif[:20, :, flat[:200]] = 0                      # all-zero balancer: bandwidth 0 on all channels
    tile.cif[:20, 1, flat[200:]] = 77                     # one constant channel
    inv = invert_colormatrix(tile.matrix)
    p = TileProcessor(stock_cfg)
    try:
        p.upload(tile.bcl, tile.cif, inv)
        p.process()
        o = oracle.process(stock_cfg, tile.bcl, tile.cif, inv)
        compare_tile(p, stock_cfg, None, o)
    finally:
        p.close()


@pytest.mark.parametrize("n,seed", [(20000, 1), (4099, 2), (129, 3), (1, 4)])
def test_tile_matches_oracle(proc, stock_cfg, oracle, n, seed):
    _, o = _run(proc, stock_cfg, oracle, n, seed)
    compare_tile(proc, stock_cfg, None, o)


def test_empty_tile(proc, stock_cfg):
    T, L3 = stock_cfg.total_cycles, stock_cfg.threep_length
    proc.reset_counters()
    proc.upload(np.zeros((T, 0), np.uint8), np.zeros((L3, 4, 0), np.int16), np.eye(4, dtype=np.float32))
    proc.process()
    assert proc.calls().size == 0
    assert proc.counters().sum() == 0
    pos, neg = proc.hist()
    assert pos.sum() == 0 and neg.sum() == 0


def test_dark_and_lowdiversity_heavy(proc, stock_cfg, oracle):
    """Many dark clusters (withdrawn record slots, undo of inline negative sampling) and many
    identical inserts (over-subscribed fair-sampling buckets)."""
    _, o = _run(proc, stock_cfg, oracle, 12000, 11, dark_cluster_frac=0.08, dark_cycle_frac=0.01,
                lowdiv_frac=0.25)
    calls = compare_tile(proc, stock_cfg, None, o)
    assert (calls["flags"] & 0x800).sum() > 50


def test_keep_options_and_hiseq_layout(oracle):
    from tailseeker_b200.config import stock_config
    from tailseeker_b200.importer import TileProcessor
    cfg = stock_config("hiseq", keep_no_delimiter=1, keep_low_quality_balancer=1)
    p = TileProcessor(cfg)
    try:
        _, o = _run(p, cfg, oracle, 6000, 5)
        compare_tile(p, cfg, None, o)
    finally:
        p.close()


def test_nondefault_parameters(oracle):
    """Other trigger lengths / weights / sampling parameters than conf/defaults.conf."""
    from tailseeker_b200.config import stock_config
    from tailseeker_b200.importer import TileProcessor
    cfg = stock_config("miseq", seed_trigger_polya_length=6, signal_analysis_trigger=12,
                       negative_sample_polya_length=5, minimum_polya_length=3,
                       maximum_modifications=10, polya_boundary_pos=60, polya_sampling_gap=5,
                       max_cctr_scan_left_space=10, max_cctr_scan_right_space=15,
                       fair_sampling_max_count=2, fair_sampling_hash_space_size=1000003,
                       dist_sampling_bins=500, maximum_dark_cycles=3, dark_cycles_threshold=25,
                       balancer_num_positive_samples=3, balancer_num_negative_samples=5,
                       t_intensity_k=20.0, t_intensity_center=0.75)
    p = TileProcessor(cfg)
    try:
        _, o = _run(p, cfg, oracle, 8000, 6, lowdiv_frac=0.05)
        compare_tile(p, cfg, None, o)
    finally:
        p.close()


@pytest.mark.parametrize("buckets,maxcount,mode", [(64, 5, 0), (64, 5, 2), (7, 3, 0), (4096, 1, 0), (4096, 1, 1)])
def test_fair_sampling_contention(oracle, buckets, maxcount, mode):
    """A tiny hash space: nearly every negative candidate sits in an over-subscribed bucket, so the
    first `maxcount` of each bucket in cluster order are found by the slot walk and sampled
    afterwards -- from the scratch columns the compact kernel keeps (mode 0), or re-scored from
    the intensities (row form, literal kernel)."""
    from tailseeker_b200.config import stock_config
    from tailseeker_b200.importer import TileProcessor
    cfg = stock_config("miseq", fair_sampling_hash_space_size=buckets, fair_sampling_max_count=maxcount)
    p = TileProcessor(cfg)
    try:
        p.force_generic(mode)
        _, o = _run(p, cfg, oracle, 9000, 31, lowdiv_frac=0.05, dark_cluster_frac=0.02)
        compare_tile(p, cfg, None, o)
    finally:
        p.close()


def test_wide_seed_weights(oracle):
    """find_polya weights that do not fit the packed int8 form of the seed scan."""
    from tailseeker_b200.config import stock_config
    from tailseeker_b200.importer import TileProcessor
    cfg = stock_config("miseq", polyA_weights=dict(T=200, A=-900, C=-900, G=-900, N=-100),
                       nonA_weights=dict(T=-100, A=0, C=-400, G=-400, N=0))
    p = TileProcessor(cfg)
    try:
        _, o = _run(p, cfg, oracle, 5000, 7)
        compare_tile(p, cfg, None, o)
    finally:
        p.close()


@pytest.mark.parametrize("minlen,maxmods", [(1, 5), (9, 30), (12, 30), (5, 0)])
def test_seed_finder_shapes(oracle, minlen, maxmods):
    """find_polya with other minimum lengths / modification limits: the shift-register form of
    the seed scan (minimum length <= 9) and its array form (longer)."""
    from tailseeker_b200.config import stock_config
    from tailseeker_b200.importer import TileProcessor
    cfg = stock_config("miseq", minimum_polya_length=minlen, maximum_modifications=maxmods)
    p = TileProcessor(cfg)
    try:
        _, o = _run(p, cfg, oracle, 6000, 40 + minlen)
        compare_tile(p, cfg, None, o)
    finally:
        p.close()


def test_firstclusterno_and_pitch(stock_cfg, oracle):
    """Second read block of a tile (global cluster numbers) and padded row pitches."""
    from tailseeker_b200.synth import make_tile
    from tailseeker_b200.importer import TileProcessor, invert_colormatrix
    tile = make_tile(stock_cfg, 3000, seed=8)
    inv = invert_colormatrix(tile.matrix)
    p = TileProcessor(stock_cfg)
    try:
        p.upload(tile.bcl, tile.cif, inv, firstclusterno=123456)
        p.process()
        o = oracle.process(stock_cfg, tile.bcl, tile.cif, inv, firstclusterno=123456)
        compare_tile(p, stock_cfg, None, o, firstclusterno=123456)
    finally:
        p.close()


def test_ruler_matches_oracle(proc, stock_cfg, oracle):
    tile, o = _run(proc, stock_cfg, oracle, 20000, 9)
    T = stock_cfg.total_cycles
    rng = np.random.default_rng(3)
    cut = np.clip(0.25 + 0.1 * rng.standard_normal(T), 0.01, 0.9)
    cut[rng.random(T) < 0.05] = np.nan          # "nan" cycles
    packed = proc.cutoffs_to_packed(cut)
    assert (packed == oracle.cutoffs_to_packed(cut)).all()
    proc.ruler_reset(T)
    want_hist = np.zeros((T, stock_cfg.dist_sampling_bins), dtype=np.uint32)
    nmeasured = 0
    for s in range(proc.params.num_samples):
        recs = o["records"][s]
        if recs.shape[0] == 0:
            continue
        want, want_hist = oracle.ruler(recs, packed, pos=want_hist)
        # (a) on host-provided records (the tailseq-polya-ruler binary's path)
        got = proc.ruler_records(recs, packed)
        assert (got == want).all()
        nmeasured += int((want >= 0).sum())
    assert nmeasured > 1000
    assert (proc.ruler_hist() == want_hist).all()
    # (b) on the resident record slots of the tile (fused pipeline), incl. withdrawn slots
    proc.ruler_reset(T)
    for s in range(proc.params.num_samples):
        slots = proc.records(s, drop_withdrawn=False)
        if slots.shape[0] == 0:
            continue
        got = proc.ruler_tile(s, packed)
        valid = (slots[:, 1] >> 16).astype(np.uint16).view(np.int16) >= 0
        want, _ = oracle.ruler(o["records"][s], packed)
        assert (got[valid] == want).all()
        assert (got[~valid] == -1).all()
    assert (proc.ruler_hist() == want_hist).all()
    # (c) every sample in one launch
    proc.ruler_reset(T)
    lens, offs = proc.ruler_tile_all(packed)
    for s in range(proc.params.num_samples):
        slots = proc.records(s, drop_withdrawn=False)
        got = lens[offs[s]:offs[s + 1]]
        assert got.size == slots.shape[0]
        if got.size:
            valid = (slots[:, 1] >> 16).astype(np.uint16).view(np.int16) >= 0
            want, _ = oracle.ruler(o["records"][s], packed)
            assert (got[valid] == want).all() and (got[~valid] == -1).all()
    assert (proc.ruler_hist() == want_hist).all()


@pytest.mark.parametrize("bins", [470, 64, 1000, 30000])
def test_ruler_histogram_bins(proc, stock_cfg, oracle, bins):
    """Positive resampling with other bin counts: 470 = 10 * 47 shares a factor with 2^23 - 1 and
    takes the literal fp64 binning, the others the integer form; 30000 bins do not fit the
    histogram kernel's private shared-memory rows and go straight to the global histogram."""
    tile, o = _run(proc, stock_cfg, oracle, 6000, 12)
    T = stock_cfg.total_cycles
    packed = proc.cutoffs_to_packed(np.full(T, 0.2))
    recs = o["records"][proc.params.num_samples - 1]
    want, want_hist = oracle.ruler(recs, packed, bins=bins, gap=0.25)
    proc.ruler_reset(T, bins)
    got = proc.ruler_records(recs, packed, bins=bins, gap=0.25)
    assert (got == want).all() and int((want >= 0).sum()) > 100
    assert (proc.ruler_hist() == want_hist).all()


def test_two_tiles_uploaded_ahead(stock_cfg, oracle):
    """Upload queue of depth two: tile B is copied in while tile A is processed; results of each
    process() are those of the oldest uploaded tile; a third upload is refused."""
    from tailseeker_b200 import _lib
    from tailseeker_b200.synth import make_tile
    from tailseeker_b200.importer import TileProcessor, invert_colormatrix
    p = TileProcessor(stock_cfg)
    try:
        tiles = [make_tile(stock_cfg, n, seed=40 + i) for i, n in enumerate((2500, 1300, 3100))]
        invs = [invert_colormatrix(t.matrix) for t in tiles]
        p.upload(tiles[0].bcl, tiles[0].cif, invs[0])
        p.upload(tiles[1].bcl, tiles[1].cif, invs[1])
        with pytest.raises(_lib.TskError):
            p.upload(tiles[2].bcl, tiles[2].cif, invs[2])
        for i in range(3):
            p.reset_counters()
            p.process()
            if i == 0:
                p.upload(tiles[2].bcl, tiles[2].cif, invs[2])
            o = oracle.process(stock_cfg, tiles[i].bcl, tiles[i].cif, invs[i])
            compare_tile(p, stock_cfg, None, o)
    finally:
        p.close()


# ---- PhiX control classification (controlaligner.c:107-145) ------------------------------

def _control_cfg(**kw):
    from tailseeker_b200.config import stock_config
    return stock_config("miseq", phix_match_name="PhiX", **kw)


@pytest.mark.parametrize("kw,n", [(dict(), 20000),
                                  (dict(phix_match_start=4, phix_match_length=36), 6000),
                                  (dict(phix_match_start=1, phix_match_length=48), 6000),
                                  # 64 > 20 / 0.35: alignments may cross the spacer, whole-target scan
                                  (dict(phix_match_start=2, phix_match_length=64), 1500)])
def test_control_classification(oracle, kw, n):
    from tailseeker_b200.importer import TileProcessor
    cfg = _control_cfg(**kw)
    p = TileProcessor(cfg)
    try:
        _, o = _run(p, cfg, oracle, n, 21, unknown_frac=0.2, phix_frac=0.6)
        calls = compare_tile(p, cfg, None, o)
        assert (calls["sample"] == 1).sum() > n // 50 and (calls["sample"] == 0).sum() > n // 50
    finally:
        p.close()


def test_control_golden_reads():
    """The reads scored by the reference aligner itself (tests/golden/control_reads.npz), laid
    out as a tile whose index reads are all no-calls, hence all barcode-unassigned."""
    import os
    from tailseeker_b200.importer import TileProcessor
    g = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "control_reads.npz"))
    cfg = _control_cfg(phix_match_start=int(g["first"]) + 1, phix_match_length=int(g["length"]))
    reads = g["reads"]
    n, span = reads.shape
    code = np.full(256, 255, np.uint8)
    code[np.frombuffer(b"ACGT", np.uint8)] = np.arange(4)
    bcl = np.zeros((cfg.total_cycles, n), np.uint8)
    span = min(span, cfg.index_start - 1)          # keep the index read all no-calls
    assert span >= int(g["first"]) + int(g["length"])
    c = code[reads[:, :span]].T
    bcl[:span] = np.where(c == 255, 0, (30 << 2) | c)
    cif = np.zeros((cfg.threep_length, 4, n), np.int16)
    p = TileProcessor(cfg)
    try:
        p.upload(bcl, cif, np.eye(4, dtype=np.float32))
        p.process()
        calls = p.calls()
        assert (calls["sample"] == g["decision"]).all()
        cnt = p.counters()
        assert cnt[1, 0] == int(g["decision"].sum()) and cnt[0, 0] == n - int(g["decision"].sum())
    finally:
        p.close()


@pytest.mark.parametrize("dump_len", [300, 231, 64, 7])
def test_ruler_synthetic_records(proc, oracle, dump_len):
    """Hand-made records: every valid length 0..dump_len (ragged ends, block boundaries at
    multiples of 32, records longer than 256 cycles), runs of +1 / -1 steps of random length,
    downhill tails, zero-score packets."""
    rng = np.random.default_rng(dump_len)
    n, ncut = 4000, 400
    cut = np.clip(0.3 + 0.05 * rng.standard_normal(ncut), 0.05, 0.9)
    cut[rng.random(ncut) < 0.03] = np.nan
    packed = proc.cutoffs_to_packed(cut)
    first = rng.integers(0, 60, n)
    valid = rng.integers(0, dump_len + 1, n)
    valid[:dump_len + 1] = np.arange(dump_len + 1)
    recs = np.zeros((n, 2 + dump_len), np.uint32)
    recs[:, 0] = np.arange(n)
    recs[:, 1] = (first & 0xffff) | (valid << 16)
    for r in range(n):
        tail = int(rng.integers(0, valid[r] + 1))
        hi = rng.random(valid[r]) < np.where(np.arange(valid[r]) < tail, 0.9, 0.15)
        score = np.where(hi, rng.uniform(0.4, 1.0, valid[r]), rng.uniform(0.0, 0.25, valid[r]))
        score[rng.random(valid[r]) < 0.05] = 0.0
        word = (np.round(score * 16777215).astype(np.uint32) << 1) | (rng.random(valid[r]) < 0.6)
        recs[r, 2:2 + valid[r]] = word
    want, want_hist = oracle.ruler(recs, packed)
    proc.ruler_reset(ncut)
    got = proc.ruler_records(recs, packed)
    assert (got == want).all() and (dump_len < 8 or int((want >= 0).sum()) > 500)
    assert (proc.ruler_hist() == want_hist).all()


def test_streamed_mode(stock_cfg, oracle):
    """tsk_tile_process_async / tsk_ruler_tile_all_async: two tiles queued back to back without
    reading anything in between; results of the second tile, the accumulated ruler histogram
    and the called counter equal the blocking calls'."""
    from tailseeker_b200.synth import make_tile
    from tailseeker_b200.importer import TileProcessor, invert_colormatrix
    cfg = stock_cfg
    T = cfg.total_cycles
    tiles = [make_tile(cfg, n, seed=sd) for n, sd in ((9000, 31), (7000, 32))]
    packed = None
    p = TileProcessor(cfg)
    try:
        packed = p.cutoffs_to_packed(np.full(T, 0.25))
        # blocking reference run
        want_called, want_hist = 0, None
        p.ruler_reset(T)
        for t in tiles:
            p.upload(t.bcl, t.cif, invert_colormatrix(t.matrix))
            p.process()
            lens, offs = p.ruler_tile_all(packed)
            want_called += int((lens >= 0).sum())
        want_hist = p.ruler_hist()
        want_lens, want_calls, want_counters = lens, p.calls(), p.counters()
        assert p.ruler_called() == want_called
        # streamed run
        p.reset_counters()
        p.ruler_reset(T)
        for t in tiles:
            p.upload(t.bcl, t.cif, invert_colormatrix(t.matrix))
            p.process_async()
            p.ruler_tile_all_async(packed)
        assert p.ruler_called() == want_called
        assert (p.ruler_hist() == want_hist).all()
        lens, offs2 = p.ruler_lengths()
        assert (offs2 == offs).all() and (lens == want_lens).all()
        assert (p.calls() == want_calls).all()
        assert (p.counters() == want_counters).all()
        o = oracle.process(cfg, tiles[1].bcl, tiles[1].cif, oracle.invert_matrix(tiles[1].matrix))
        for s in range(p.params.num_samples):
            assert (p.records(s) == o["records"][s]).all()
    finally:
        p.close()


def test_device_pointer_requirements(stock_cfg):
    """Zero-copy uploads: base-call rows must be 16-byte aligned (TMA); a misaligned pitch is refused
    at upload time with a message, nothing is launched."""
    import torch
    from tailseeker_b200 import _lib
    from tailseeker_b200.importer import TileProcessor
    T, L3 = stock_cfg.total_cycles, stock_cfg.threep_length
    n, pitch = 1000, 1004
    bcl = torch.zeros((T, pitch), dtype=torch.uint8, device="cuda")
    cif = torch.zeros((L3, 4, pitch), dtype=torch.int16, device="cuda")
    p = TileProcessor(stock_cfg)
    try:
        with pytest.raises(_lib.TskError, match="16-byte"):
            p.upload(bcl.data_ptr(), cif.data_ptr(), np.eye(4, dtype=np.float32), nclusters=n, on_device=True,
                     bcl_pitch=pitch, cif_pitch=pitch)
    finally:
        p.close()


def test_streamed_host_uploads_reuse_slots(stock_cfg, oracle):
    """Five host tiles through the two upload slots with nothing read in between: the copy of tile
    t+2 must wait for the kernels of tile t (same slot).  Pinned buffers and tiles whose copy is
    much shorter than their kernels make the overwrite visible if the wait is missing: the ruler
    histogram accumulated over all tiles, the counters and the last tile must equal the blocking run."""
    import torch
    from tailseeker_b200.synth import make_tile
    from tailseeker_b200.importer import TileProcessor, invert_colormatrix
    cfg = stock_cfg
    T = cfg.total_cycles
    tiles = [make_tile(cfg, n, seed=sd) for n, sd in ((6000, 41), (5000, 42), (6500, 43), (5500, 44), (6000, 45))]
    pinned = [(torch.from_numpy(t.bcl.copy()).pin_memory(), torch.from_numpy(t.cif.copy()).pin_memory()) for t in tiles]
    p = TileProcessor(cfg)
    try:
        packed = p.cutoffs_to_packed(np.full(T, 0.25))
        p.ruler_reset(T)
        for t in tiles:
            p.upload(t.bcl, t.cif, invert_colormatrix(t.matrix))
            p.process()
            p.ruler_tile_all(packed)
        want_hist, want_counters, want_called = p.ruler_hist(), p.counters(), p.ruler_called()
        for rep in range(3):
            p.reset_counters()
            p.ruler_reset(T)
            for t, (pb, pc) in zip(tiles, pinned):
                p.upload(pb.numpy(), pc.numpy(), invert_colormatrix(t.matrix))
                p.process_async()
                p.ruler_tile_all_async(packed)
            assert p.ruler_called() == want_called
            assert (p.ruler_hist() == want_hist).all()
            assert (p.counters() == want_counters).all()
        o = oracle.process(cfg, tiles[-1].bcl, tiles[-1].cif, oracle.invert_matrix(tiles[-1].matrix))
        for s in range(p.params.num_samples):
            assert (p.records(s) == o["records"][s]).all()
    finally:
        p.close()
