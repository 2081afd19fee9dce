"""Full-size checks (BASELINE.json configs[1]: one MiSeq v2 tile of 1.07 M clusters) through
size-independent properties: the three scoring kernels (compact work-list form, row form,
literal form -- independent code paths) must agree bit for bit on every output, a second pass
must reproduce the first, the counters must account for every cluster, and the ruler's three
entry points must agree.  Parity with the oracle and with the reference binaries at these sizes
is test_gpu_fullsize_oracle.py."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu

N = 1_070_000


def _checks(p):
    """Everything a tile produces, reduced to comparable numbers."""
    calls = p.calls()
    S = p.params.num_samples
    recs = []
    for s in range(S):
        r = p.records(s)
        # order-sensitive checksum of the record words
        w = np.arange(1, r.size + 1, dtype=np.uint64)
        recs.append((r.shape, int((r.astype(np.uint64).reshape(-1) * w).sum(dtype=np.uint64))))
    pos, neg = p.hist()
    return dict(calls=calls.tobytes(), recs=recs, pos=pos, neg=neg, counters=p.counters())


def _same(a, b):
    assert a["calls"] == b["calls"]
    assert a["recs"] == b["recs"]
    assert (a["pos"] == b["pos"]).all() and (a["neg"] == b["neg"]).all()
    assert (a["counters"] == b["counters"]).all()


@pytest.mark.parametrize("layout,N", [("miseq", N), ("hiseq", 1_600_000)])
def test_full_tile_three_kernels_agree(layout, N):
    from tailseeker_b200.config import stock_config
    from tailseeker_b200.importer import TileProcessor, invert_colormatrix
    from tailseeker_b200.synth_torch import make_tile_device
    cfg = stock_config(layout, phix_match_name="PhiX")
    bcl, cif, mtext, pitch = make_tile_device(cfg, N, seed=4242, device="cuda", phix_frac=0.6)
    inv = invert_colormatrix(np.array([np.float32(x) for x in mtext.split()], dtype=np.float32))
    p = TileProcessor(cfg)
    try:
        def run(mode):
            p.force_generic(mode)
            p.reset_counters()
            p.upload(bcl.data_ptr(), cif.data_ptr(), inv, nclusters=N, on_device=True, bcl_pitch=pitch, cif_pitch=pitch)
            p.process()
            return _checks(p)
        compact = run(0)
        _same(compact, run(0))                 # a second pass reproduces the first
        _same(compact, run(2))                 # row form of the fast kernel
        _same(compact, run(1))                 # literal kernel, every cluster
        p.force_generic(0)

        calls = np.frombuffer(compact["calls"], dtype=p.calls().dtype)
        c = compact["counters"].astype(np.int64)
        # every cluster is counted once under its sample's index-mismatch columns
        assert c[:, :3].sum() == N
        assert (np.bincount(calls["sample"], minlength=c.shape[0]) == c[:, :3].sum(axis=1)).all()
        assert c[1, 0] > 5000 and c[0, 0] > 5000            # PhiX and Unknown both present
        # one record slot per emitting cluster, in cluster order
        for s in range(p.params.num_samples):
            mine = np.where((calls["sample"] == s) & (calls["emitted"] == 1))[0]
            assert compact["recs"][s][0][0] == mine.size

        # ruler: one launch over every sample == one launch per sample == counter
        T = cfg.total_cycles
        packed = p.cutoffs_to_packed(np.full(T, 0.25))
        p.upload(bcl.data_ptr(), cif.data_ptr(), inv, nclusters=N, on_device=True, bcl_pitch=pitch, cif_pitch=pitch)
        p.process()
        p.ruler_reset(T)
        lens, offs = p.ruler_tile_all(packed)
        hist_all, called = p.ruler_hist(), p.ruler_called()
        assert called == int((lens >= 0).sum()) and called > N // 3
        p.ruler_reset(T)
        for s in range(p.params.num_samples):
            if offs[s + 1] > offs[s]:
                assert (p.ruler_tile(s, packed) == lens[offs[s]:offs[s + 1]]).all()
        assert (p.ruler_hist() == hist_all).all()
        # the positive samples of a called tail lie inside it: no more samples than called bases
        assert int(hist_all.sum()) <= int(lens[lens >= 0].astype(np.int64).sum())
    finally:
        p.close()
