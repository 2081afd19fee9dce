"""Shared comparison helpers for the parity tests."""
import numpy as np


def compare_tile(proc, cfg, tile_or_arrays, oracle_out, firstclusterno=0):
    """Assert the CUDA results of `proc` (TileProcessor, already processed) equal the oracle's."""
    calls = proc.calls()
    ocalls = oracle_out["calls"]
    # withdrawn slots keep their slot number on the GPU; the oracle numbers only written records
    for field in ("sample", "flags", "delimiter_end", "polya_len", "terminal_mods", "written", "emitted"):
        a, b = calls[field], ocalls[field]
        bad = np.where(a != b)[0]
        assert bad.size == 0, "%s differs at clusters %s: gpu %s oracle %s" % (
            field, bad[:8], a[bad[:8]], b[bad[:8]])
    S = proc.params.num_samples
    for s in range(S):
        g = proc.records(s)
        o = oracle_out["records"][s]
        assert g.shape == o.shape, "sample %d: %s record slots vs oracle %s" % (s, g.shape, o.shape)
        if g.size:
            bad = np.where((g != o).any(axis=1))[0]
            assert bad.size == 0, "sample %d: %d records differ, first cluster %d" % (
                s, bad.size, int(o[bad[0], 0]))
        # slot numbers must index the (un-compacted) slot array in cluster order
        allslots = proc.records(s, drop_withdrawn=False)
        mine = np.where((calls["sample"] == s) & (calls["emitted"] == 1))[0]
        if mine.size:
            assert (allslots[calls["record_slot"][mine], 0] == mine + firstclusterno).all()
    pos, neg = proc.hist()
    assert (pos == oracle_out["pos"]).all(), "positive histogram differs (%d vs %d samples)" % (
        pos.sum(), oracle_out["pos"].sum())
    assert (neg == oracle_out["neg"]).all(), "negative histogram differs (%d vs %d samples)" % (
        neg.sum(), oracle_out["neg"].sum())
    assert (proc.counters() == oracle_out["counters"]).all(), "demultiplexing counters differ"
    return calls
