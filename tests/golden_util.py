import os

import numpy as np

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def load(name):
    return np.load(os.path.join(GOLDEN, name + ".npz"))


def fixture_config(name):
    from tailseeker_b200.config import stock_config
    if name == "importer_miseq":
        return stock_config("miseq")
    if name == "importer_hiseq_keep":
        return stock_config("hiseq", keep_no_delimiter=1, keep_low_quality_balancer=1)
    raise KeyError(name)


def matrix_of(g):
    text = bytes(g["matrix_text"]).decode()
    return np.array([np.float32(t) for t in text.split()], dtype=np.float32)


def check_importer_outputs(cfg, g, calls, records_by_sample, pos, neg, counters):
    """Compare importer products (from the oracle or from the GPU) with the reference's files."""
    from tailseeker_b200 import textio
    seq, qual = textio.decode_basecalls(g["bcl"])
    for s in cfg.sample_list():
        if not s["has_streams"]:
            continue
        name = s["name"]
        assert textio.format_taginfo(cfg, s, calls, seq) == bytes(g["taginfo_" + name]), name
        assert textio.format_seqqual(cfg, s, calls, seq, qual) == bytes(g["seqqual_" + name]), name
        want = g["sigpack_" + name]
        got = records_by_sample[s["numindex"]]
        assert got.shape == want.shape and (got == want).all(), name
        total, dump = g["sigpack_meta_" + name]
        assert total == g["bcl"].shape[1] and dump == s["signal_dump_length"]
    assert (pos == g["pos"]).all() and (neg == g["neg"]).all()
    assert textio.format_stats_csv(cfg, counters) == bytes(g["stats_csv"]).decode()


def check_ruler_outputs(cfg, g, ruler_fn, taginfo_by_sample=None):
    """ruler_fn(records u32[n][w], packed cutoffs) -> (lengths, pos hist of this sample)."""
    from tailseeker_b200 import textio
    for s in cfg.sample_list():
        if not s["has_streams"]:
            continue
        name = s["name"]
        recs = g["sigpack_" + name]
        if recs.shape[0] == 0:
            assert g["ruler_pos_" + name].sum() == 0
            continue
        lens, hist = ruler_fn(recs, g["cutoffs"])
        assert (hist == g["ruler_pos_" + name]).all(), name
        measured = {int(c): int(l) for c, l in zip(recs[:, 0], lens) if l >= 0}
        text = textio.format_ruler_output(cfg.tile_id, bytes(g["taginfo_" + name]), measured)
        assert text == bytes(g["ruler_taginfo_" + name]), name
