"""CPU: the oracle against the reference binaries run LIVE (oracle/_ref, built from the unmodified
sources by oracle/Makefile; present in the build container and, prebuilt, on the GPU box)."""
import os
import tempfile

import numpy as np
import pytest

from oracle import refrun

pytestmark = pytest.mark.skipif(not refrun.available(), reason="oracle/_ref not built")


def test_importer_and_ruler_live(oracle, stock_cfg, tmp_path):
    from tailseeker_b200 import textio
    from tailseeker_b200.synth import make_tile
    cfg = stock_cfg
    tile = make_tile(cfg, 3000, seed=77, lowdiv_frac=0.05, dark_cluster_frac=0.01)
    work = str(tmp_path)
    r = refrun.run_import(cfg, tile, workdir=work)
    inv = oracle.invert_matrix(tile.matrix)
    o = oracle.process(cfg, tile.bcl, tile.cif, inv)
    seq, qual = textio.decode_basecalls(tile.bcl)
    for s in cfg.sample_list():
        if not s["has_streams"]:
            continue
        assert textio.format_taginfo(cfg, s, o["calls"], seq) == r["taginfo"][s["name"]]
        assert textio.format_seqqual(cfg, s, o["calls"], seq, qual) == r["seqqual"][s["name"]]
        assert (o["records"][s["numindex"]] == r["sigpack"][s["name"]][2]).all()
    assert (o["pos"] == r["pos"]).all() and (o["neg"] == r["neg"]).all()
    assert textio.format_stats_csv(cfg, o["counters"]) == r["stats_csv"]

    cut = np.full(cfg.total_cycles, 0.3)
    cut[100:110] = np.nan
    cpath = os.path.join(work, "cutoffs.txt")
    refrun.write_cutoffs_file(cpath, cfg.tile_id, [float(c) for c in cut])
    packed = oracle.cutoffs_to_packed(cut)
    s = cfg.sample_list()[-1]
    name, tid = s["name"], cfg.tile_id
    rr = refrun.run_ruler(tid, os.path.join(work, "signals", "%s_%s.sigpack" % (name, tid)), cpath,
                          os.path.join(work, "taginfo", "%s_%s.txt.gz" % (name, tid)),
                          os.path.join(work, "rp.gz"))
    recs = o["records"][s["numindex"]]
    lens, hist = oracle.ruler(recs, packed)
    assert (hist == rr["pos"]).all()
    measured = {int(c): int(l) for c, l in zip(recs[:, 0], lens) if l >= 0}
    assert textio.format_ruler_output(tid, r["taginfo"][name], measured) == rr["taginfo"]


def test_threads_do_not_change_reference_outputs(stock_cfg):
    """taginfo / sigdists identical for threads 1 vs 4 except fair-sampling
    order effects; checked here on a tile without over-subscribed buckets."""
    from tailseeker_b200.synth import make_tile
    tile = make_tile(stock_cfg, 1500, seed=78, lowdiv_frac=0.0, spikein_weight=0.0)
    a = refrun.run_import(stock_cfg, tile, threads=1)
    b = refrun.run_import(stock_cfg, tile, threads=4)
    assert a["taginfo"] == b["taginfo"] and a["stats_csv"] == b["stats_csv"]
    assert (a["pos"] == b["pos"]).all() and (a["neg"] == b["neg"]).all()
