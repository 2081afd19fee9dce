"""Run-wide sums on the devices (csrc/tsk_merge.cu).

One GPU: the accumulators equal the sum of the tiles' histograms / counters, and the batched fit on
device memory equals the per-tile host-histogram fit.
Two or more GPUs (skipped on a one-GPU box; `gpurun --gpus 2`): the CUDA path on shards of the tiles
followed by the NCCL all-reduce equals the single-GPU result on all tiles, through both
tsk_merge_allreduce (one packed u64 collective) and tsk_hist_allreduce (§8b, in place on the context);
and `tailseq-import --devices 0,1 --merged-sigdists ... --merged-stats ...` writes the sum of its
per-tile files."""
import csv
import os
import subprocess

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SEEDS = [(2500, 71), (1800, 72), (2200, 73), (2000, 74), (1500, 75)]


def _ngpus():
    import torch
    return torch.cuda.device_count()


def _tiles(cfg):
    from tailseeker_b200.synth import make_tile
    return [make_tile(cfg, n, seed=sd, unknown_frac=0.1) for n, sd in SEEDS]


def _run_tiles(proc, merge, tiles, packed, T):
    from tailseeker_b200.importer import invert_colormatrix
    per_tile = []
    proc.ruler_reset(T)
    for t in tiles:
        proc.upload(t.bcl, t.cif, invert_colormatrix(t.matrix))
        proc.process_async()
        merge.add_tile()
        proc.ruler_tile_all_async(packed)
        per_tile.append(proc.hist())
    merge.add_totals()
    return per_tile


def test_accumulators_and_batched_fit(stock_cfg):
    from tailseeker_b200.importer import TileProcessor
    from tailseeker_b200.merge import RunMerge
    cfg = stock_cfg
    T = cfg.total_cycles
    tiles = _tiles(cfg)
    p = TileProcessor(cfg)
    m = RunMerge(p, max_tiles=len(tiles))
    try:
        packed = p.cutoffs_to_packed(np.full(T, 0.25))
        per_tile = _run_tiles(p, m, tiles, packed, T)
        got = m.get()
        assert (got["pos"] == sum(h[0].astype(np.uint64) for h in per_tile)).all()
        assert (got["neg"] == sum(h[1].astype(np.uint64) for h in per_tile)).all()
        assert (got["ruler_pos"] == p.ruler_hist().astype(np.uint64)).all() and got["ruler_pos"].sum() > 0
        assert (got["counters"] == p.counters().astype(np.uint64)).all()
        assert m.tile_count() == len(tiles)
        # batched fit on device memory == the host-histogram entry point, tile by tile and run-wide
        cut = m.fit(positive=0, min_spots=30)
        assert cut.shape == (1 + len(tiles), T)
        want = p.fit_cutoffs(got["pos"], got["neg"], min_spots=30)
        assert np.array_equal(cut[0], want, equal_nan=True) and np.isfinite(cut[0]).sum() > 5
        for i, (pos, neg) in enumerate(per_tile):
            w = p.fit_cutoffs(pos.astype(np.uint64), neg.astype(np.uint64), min_spots=30)
            assert np.array_equal(cut[1 + i], w, equal_nan=True)
        # round >= 1 of the pipeline: the ruler's positives against the importer's negatives
        cut1 = m.fit(positive=1, min_spots=30, with_tiles=False)
        assert np.array_equal(cut1[0], p.fit_cutoffs(got["ruler_pos"], got["neg"], min_spots=30), equal_nan=True)
        m.reset(7)
        z = m.get()
        assert z["pos"].sum() == 0 and z["counters"].sum() == 0 and m.tile_count() == 0
    finally:
        m.close()
        p.close()


@pytest.mark.skipif("_ngpus() < 2")
def test_sharded_cuda_path_plus_nccl_equals_single_gpu(stock_cfg):
    from tailseeker_b200 import merge as M
    from tailseeker_b200.importer import TileProcessor
    from tailseeker_b200.sharding import shard_tiles
    cfg = stock_cfg
    T = cfg.total_cycles
    tiles = _tiles(cfg)
    ndev = min(_ngpus(), 4)
    # single GPU, all tiles
    p0 = TileProcessor(cfg, device=0)
    m0 = M.RunMerge(p0)
    packed = p0.cutoffs_to_packed(np.full(T, 0.25))
    _run_tiles(p0, m0, tiles, packed, T)
    want = m0.get()
    want_cut = m0.fit(positive=0, min_spots=30)[0]
    m0.close(); p0.close()
    # shards
    procs = [TileProcessor(cfg, device=d) for d in range(ndev)]
    merges = [M.RunMerge(p) for p in procs]
    comms = M.NcclComm.init_all(list(range(ndev)))
    try:
        for d in range(ndev):
            _run_tiles(procs[d], merges[d], shard_tiles(tiles, d, ndev), packed, T)
        M.group_start()
        for d in range(ndev):
            merges[d].allreduce(comms[d])
        M.group_end()
        for d in range(ndev):
            got = merges[d].get()
            for k in ("pos", "neg", "ruler_pos", "counters"):
                assert (got[k] == want[k]).all(), (d, k)
            assert np.array_equal(merges[d].fit(positive=0, min_spots=30)[0], want_cut, equal_nan=True)
        # §8b: in place on the contexts' own histograms (last tile of each shard) + counters + ruler histogram
        local = [(p.hist(), p.counters(), p.ruler_hist()) for p in procs]
        M.group_start()
        for d in range(ndev):
            M._lib.check(procs[d].lib.tsk_hist_allreduce(procs[d]._ctx, comms[d].handle))
        M.group_end()
        for d in range(ndev):
            pos, neg = procs[d].hist()
            assert (pos == sum(l[0][0] for l in local)).all() and (neg == sum(l[0][1] for l in local)).all()
            assert (procs[d].counters() == sum(l[1] for l in local)).all()
            assert (procs[d].ruler_hist() == sum(l[2] for l in local)).all()
            assert (procs[d].counters().astype(np.uint64) == want["counters"]).all()
    finally:
        for c in comms:
            c.destroy()
        for m in merges:
            m.close()
        for p in procs:
            p.close()


@pytest.mark.parametrize("devices", ["0", pytest.param("0,1", marks=pytest.mark.skipif("_ngpus() < 2"))])
def test_program_over_devices_writes_the_sums(tmp_path, devices):
    from oracle import refrun
    from tailseeker_b200.config import stock_config
    from tailseeker_b200.synth import make_tile, write_tile_files
    exe = os.path.join(ROOT, "tailseeker_b200", "bin", "tailseq-import")
    inis, tids = [], []
    for k, (n, sd) in enumerate(SEEDS[:4]):
        cfg = stock_config("miseq", phix_match_name="PhiX", tile=1101 + k)
        tile = make_tile(cfg, n, seed=sd, unknown_frac=0.1, phix_frac=0.5)
        w = tmp_path / ("t%d" % k)
        data = str(w / "data")
        mpath = write_tile_files(tile, cfg, data)
        for sub in ("seqqual", "taginfo", "signals", "sigdists-r00", "stats"):
            os.makedirs(str(tmp_path / sub), exist_ok=True)
        ini = str(w / "tile.ini")
        with open(ini, "w") as f:
            f.write(cfg.replace(data_dir=data, threep_colormatrix=mpath, threads=4).to_ini())
        inis.append(ini); tids.append(cfg.tile_id)
    r = subprocess.run([exe, "--devices", devices, "--merged-sigdists", str(tmp_path / "run_{posneg}.sigdists"),
                        "--merged-stats", str(tmp_path / "run_stats.csv")] + inis, cwd=str(tmp_path), capture_output=True)
    assert r.returncode == 0, r.stderr.decode()[-2000:]
    for kind in ("pos", "neg"):
        tot = sum(refrun.read_sigdists(str(tmp_path / "sigdists-r00" / ("%s_%s.sigdists" % (kind, t)))).astype(np.uint64) for t in tids)
        assert (refrun.read_sigdists(str(tmp_path / ("run_%s.sigdists" % kind))) == tot).all() and tot.sum() > 0
    rows = {}
    for t in tids:
        for row in csv.DictReader(open(str(tmp_path / "stats" / ("signal-proc-%s.csv" % t)))):
            acc = rows.setdefault(row["name"], [0] * 6)
            for i, k in enumerate(("cln_no_index_mm", "cln_1_index_mm", "cln_2+_index_mm", "cln_fp_mm", "cln_qc_fail", "cln_no_delim")):
                acc[i] += int(row[k])
    merged = list(csv.DictReader(open(str(tmp_path / "run_stats.csv"))))
    assert [m["name"] for m in merged] == sorted(rows) and all(m["tile"] == "(total)" for m in merged)
    for mrow in merged:
        a = rows[mrow["name"]]
        assert [int(mrow[k]) for k in ("cln_no_index_mm", "cln_1_index_mm", "cln_2+_index_mm", "cln_fp_mm", "cln_qc_fail", "cln_no_delim")] == a
        assert int(mrow["cln_passed"]) == a[0] + a[1] + a[2] - a[3] - a[4] - a[5]
