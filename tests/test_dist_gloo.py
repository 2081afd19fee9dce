"""CPU, world_size 2 over gloo: sharding tiles across ranks and all-reducing the integer
statistics gives exactly the single-process result (merge(shards) == single)."""
import os
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _worker(rank, world, port, seeds, outdir):
    sys.path.insert(0, ROOT)
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from oracle import pyoracle
    from tailseeker_b200.config import stock_config
    from tailseeker_b200.sharding import allreduce_stats, shard_tiles
    from tailseeker_b200.synth import make_tile
    cfg = stock_config("miseq")
    T, bins = cfg.total_cycles, cfg.dist_sampling_bins
    pos = torch.zeros((T, bins), dtype=torch.int64)
    neg = torch.zeros((T, bins), dtype=torch.int64)
    cnt = torch.zeros((cfg.to_cparams().num_samples, 6), dtype=torch.int64)
    for seed in shard_tiles(seeds, rank, world):
        tile = make_tile(cfg, 400, seed=seed)
        o = pyoracle.process(cfg, tile.bcl, tile.cif, pyoracle.invert_matrix(tile.matrix))
        pos += torch.from_numpy(o["pos"].astype(np.int64))
        neg += torch.from_numpy(o["neg"].astype(np.int64))
        cnt += torch.from_numpy(o["counters"].astype(np.int64))
    allreduce_stats([pos, neg, cnt])
    np.savez(os.path.join(outdir, "rank%d.npz" % rank), pos=pos.numpy(), neg=neg.numpy(), cnt=cnt.numpy())
    dist.destroy_process_group()


def test_sharded_statistics_equal_single_process(tmp_path, oracle, stock_cfg):
    from tailseeker_b200.synth import make_tile
    seeds = [301, 302, 303]
    port = 29500 + (os.getpid() % 2000)
    mp.spawn(_worker, args=(2, port, seeds, str(tmp_path)), nprocs=2, join=True)
    T, bins = stock_cfg.total_cycles, stock_cfg.dist_sampling_bins
    pos = np.zeros((T, bins), np.int64); neg = np.zeros((T, bins), np.int64)
    cnt = None
    for seed in seeds:
        tile = make_tile(stock_cfg, 400, seed=seed)
        o = oracle.process(stock_cfg, tile.bcl, tile.cif, oracle.invert_matrix(tile.matrix))
        pos += o["pos"]; neg += o["neg"]
        cnt = o["counters"].astype(np.int64) if cnt is None else cnt + o["counters"]
    for rank in range(2):
        r = np.load(os.path.join(str(tmp_path), "rank%d.npz" % rank))
        assert (r["pos"] == pos).all() and (r["neg"] == neg).all() and (r["cnt"] == cnt).all()


def test_shard_tiles_partition():
    from tailseeker_b200.sharding import shard_tiles
    tiles = list(range(14))
    for world in (1, 2, 4, 8):
        got = sorted(t for r in range(world) for t in shard_tiles(tiles, r, world))
        assert got == tiles
