#!/usr/bin/env python
"""Experiment: tiles alternate between two contexts on two streams (kernel tails of one tile
overlap the next tile's kernels).  Prints device-resident throughput for 1 and 2 contexts."""
import os, sys, time
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from tailseeker_b200.config import stock_config
from tailseeker_b200.importer import TileProcessor, invert_colormatrix
from tailseeker_b200.synth_torch import make_tile_device

cfg = stock_config("miseq", phix_match_name="PhiX")
N, ntiles = 1070000, 8
dev = torch.device("cuda", 0)
tiles = []
for t in range(ntiles):
    bcl, cif, mtext, pitch = make_tile_device(cfg, N, seed=1000 + t, device=dev, phix_frac=0.6)
    mtx = np.array([np.float32(x) for x in mtext.split()], dtype=np.float32)
    tiles.append((bcl, cif, invert_colormatrix(mtx), pitch))
T = cfg.total_cycles
for P in (1, 2, 3):
    procs = [TileProcessor(cfg) for _ in range(P)]
    streams = [torch.cuda.Stream() for _ in range(P)]
    for p, s in zip(procs, streams):
        p.set_stream(s.cuda_stream)
    packed = procs[0].cutoffs_to_packed(np.full(T, 0.25))
    def step():
        for p in procs:
            p.ruler_reset(T)
        for t, (bcl, cif, inv, pitch) in enumerate(tiles):
            p = procs[t % P]
            p.upload(bcl.data_ptr(), cif.data_ptr(), inv, nclusters=N, on_device=True, bcl_pitch=pitch, cif_pitch=pitch)
            p.process_async()
            p.ruler_tile_all_async(packed)
        return sum(p.ruler_called() for p in procs)
    for _ in range(2):
        step()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(3):
        called = step()
    torch.cuda.synchronize()
    dt = (time.perf_counter() - t0) / 3
    print("contexts", P, "ms/step %.2f" % (dt * 1e3), "M clusters/s %.1f" % (ntiles * N / dt / 1e6), "called", called)
    for p in procs:
        p.close()
