#!/usr/bin/env python
"""Stage times on tiles with a large share of barcode-unassigned clusters (PhiX spike-in):
work items of the compact scoring kernel then span more clusters."""
import os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from tailseeker_b200.config import stock_config
from tailseeker_b200.importer import TileProcessor, invert_colormatrix
from tailseeker_b200.synth_torch import make_tile_device

cfg = stock_config("miseq", phix_match_name="PhiX")
N = 1070000
names = ["decode", "scan", "score", "fixups"]
for unk in (0.03, 0.15, 0.3, 0.4, 0.5, 0.7):
    bcl, cif, mtext, pitch = make_tile_device(cfg, N, seed=77, device="cuda", unknown_frac=unk, phix_frac=0.8)
    mtx = np.array([np.float32(x) for x in mtext.split()], dtype=np.float32)
    p = TileProcessor(cfg)
    p.upload(bcl.data_ptr(), cif.data_ptr(), invert_colormatrix(mtx), nclusters=N, on_device=True, bcl_pitch=pitch, cif_pitch=pitch)
    p.process()
    p.upload(bcl.data_ptr(), cif.data_ptr(), invert_colormatrix(mtx), nclusters=N, on_device=True, bcl_pitch=pitch, cif_pitch=pitch)
    ms = p.process_timed(3)
    kept = int((p.calls()["written"] == 1).sum())
    print("unassigned %.2f: written %.2f  " % (unk, kept / N) + "  ".join("%s %.3f ms" % (n, float(m)) for n, m in zip(names, ms[:4])))
    p.close()
    del bcl, cif
