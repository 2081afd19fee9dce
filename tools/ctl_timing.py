#!/usr/bin/env python
"""Timing of the control classifier per read class (GPU box): tiles whose index reads are all
no-calls, so every cluster goes through k_control_classify; decode stage time with and
without a [control] section gives the classifier's cost."""
import sys, os
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from tailseeker_b200.config import stock_config
from tailseeker_b200.importer import TileProcessor
from tailseeker_b200.synth import phix_genome

N = int(sys.argv[1]) if len(sys.argv) > 1 else 32768
rng = np.random.default_rng(7)
fwd = phix_genome(); rev = (3 - fwd)[::-1]


def reads(kind):
    out = np.empty((N, 51), np.uint8)
    if kind == "random":
        return rng.integers(0, 4, (N, 51)).astype(np.uint8)
    for i in range(N):
        st = fwd if rng.random() < 0.5 else rev
        s0 = int(rng.integers(0, st.size - 60))
        r = st[s0:s0 + 51].copy()
        if kind == "noisy":
            ns = int(rng.integers(1, 9)); w = rng.integers(0, 46, ns)
            r[w] = (r[w] + rng.integers(1, 4, ns)) & 3
        out[i] = r
    return out


for kind in ("random", "verbatim", "noisy"):
    r = reads(kind)
    for control in ("", "PhiX"):
        cfg = stock_config("miseq", phix_match_name=control)
        bcl = np.zeros((cfg.total_cycles, N), np.uint8)
        bcl[:51] = ((30 << 2) | r).T
        cif = np.zeros((cfg.threep_length, 4, N), np.int16)
        p = TileProcessor(cfg)
        p.upload(bcl, cif, np.eye(4, dtype=np.float32))
        p.process()
        ms = p.process_timed(5)
        c = p.counters()
        print(kind, control or "-", "decode stage ms", float(ms[0]), "counters", c[0, 0], c[1, 0] if control else "")
        p.close()

if len(sys.argv) > 2:
    import torch
    from tailseeker_b200.synth_torch import make_tile_device
    from tailseeker_b200.importer import invert_colormatrix
    cfg = stock_config("miseq", phix_match_name="PhiX")
    Nb = int(sys.argv[2])
    bcl, cif, mtext, pitch = make_tile_device(cfg, Nb, seed=1000, device="cuda", phix_frac=0.6)
    mtx = np.array([np.float32(x) for x in mtext.split()], dtype=np.float32)
    p = TileProcessor(cfg)
    p.upload(bcl.data_ptr(), cif.data_ptr(), invert_colormatrix(mtx), nclusters=Nb, bcl_pitch=pitch,
             cif_pitch=pitch, on_device=True)
    p.process()
    calls = p.calls()
    un = np.where(calls["sample"] <= 1)[0]
    head = bcl[:51].cpu().numpy()[:, un]
    nn = (head[5:45] == 0).sum(axis=0)
    print("unassigned", un.size, "control", int((calls["sample"] == 1).sum()), "N-count histogram", np.bincount(nn))
    p.upload(bcl.data_ptr(), cif.data_ptr(), invert_colormatrix(mtx), nclusters=Nb, bcl_pitch=pitch,
             cif_pitch=pitch, on_device=True)
    print("decode stage ms", float(p.process_timed(3)[0]))
