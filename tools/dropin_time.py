#!/usr/bin/env python
"""Wall time of the drop-in tailseq-import against the reference binary on one synthetic tile
(GPU box; files on /dev/shm)."""
import os, sys, time
os.environ["TSK_HOST_TIMING"] = "1"
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import refrun
from tailseeker_b200.config import stock_config
from tailseeker_b200.synth import make_tile

n = int(sys.argv[1]) if len(sys.argv) > 1 else 200000
cfg = stock_config("miseq", phix_match_name="PhiX")
tile = make_tile(cfg, n, seed=5, phix_frac=0.6)
ours = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tailseeker_b200", "bin", "tailseq-import")
for name, binary, threads in (("reference x16", None, 16), ("drop-in", ours, 16), ("drop-in", ours, 1)):
    r = refrun.run_import(cfg, tile, threads=threads, binary=binary)
    print("%-14s threads=%-2d %7.2f s  %9.0f clusters/s" % (name, threads, r["wall_s"], n / r["wall_s"]))
    print("\n".join(l for l in r["stderr"].splitlines() if l.startswith("[timing]")))

# ---- the ruler programs on the largest sample of that tile
import numpy as np
work = "/dev/shm/tsk_dropin_ruler"
os.makedirs(work, exist_ok=True)
r = refrun.run_import(cfg, tile, workdir=work, threads=16, binary=ours)
tid = cfg.tile_id
cpath = os.path.join(work, "cutoffs.txt")
refrun.write_cutoffs_file(cpath, tid, [0.25] * cfg.total_cycles)
best = max(cfg.samples, key=lambda s: r["sigpack"][s.name][2].shape[0])
args = (tid, os.path.join(work, "signals", "%s_%s.sigpack" % (best.name, tid)), cpath,
        os.path.join(work, "taginfo", "%s_%s.txt.gz" % (best.name, tid)))
our_ruler = os.path.join(os.path.dirname(ours), "tailseq-polya-ruler")
for name, binary in (("reference ruler", None), ("drop-in ruler", our_ruler), ("drop-in ruler", our_ruler)):
    rr = refrun.run_ruler(*args, os.path.join(work, "rp.gz"), binary=binary)
    print("%-16s %6.2f s  (%d records of %s)" % (name, rr["wall_s"], r["sigpack"][best.name][2].shape[0], best.name))

# ---- --batch: all samples of the tile on one context
import subprocess
jobs = []
for s in cfg.samples:
    jobs.append("\t".join([tid, os.path.join(work, "signals", "%s_%s.sigpack" % (s.name, tid)), cpath, "8", "0.49",
                           os.path.join(work, "taginfo", "%s_%s.txt.gz" % (s.name, tid)), "1000", "0.1",
                           os.path.join(work, "bpos_%s.gz" % s.name), os.path.join(work, "btag_%s.txt" % s.name)]))
open(os.path.join(work, "jobs.tsv"), "w").write("\n".join(jobs) + "\n")
t0 = time.perf_counter()
subprocess.run([our_ruler, "--batch", os.path.join(work, "jobs.tsv")], check=True)
tb = time.perf_counter() - t0
t0 = time.perf_counter()
for s in cfg.samples:
    refrun.run_ruler(tid, os.path.join(work, "signals", "%s_%s.sigpack" % (s.name, tid)), cpath,
                     os.path.join(work, "taginfo", "%s_%s.txt.gz" % (s.name, tid)), os.path.join(work, "rp.gz"))
tr = time.perf_counter() - t0
print("all %d samples: reference ruler (one process each) %.2f s, drop-in --batch %.2f s" % (len(cfg.samples), tr, tb))
