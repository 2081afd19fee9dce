"""SURVEY.md §8(d) "CPU baseline beside it": the reference binaries (oracle/_ref, unmodified sources)
on this box's host cores, MiSeq tile sample: tailseq-import with threads=1 and threads=nproc,
tailseq-polya-ruler over all samples, the drop-in programs next to them, and the share of the
importer's time spent in output compression (same run at zlib level 6 = htslib's default, level 1
and level 0).  Writes a small text table; run on the GPU box:  python tools/cpu_baseline_report.py"""
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import refrun  # noqa: E402
from tailseeker_b200.config import stock_config  # noqa: E402
from tailseeker_b200.synth import make_tile  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 100_000
cfg = stock_config("miseq", phix_match_name="PhiX")
tile = make_tile(cfg, n, seed=77, unknown_frac=0.03, phix_frac=0.6, lowdiv_frac=0.02, dark_cluster_frac=0.003,
                 spikein_weight=0.25)
nproc = os.cpu_count() or 1
ours = os.path.join(ROOT, "tailseeker_b200", "bin", "tailseq-import")
rows = []


def imp(label, threads, level, binary=None):
    r = refrun.run_import(cfg, tile, threads=threads, gz_level=level, binary=binary)
    rows.append((label, threads, level, r["wall_s"], n / r["wall_s"]))
    return r["wall_s"]


t1_6 = imp("reference tailseq-import", 1, 6)
t1_1 = imp("reference tailseq-import", 1, 1)
t1_0 = imp("reference tailseq-import", 1, 0)
tn_6 = imp("reference tailseq-import", nproc, 6)
tn_1 = imp("reference tailseq-import", nproc, 1)
if os.access(ours, os.X_OK):
    os.environ["TSK_GZ_LEVEL"] = "1"
    try:
        imp("B200 tailseq-import (1 process, incl. CUDA start-up)", nproc, 1, binary=ours)
    except RuntimeError as e:                       # no GPU here
        print("drop-in importer not timed:", str(e).splitlines()[0])

# ruler: every sample of the tile, one process each (what the pipeline does)
import tempfile  # noqa: E402
work = tempfile.mkdtemp(prefix="tskcpu_", dir="/dev/shm" if os.path.isdir("/dev/shm") else None)
r = refrun.run_import(cfg, tile, workdir=work, threads=nproc, gz_level=1)
tid = cfg.tile_id
cpath = os.path.join(work, "cutoffs.txt")
refrun.write_cutoffs_file(cpath, tid, [0.25] * cfg.total_cycles)
secs, nrec = 0.0, 0
for s in cfg.samples:
    rr = refrun.run_ruler(tid, os.path.join(work, "signals", "%s_%s.sigpack" % (s.name, tid)), cpath,
                          os.path.join(work, "taginfo", "%s_%s.txt.gz" % (s.name, tid)), os.path.join(work, "sd_%s.gz" % s.name))
    secs += rr["wall_s"]
    nrec += int((r["sigpack"][s.name][2][:, 0] >= 0).sum()) if len(r["sigpack"][s.name][2]) else 0
import shutil  # noqa: E402
shutil.rmtree(work, ignore_errors=True)

out = ["# reference binaries on this box: %d host cores, MiSeq tile sample of %d clusters, files on /dev/shm" % (nproc, n),
       "%-58s %7s %5s %9s %14s" % ("program", "threads", "zlib", "wall s", "clusters/s")]
for label, th, lv, w, cps in rows:
    out.append("%-58s %7d %5d %9.2f %14.0f" % (label, th, lv, w, cps))
out.append("%-58s %7d %5d %9.2f %14.0f   (%d processes, %d records)" % ("reference tailseq-polya-ruler, all samples", 1, 1, secs, n / secs, len(cfg.samples), nrec))
out.append("compression share of tailseq-import, threads=1: %.0f %% at zlib level 6 (htslib's default), %.0f %% at level 1"
           % (100 * (t1_6 - t1_0) / t1_6, 100 * (t1_1 - t1_0) / t1_1))
out.append("threads=%d over threads=1 at level 6: %.2fx" % (nproc, t1_6 / tn_6))
text = "\n".join(out) + "\n"
print(text)
dst = os.path.join(ROOT, "gpurun_out")
os.makedirs(dst, exist_ok=True)
open(os.path.join(dst, "r1_cpu_baseline.txt"), "w").write(text)
