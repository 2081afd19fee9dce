"""Wall time of the drop-in tailseq-dedup-perfect --sort against the reference pipeline
`sort -k6,6 -k1,1 -k2,2n | tailseq-dedup-perfect | sort -k1,1 -k2,2n` (tailseeker/main.py:272-278,
bgzip stages left out) on the same text, files on /dev/shm, LC_ALL=C, gzip level 1."""
import os
import subprocess
import sys
import tempfile
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tools"))
import dedup_bench  # noqa: E402
from oracle import dedup_oracle  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 4_000_000
hi, lo, flags, polya = dedup_bench.synth_keys(n, seed=3)
text = dedup_bench.keys_to_text(hi, lo, flags, polya)
threads = os.cpu_count() or 1
env = dict(os.environ, LC_ALL="C", TSK_SHIM_GZ_LEVEL="1", TSK_GZ_LEVEL="1", TSK_HOST_TIMING="1")
with tempfile.TemporaryDirectory(dir="/dev/shm") as d:
    src = os.path.join(d, "in.txt")
    open(src, "wb").write(text)
    ref = ("sort -t '\t' -k6,6 -k1,1 -k2,2n --parallel=%d -S 4G %s | %s %s/rt.gz | sort -t '\t' -k1,1 -k2,2n --parallel=%d -S 4G > %s/ro.txt"
           % (threads, src, dedup_oracle.REF_BIN, d, threads, d))
    our = "%s --sort %s/ot.gz < %s > %s/oo.txt" % (os.path.join(ROOT, "tailseeker_b200", "bin", "tailseq-dedup-perfect"), d, src, d)
    for name, cmd in (("reference pipeline", ref), ("drop-in --sort", our), ("drop-in --sort (again)", our)):
        t0 = time.perf_counter()
        subprocess.run(["bash", "-c", cmd], check=True, env=env)
        print("%-24s %.2f s  (%d records, %.2f M records/s)" % (name, time.perf_counter() - t0, n, n / (time.perf_counter() - t0) / 1e6))
    same_out = open(d + "/ro.txt", "rb").read() == open(d + "/oo.txt", "rb").read()
    import gzip
    same_trace = gzip.open(d + "/rt.gz").read() == gzip.open(d + "/ot.gz").read()
    print("stdout identical:", same_out, " trace identical:", same_trace)
