"""Measurement of the duplicate-collapsing row (tsk_dedup_*): records/s for UMI sort + collapse of
one sample's tags, device-resident and end to end through the C ABI, with the reference pipeline
(`sort | tailseq-dedup-perfect | sort`, tailseeker/main.py:272-278) timed beside it on a bounded
sample.  Prints one JSON line in bench.py's format.  `python bench.py --workload dedup` runs this."""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import tempfile
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def synth_keys(n, seed=0, umi_len=30):
    """Packed 30-base UMIs: 55 % of the records unique, the rest in groups of 2..~40 and a few
    groups of thousands (abundant transcripts)."""
    rng = np.random.default_rng(seed)
    n_umis = int(n * 0.62)
    bases = rng.integers(1, 6, (n_umis, umi_len), dtype=np.uint64)
    bases[bases == 4] = 5                     # N is rare: most codes 4 become T
    bases[rng.random(bases.shape) < 0.001] = 4
    hi = np.zeros(n_umis, np.uint64)
    lo = np.zeros(n_umis, np.uint64)
    for i in range(min(21, umi_len)):
        hi |= bases[:, i] << np.uint64(60 - 3 * i)
    for i in range(21, umi_len):
        lo |= bases[:, i] << np.uint64(60 - 3 * (i - 21))
    pick = np.minimum(rng.integers(0, n_umis, n), rng.integers(0, n_umis, n))      # skewed multiplicities
    at = 0
    for size in (30000, 8000, 5000, 2500, 1000, 600):
        pick[at:at + size] = rng.integers(0, n_umis)
        at += size
    rng.shuffle(pick)
    flag_bits = np.array([0x1, 0x2, 0x4, 0x8, 0x10, 0x20, 0x40, 0x80, 0x100, 0x200, 0x400, 0x800], np.int32)
    probs = np.array([0.7, 0.1, 0.3, 0.4, 0.15, 0.05, 0.1, 0.02, 0.05, 0.03, 0.03, 0.02])
    flags = np.zeros(n, np.int32)
    for b, p in zip(flag_bits, probs):
        flags |= np.where(rng.random(n) < p, b, 0).astype(np.int32)
    polya = np.where(rng.random(n) < 0.1, -1, rng.integers(0, 231, n)).astype(np.int32)
    return hi[pick], lo[pick], flags, polya


def keys_to_text(hi, lo, flags, polya, umi_len=30):
    """taginfo lines for the reference pipeline, in (tile, cluster) order."""
    n = len(hi)
    sym = np.frombuffer(b"?ACGNT??", np.uint8)
    umi = np.empty((n, umi_len), np.uint8)
    for i in range(umi_len):
        w, r = (hi, i) if i < 21 else (lo, i - 21)
        umi[:, i] = sym[((w >> np.uint64(60 - 3 * r)) & np.uint64(7)).astype(np.int64)]
    tiles = [b"a%d" % (1101 + k) for k in range(14)]
    per = (n + 13) // 14
    out = []
    for i in range(n):
        out.append(b"%s\t%d\t%d\t%d\tT\t%s\n" % (tiles[i // per], i % per + 1, flags[i], polya[i], umi[i].tobytes()))
    return b"".join(out)


def reference_pipeline(text, threads):
    """`sort -k6,6 -k1,1 -k2,2n | tailseq-dedup-perfect trace | sort -k1,1 -k2,2n` (main.py:272-278;
    the bgzip stages are left out), C locale."""
    from oracle import dedup_oracle
    env = dict(os.environ, LC_ALL="C", TSK_SHIM_GZ_LEVEL="1")
    with tempfile.TemporaryDirectory() as d:
        src = os.path.join(d, "in.txt")
        with open(src, "wb") as f:
            f.write(text)
        cmd = ("sort -t '\t' -k6,6 -k1,1 -k2,2n --parallel=%d -S 2G %s | %s %s/trace.gz | "
               "sort -t '\t' -k1,1 -k2,2n --parallel=%d -S 2G > %s/out.txt") % (threads, src, dedup_oracle.REF_BIN, d, threads, d)
        t0 = time.perf_counter()
        subprocess.run(["bash", "-c", cmd], check=True, env=env)
        return time.perf_counter() - t0


def main(argv=None):
    ap = argparse.ArgumentParser()
    ap.add_argument("--records", type=int, default=16_000_000)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--ref-records", type=int, default=1_000_000)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args(argv)
    import torch
    from tailseeker_b200 import dedup
    from oracle import dedup_oracle
    n = args.records
    hi, lo, flags, polya = synth_keys(n)
    pinned = [torch.from_numpy(a).pin_memory() for a in (hi, lo, flags, polya)]
    d = dedup.Deduplicator(n)
    # device-resident: keys already in HBM (the sort starts from the uploaded arrays every time)
    d.upload(hi, lo, flags, polya)
    ms_sort, ms_col = [], []
    for it in range(args.warmup + args.steps):
        ng = d.run(sort_by_umi=True)
        if it >= args.warmup:
            a, b = d.elapsed_ms()
            ms_sort.append(a); ms_col.append(b)
    launches = d.launches() // (args.warmup + args.steps)
    ms = float(np.mean(ms_sort) + np.mean(ms_col))
    # end to end through the C ABI from pinned host arrays, results back to the host
    out = {k: torch.empty(n, dtype=torch.int32).pin_memory() for k in ("order", "group", "tval", "tslot")}
    gout = {k: torch.empty(n, dtype=torch.int32).pin_memory() for k in ("rep", "polya", "ndup", "kept")}
    L = d.L
    e2e = []
    for it in range(args.warmup + args.steps):
        t0 = time.perf_counter()
        d._check(L.tsk_dedup_upload(d.h, n, *[t.data_ptr() for t in pinned]))
        d.run(sort_by_umi=True)
        d._check(L.tsk_dedup_get_records(d.h, *[out[k].data_ptr() for k in ("order", "group", "tval", "tslot")]))
        d._check(L.tsk_dedup_get_groups(d.h, *[gout[k].data_ptr() for k in ("rep", "polya", "ndup", "kept")]))
        if it >= args.warmup:
            e2e.append(time.perf_counter() - t0)
    e2e_s = float(np.mean(e2e))
    peak = 6650.0
    try:
        peak = float(json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"])
    except Exception:
        pass
    alg = n * (16 + 4 + 4 + 16) + ng * 16          # keys, flags, length in; order, group, value, slot out; 4 words per group
    line = {
        "metric": "records/sec for UMI sort + duplicate collapse (tailseq-dedup-perfect)", "value": n / (ms * 1e-3),
        "unit": "records/s", "n_gpus": 1, "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "u64 keys / int32",
        "data": "synthetic", "config": {"workload": "one sample of a MiSeq run: %d tags, 30-base UMIs, %d groups" % (n, ng),
                                        "inputs_larger_than_L2": True},
        "stages_ms": {"umi_radix_sort": float(np.mean(ms_sort)), "group_scan_collapse": float(np.mean(ms_col))},
        "roofline": {"bound": "hbm", "achieved": alg / (ms * 1e-3) / 1e9, "peak": peak, "unit": "GB/s",
                     "frac": alg / (ms * 1e-3) / 1e9 / peak, "traffic": None,
                     "note": "algorithmic bytes = 40 B per record + 16 B per group; the 12-pass radix sort moves about 45 B per record and pass (histogram read + scatter)"},
        "e2e": {"value": n / e2e_s, "unit": "records/s", "h2d_bytes_per_step": 24 * n, "d2h_bytes_per_step": 16 * n + 16 * int(ng)},
        "gpu_launches": int(launches),
    }
    if not args.no_cpu_baseline and dedup_oracle.ref_available():
        m = min(args.ref_records, n)
        text = keys_to_text(hi[:m], lo[:m], flags[:m], polya[:m])
        threads = os.cpu_count() or 1
        secs = reference_pipeline(text, threads)
        line["cpu_baseline"] = {"value": m / secs, "unit": "records/s", "cores": threads, "kind": "reference",
                                "sample": "first %d records as text through GNU sort | oracle/_ref/tailseq-dedup-perfect | GNU sort (LC_ALL=C)" % m}
    print(json.dumps(line))


if __name__ == "__main__":
    main()
