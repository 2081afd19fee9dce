"""TEST INFRASTRUCTURE ONLY -- drives the UNMODIFIED reference binaries built into
oracle/_ref/ (see oracle/Makefile) on a synthetic tile and parses their outputs.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs
may import this module.  Nothing here reads /root/reference at run time: the binaries are
prebuilt, so this also works on the GPU box.

Recipe = SURVEY.md appendix D: lay the tile out as a run-folder fragment, write the INI that
scripts/generate-signalproc-conf.py would emit, run `tailseq-import tile.ini` with
threads=1 (deterministic fair sampling / record order) and a read buffer large enough for one
block per tile (the per-block histogram rewrite of tailseq-import.c:358-370,559).
"""
from __future__ import annotations

import gzip
import os
import shutil
import struct
import subprocess
import tempfile
import time
from typing import Dict, List, Optional

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
REF_IMPORT = os.path.join(HERE, "_ref", "tailseq-import")
REF_RULER = os.path.join(HERE, "_ref", "tailseq-polya-ruler")


def available() -> bool:
    return os.access(REF_IMPORT, os.X_OK) and os.access(REF_RULER, os.X_OK)


def read_sigpack(path: str):
    """signal-packs.h:30-41 + tailseq-import.c:106-110.  Returns (total_clusters, dump_len,
    u32 array [nrecords][2 + dump_len])."""
    with gzip.open(path, "rb") as f:
        raw = f.read()
    elemsize, total, dump_len = struct.unpack("<III", raw[:12])
    assert elemsize == 4
    body = np.frombuffer(raw, dtype="<u4", offset=12)
    assert body.size % (2 + dump_len) == 0
    return total, dump_len, body.reshape(-1, 2 + dump_len).copy()


def read_sigdists(path: str) -> np.ndarray:
    with gzip.open(path, "rb") as f:
        raw = f.read()
    elemsize, cycles, bins = struct.unpack("<III", raw[:12])
    assert elemsize == 4
    return np.frombuffer(raw, dtype="<u4", offset=12).reshape(cycles, bins).copy()


def write_sigdists(path: str, counts: np.ndarray) -> None:
    with gzip.open(path, "wb", compresslevel=1) as f:
        f.write(struct.pack("<III", 4, counts.shape[0], counts.shape[1]))
        f.write(counts.astype("<u4").tobytes())


def run_import(cfg, tile, workdir: Optional[str] = None, threads: int = 1,
               keep: bool = False, gz_level: Optional[int] = 1, binary: Optional[str] = None,
               read_buffer_size: Optional[int] = None) -> Dict:
    """Run oracle/_ref/tailseq-import on `tile` with configuration `cfg`
    (tailseeker_b200.config.ImporterConfig).  Returns dict with per-sample taginfo / seqqual
    text, sigpack arrays, pos/neg histograms, the stats CSV text and the wall time."""
    from tailseeker_b200.synth import write_tile_files

    own = workdir is None
    if own:
        base = "/dev/shm" if os.path.isdir("/dev/shm") else None
        workdir = tempfile.mkdtemp(prefix="tskref_", dir=base)
    try:
        data = os.path.join(workdir, "data")
        mpath = write_tile_files(tile, cfg, data)
        for sub in ("seqqual", "taginfo", "signals", "sigdists-r00", "stats"):
            os.makedirs(os.path.join(workdir, sub), exist_ok=True)
        rcfg = cfg.replace(data_dir=data, threep_colormatrix=mpath, threads=threads,
                           read_buffer_size=(read_buffer_size if read_buffer_size is not None
                                             else max(cfg.read_buffer_size, 2147483648)))
        ini = os.path.join(workdir, "tile.ini")
        with open(ini, "w") as f:
            f.write(rcfg.to_ini())
        env = dict(os.environ)
        if gz_level is not None:
            env["TSK_SHIM_GZ_LEVEL"] = str(gz_level)
        t0 = time.perf_counter()
        proc = subprocess.run([binary or REF_IMPORT, ini], cwd=workdir, env=env,
                              stdout=subprocess.PIPE, stderr=subprocess.PIPE)
        wall = time.perf_counter() - t0
        if proc.returncode != 0:
            raise RuntimeError("tailseq-import failed (%d): %s"
                               % (proc.returncode, proc.stderr.decode()[-2000:]))
        out = dict(wall_s=wall, stdout=proc.stdout.decode(), stderr=proc.stderr.decode(),
                   taginfo={}, seqqual={}, sigpack={}, workdir=workdir, ini=ini)
        tid = rcfg.tile_id
        for s in cfg.samples:
            with gzip.open(os.path.join(workdir, "taginfo", "%s_%s.txt.gz" % (s.name, tid)),
                           "rb") as f:
                out["taginfo"][s.name] = f.read()
            with gzip.open(os.path.join(workdir, "seqqual", "%s_%s.txt.gz" % (s.name, tid)),
                           "rb") as f:
                out["seqqual"][s.name] = f.read()
            out["sigpack"][s.name] = read_sigpack(
                os.path.join(workdir, "signals", "%s_%s.sigpack" % (s.name, tid)))
        out["pos"] = read_sigdists(os.path.join(workdir, "sigdists-r00", "pos_%s.sigdists" % tid))
        out["neg"] = read_sigdists(os.path.join(workdir, "sigdists-r00", "neg_%s.sigdists" % tid))
        with open(os.path.join(workdir, "stats", "signal-proc-%s.csv" % tid)) as f:
            out["stats_csv"] = f.read()
        return out
    finally:
        if own and not keep:
            shutil.rmtree(workdir, ignore_errors=True)


def run_ruler(tile_id: str, sigpack_path: str, cutoffs_path: str, taginfo_path: str,
              sigdist_out: str, min_polya: int = 8, downhill_w: float = 0.49,
              bins: int = 1000, gap: float = 0.1, binary: Optional[str] = None) -> Dict:
    """Run oracle/_ref/tailseq-polya-ruler with the 9 positional arguments of
    polyaruler.c:402-418 (as measure-polya-lengths.py:55-60 does)."""
    t0 = time.perf_counter()
    proc = subprocess.run([binary or REF_RULER, tile_id, sigpack_path, cutoffs_path, str(min_polya),
                           repr(downhill_w), taginfo_path, str(bins), repr(gap), sigdist_out],
                          stdout=subprocess.PIPE, stderr=subprocess.PIPE)
    wall = time.perf_counter() - t0
    if proc.returncode != 0:
        raise RuntimeError("reference tailseq-polya-ruler failed (%d): %s"
                           % (proc.returncode, proc.stderr.decode()[-2000:]))
    return dict(wall_s=wall, taginfo=proc.stdout, pos=read_sigdists(sigdist_out))


def write_cutoffs_file(path: str, tile_id: str, cutoffs) -> None:
    """signal-cutoffs.txt line format, calculate-optimal-parameters.py:173-182."""
    with open(path, "w") as f:
        f.write(tile_id + "\t")
        for v in cutoffs:
            if v is None or (isinstance(v, float) and v != v):
                f.write("nan\t")
            else:
                f.write(format(v, ".6f") + "\t")
        f.write("\n")


def run_reference_estimator(sigdists, total_cycles, conf_seeder, threads=2, workdir=None):
    """Run the reference's scripts/calculate-optimal-parameters.py (needs /root/reference; used
    only to GENERATE golden vectors in the build container, never on the GPU box).
    sigdists: {(tile_id, 'pos'|'neg'): u32 array [cycles][bins]}.  Returns
    {tile_id: list of float|None per cycle} and the bases report text."""
    import json
    import sys
    ref = "/root/reference"
    if not os.path.isdir(ref):
        raise RuntimeError("reference tree not present")
    own = workdir is None
    if own:
        workdir = tempfile.mkdtemp(prefix="tskest_")
    try:
        inputs = []
        tiles = sorted({t for t, _ in sigdists})
        for (tile, kind), arr in sigdists.items():
            d = os.path.join(workdir, "sigdists")
            os.makedirs(d, exist_ok=True)
            path = os.path.join(d, "%s_%s.sigdists" % (kind, tile))
            write_sigdists(path, arr)
            inputs.append([None, path])
        out_values = os.path.join(workdir, "signal-cutoffs.txt")
        out_bases = os.path.join(workdir, "signal-cutoff-bases.txt")
        params = {
            "input": inputs,
            "output": [["cutoff_values", out_values], ["cutoff_bases", out_bases]],
            "threads": threads,
            "wildcards": [],
            "params": [["conf", {"polyA_seeder": conf_seeder}],
                       ["tileinfo", {t: {"id": t, "source": "run"} for t in tiles}],
                       ["total_cycles", total_cycles]],
        }
        pj = os.path.join(workdir, "params.json")
        with open(pj, "w") as f:
            json.dump(params, f)
        env = dict(os.environ)
        env["SNAKEMAKE_PARAMS"] = pj
        env["PYTHONPATH"] = os.path.join(HERE, "snakemake_stub") + ":" + ref
        proc = subprocess.run([sys.executable, os.path.join(ref, "scripts", "calculate-optimal-parameters.py")],
                              env=env, cwd=workdir, stdout=subprocess.PIPE, stderr=subprocess.PIPE)
        if proc.returncode != 0:
            raise RuntimeError("reference estimator failed: " + proc.stderr.decode()[-3000:])
        text = open(out_values).read()
        bases = open(out_bases).read()
        cut = {}
        for line in text.splitlines():
            toks = line.split("\t")
            cut[toks[0]] = [None if t == "nan" else float(t) for t in toks[1:] if t != ""]
        return dict(cutoffs=cut, text=text, bases=bases)
    finally:
        if own:
            shutil.rmtree(workdir, ignore_errors=True)
