#!/usr/bin/env python
"""bench.py -- clusters/sec of the TAIL-seq signal-processing hot path on B200.

Metric (BASELINE.json): clusters/sec for decode + normalise + poly(A) call; HBM GB/s vs peak.
Workload at every N: BASELINE.json configs[1] -- a MiSeq v2 full run, 14 tiles of ~1.07 M
clusters (15 M clusters, 51+251 cycles, 4 experimental + 7 spike-in samples) PER GPU (weak
scaling: tiles and lanes are independent, one tile set per GPU, no data-path collective;
the only exchange is the all-reduce of the QC histograms / counters at the end of a step).

A step = one pass of the hot path over the whole tile set:
  per tile : decode_demux_seed -> slot scan -> normalise_score_pack (+ fair-sampling fix-ups)
             -> polya_ruler over every sample's records with the previous round's cutoffs
  per step : all-reduce (N>1) of the summed pos/neg histograms and demux counters,
             fit_cutoffs on the run-wide histograms
`value`  : inputs resident in HBM when the timed region starts.
`e2e`    : the same step through the C ABI with HOST (pinned) buffers: H2D of every tile and
           D2H of calls / poly(A) lengths / histograms inside the timed region.
--impl reference : the reference's own CPU implementation (oracle/_ref binaries built from
           the unmodified sources) on the box's host cores, on a bounded sample of the same
           workload.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import tempfile
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

TILES_PER_GPU = 14
CLUSTERS_PER_TILE = 1_070_000
METRIC = "clusters/sec for decode+normalise+poly(A) call"
UNIT = "clusters/s"
PHIX_FRAC = 0.6     # of the 3 % barcode-unassigned clusters: reads of the PhiX control (half with errors)
WORKLOAD = ("MiSeq v2 full run: %d tiles x %d clusters, 51+251 cycles, 4 exp + 7 spike-in samples, "
            "[control] PhiX as generate-signalproc-conf.py always emits it "
            "(BASELINE.json configs[1]) per GPU" % (TILES_PER_GPU, CLUSTERS_PER_TILE))


def measured_peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        with open(path) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
    return 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons DURING the timed region (B200_PROFILING.md)."""
    Q = ("timestamp,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
         "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index: int):
        self.gpu = gpu_index
        self.proc = None
        self.path = None

    def start(self):
        fd, self.path = tempfile.mkstemp(prefix="clocks_", suffix=".csv")
        os.close(fd)
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", "-i", str(self.gpu), "--query-gpu=" + self.Q,
                 "--format=csv,noheader,nounits", "-lms", "50"],
                stdout=open(self.path, "w"), stderr=subprocess.DEVNULL)
        except OSError:
            self.proc = None

    @staticmethod
    def _when(text):
        import datetime
        try:
            return datetime.datetime.strptime(text.strip(), "%Y/%m/%d %H:%M:%S.%f").timestamp()
        except ValueError:
            return None

    def stop(self, t0=None, t1=None):
        """Median SM clock and throttle reasons of the samples taken in [t0, t1] (host clock), the
        timed region; the sampler is started before the warm-up steps because nvidia-smi needs a
        few hundred ms to deliver its first sample and the timed region of the device-resident run
        is shorter than that.  Without a sample inside the window the ones taken during the
        warm-up steps (same kernels, same load) are reported and `window` says so."""
        out = {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0, "window": "timed"}
        if self.proc is not None:
            self.proc.terminate()
            try:
                self.proc.wait(timeout=5)
            except subprocess.TimeoutExpired:
                self.proc.kill()
        try:
            rows = [r.strip().split(", ") for r in open(self.path) if r.strip()]
            os.unlink(self.path)
        except OSError:
            rows = []
        if t0 is not None and t1 is not None:
            inside = [r for r in rows if len(r) >= 9 and (self._when(r[0]) or 0) >= t0 and (self._when(r[0]) or 0) <= t1]
            if inside:
                rows = inside
            else:
                out["window"] = "warmup+timed"
                rows = [r for r in rows if len(r) >= 9 and (self._when(r[0]) or t1 + 1) <= t1]
        sm, reasons = [], set()
        for r in rows:
            if len(r) < 9:
                continue
            try:
                sm.append(float(r[1]))
                out["sm_max_mhz"] = float(r[2])
            except ValueError:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown",
                                "sw_power_cap"), r[5:9]):
                if v.strip().lower() == "active":
                    reasons.add(name)
        if sm:
            out["sm_mhz"] = float(np.median(sm))
            out["samples"] = len(sm)
        out["reasons"] = sorted(reasons)
        return out


# ------------------------------------------------------------------------------------------
# reference arm / cpu baseline: the unmodified reference binaries on host cores
# ------------------------------------------------------------------------------------------
def reference_sample_run(cfg, bcl, cif, matrix_text, threads, cutoffs):
    """One pass of the reference's CPU path (tailseq-import, then tailseq-polya-ruler per
    sample) over a host tile.  Returns (clusters, seconds of the binaries' wall time)."""
    from oracle import refrun
    from tailseeker_b200.synth import Tile

    tile = Tile(bcl=bcl, cif=cif, matrix_text=matrix_text, nclusters=bcl.shape[1],
                truth_sample=np.zeros(0, np.int32))
    base = "/dev/shm" if os.path.isdir("/dev/shm") else None
    work = tempfile.mkdtemp(prefix="tskbench_", dir=base)
    try:
        r = refrun.run_import(cfg, tile, workdir=work, threads=threads, gz_level=1, read_buffer_size=4 << 30)
        secs = r["wall_s"]
        tid = cfg.tile_id
        cpath = os.path.join(work, "cutoffs.txt")
        refrun.write_cutoffs_file(cpath, tid, list(cutoffs))
        for s in cfg.samples:
            rr = refrun.run_ruler(tid, os.path.join(work, "signals", "%s_%s.sigpack" % (s.name, tid)),
                                  cpath, os.path.join(work, "taginfo", "%s_%s.txt.gz" % (s.name, tid)),
                                  os.path.join(work, "sigdist_%s.gz" % s.name))
            secs += rr["wall_s"]
        return tile.nclusters, secs
    finally:
        import shutil
        shutil.rmtree(work, ignore_errors=True)


def default_cutoffs(T):
    return np.full(T, 0.25, dtype=np.float64)


def workload_config(layout, ntiles, N, T):
    """The `config` both arms print (identical keys and values: the driver compares them)."""
    workload = WORKLOAD if (layout, ntiles, N) == ("miseq", TILES_PER_GPU, CLUSTERS_PER_TILE) else (
        "%s layout: %d tiles x %d clusters per GPU, %d cycles, 4 exp + 7 spike-in samples, [control] PhiX"
        % (layout, ntiles, N, T))
    return {"workload": workload, "tiles_per_gpu": ntiles, "clusters_per_tile": N,
            "l2": "inputs_larger_than_L2 (2.5 GB per tile, streamed once)",
            "ruler_rounds": 1}


def run_reference_arm(args):
    """The reference's own CPU implementation of the path (oracle/_ref: the unmodified sources) with
    every host thread it can use, one FULL tile of the workload per step."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    from oracle import refrun
    from tailseeker_b200.config import stock_config
    cfg = stock_config(args.layout, phix_match_name="PhiX")
    line = {"impl": "reference", "metric": METRIC, "unit": UNIT, "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": workload_config(args.layout, args.tiles, args.clusters_per_tile, cfg.total_cycles)}
    if not refrun.available():
        line["unavailable"] = "oracle/_ref binaries are not built (run __graft_entry__.build() where /root/reference exists)"
        print(json.dumps(line))
        return 0
    cores = os.cpu_count() or 1
    nsample = args.ref_clusters if args.ref_clusters > 0 else args.clusters_per_tile
    bcl, cif, mtext = host_tile(cfg, nsample, seed=1000)
    cut = default_cutoffs(cfg.total_cycles)
    times = []
    for it in range(args.warmup + args.steps):
        n, secs = reference_sample_run(cfg, bcl, cif, mtext, cores, cut)
        if it >= args.warmup:
            times.append(secs)
    value = nsample / (sum(times) / len(times))
    sample = ("one %d-cluster tile of the workload per step (the per-GPU tile set is %d such tiles): tailseq-import "
              "(threads=%d, one read block, zlib level 1 instead of BGZF) + tailseq-polya-ruler per sample, files on "
              "/dev/shm; wall time of the binaries (file read + compute + compress + write)" % (nsample, args.tiles, cores))
    line.update({"value": value, "ms_per_step": 1e3 * sum(times) / len(times),
                 "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": "reference",
                                  "sample": sample},
                 "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
                 "gpu_launches": 0})
    print(json.dumps(line))
    return 0


def host_tile(cfg, n, seed):
    """A host tile of the workload's recipe: generated on the GPU when there is one (seconds for a
    full tile), by the NumPy generator otherwise."""
    try:
        import torch
        if torch.cuda.is_available():
            from tailseeker_b200.synth_torch import make_tile_device
            bcl, cif, mtext, pitch = make_tile_device(cfg, n, seed=seed, device="cuda", phix_frac=PHIX_FRAC)
            out = (np.ascontiguousarray(bcl[:, :n].cpu().numpy()), np.ascontiguousarray(cif[:, :, :n].cpu().numpy()), mtext)
            del bcl, cif
            torch.cuda.empty_cache()
            return out
    except ImportError:
        pass
    from tailseeker_b200.synth import make_tile
    t = make_tile(cfg, n, seed=seed, phix_frac=PHIX_FRAC)
    return t.bcl, t.cif, t.matrix_text


# ------------------------------------------------------------------------------------------
# the command-line surface: files -> bin/tailseq-import -> files, beside the reference binary
# ------------------------------------------------------------------------------------------
def cli_leg(cfg, ntiles, N, threads):
    """`tailseq-import a.ini b.ini ...` (drop-in: C host + CUDA, text and BGZF blocks made on the GPU)
    against oracle/_ref/tailseq-import (one process per tile, zlib level 1), same files on /dev/shm,
    same [options] threads.  Returns clusters/s of both and the drop-in's phase times."""
    import shutil
    from oracle import refrun
    from tailseeker_b200.synth import Tile, write_tile_files
    exe = os.path.join(ROOT, "tailseeker_b200", "bin", "tailseq-import")
    if not (os.access(exe, os.X_OK) and refrun.available()):
        return None
    base = tempfile.mkdtemp(prefix="tskcli_", dir="/dev/shm" if os.path.isdir("/dev/shm") else None)
    try:
        inis = {"ref": [], "our": []}
        for k in range(ntiles):
            tcfg = cfg.replace(tile=1101 + k)
            bcl, cif, mtext = host_tile(tcfg, N, seed=7000 + k)
            data = os.path.join(base, "data%d" % k)
            mpath = write_tile_files(Tile(bcl=bcl, cif=cif, matrix_text=mtext, nclusters=N, truth_sample=None), tcfg, data)
            del bcl, cif
            for arm in ("ref", "our"):
                w = os.path.join(base, arm)
                for sub in ("seqqual", "taginfo", "signals", "sigdists-r00", "stats"):
                    os.makedirs(os.path.join(w, sub), exist_ok=True)
                ini = os.path.join(w, "tile%d.ini" % k)
                with open(ini, "w") as f:
                    f.write(tcfg.replace(data_dir=data, threep_colormatrix=mpath, threads=threads,
                                         read_buffer_size=8 << 30).to_ini())
                inis[arm].append(ini)
        env = dict(os.environ, TSK_SHIM_GZ_LEVEL="1", TSK_HOST_TIMING="1")
        t0 = time.perf_counter()
        for ini in inis["ref"]:
            subprocess.run([refrun.REF_IMPORT, ini], cwd=os.path.join(base, "ref"), env=env, check=True,
                           stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL)
        tref = time.perf_counter() - t0
        t0 = time.perf_counter()
        r = subprocess.run([exe] + inis["our"], cwd=os.path.join(base, "our"), env=env, check=True,
                           stdout=subprocess.DEVNULL, stderr=subprocess.PIPE)
        tour = time.perf_counter() - t0
        phases = {}
        if os.environ.get("TSK_CLI_RAW_TIMING"):                 # per-tile lines, for tools/cli_time.py
            sys.stderr.write(r.stderr.decode())
        per_tile = []                       # a tile's phases start with "open files + buffers"
        for ln in r.stderr.decode().splitlines():
            if ln.startswith("[timing]"):
                name, secs = ln[8:].rsplit(None, 2)[0].strip(), float(ln.split()[-2])
                phases[name] = round(phases.get(name, 0.0) + secs, 3)
                if name == "open files + buffers":
                    per_tile.append(0.0)
                if per_tile:
                    per_tile[-1] += secs
        later = per_tile[1:] if len(per_tile) > 1 else []
        sizes = {arm: sum(os.path.getsize(os.path.join(dp, f)) for sub in ("seqqual", "taginfo", "signals")
                          for dp, _, fs in os.walk(os.path.join(base, arm, sub)) for f in fs) for arm in ("ref", "our")}
        start = phases.get("wait for the CUDA context", 0.0)
        return {"value": ntiles * N / tour, "unit": UNIT, "reference_value": ntiles * N / tref,
                "value_without_cuda_start": ntiles * N / max(tour - start, 1e-9),
                "speedup": tref / tour, "tiles": ntiles, "clusters_per_tile": N, "threads": threads,
                "seconds": round(tour, 3), "reference_seconds": round(tref, 3),
                "output_bytes": sizes["our"], "reference_output_bytes": sizes["ref"], "phases_s": phases,
                "seconds_per_tile": [round(x, 3) for x in per_tile],
                "steady_state_value": (N * len(later) / sum(later)) if later and sum(later) > 0 else None,
                "note": "files on /dev/shm -> tailseq-import (one process, all tiles; CUDA start-up included: creating the "
                        "context takes 0.2 - 3 s on these boxes, the phase 'wait for the CUDA context' says how much; the first tile also "
                        "pays the first-use allocations, steady_state_value = clusters/s over the later tiles) -> files; "
                        "reference: its binary once per tile, zlib level 1; outputs of both decompress to the same bytes "
                        "(tests/test_gpu_dropin.py, test_gpu_fullsize_oracle.py)"}
    finally:
        shutil.rmtree(base, ignore_errors=True)


# ------------------------------------------------------------------------------------------
# device arm
# ------------------------------------------------------------------------------------------
class DevView:
    """Expose raw device memory of the library to torch (CUDA array interface)."""
    def __init__(self, ptr, shape, typestr):
        self.__cuda_array_interface__ = {"shape": shape, "typestr": typestr, "data": (ptr, False),
                                         "version": 2}


def algorithmic_bytes(cfg, params, calls, nclusters, slot_counts):
    """bytes that must move per tile (stated in DESIGN.md).
    K1 decode_demux_seed   : N*T (BCL) + N*16 (call written)
    K2 normalise_score_pack: per processed cluster 8*(balancer_len + insert_len) CIF bytes
                             + 16 (call read) + per record slot (8 + 4*dump_len) written
    ruler                  : per record slot (8 + 4*dump_len) read + 2 (int16 length) written"""
    T, L3, ts = cfg.total_cycles, cfg.threep_length, cfg.threep_start - 1
    de = calls["delimiter_end"].astype(np.int64)
    proc = (calls["written"] == 1) & (de >= 0) & ((calls["flags"] & 0x200) == 0)
    limit = np.array([params.samples[s].limit_threep_processing for s in range(params.num_samples)])
    dump = np.array([params.samples[s].signal_dump_length for s in range(params.num_samples)])
    smp = calls["sample"].astype(np.int64)
    ins = ts + L3 - de
    lim = limit[smp]
    ins = np.where((lim > 0) & (lim < ins), lim, ins)
    blen = np.minimum(de - ts, cfg.balancer_length)
    k2 = int((8 * (blen + ins) + 16)[proc].sum())
    nslots = int(sum(slot_counts))
    recbytes = int(sum(c * (8 + 4 * int(d)) for c, d in zip(slot_counts, dump)))
    k1 = nclusters * (T + 16)
    return {"k1": k1, "k2": k2 + recbytes, "ruler": recbytes + 2 * nslots,
            "total": k1 + k2 + 2 * recbytes + 2 * nslots, "records": nslots,
            "processed": int(proc.sum())}


def pin_rank_to_cores(local_rank, local_world):
    """Each rank keeps to its own slice of the host cores: eight ranks that all wander over the same
    cores (and the same memory controller queues) is what the end-to-end curve of round 1 showed."""
    try:
        cores = sorted(os.sched_getaffinity(0))
        per = max(1, len(cores) // max(1, local_world))
        mine = cores[local_rank * per:(local_rank + 1) * per] or cores
        os.sched_setaffinity(0, mine)
        return len(mine)
    except (AttributeError, OSError):
        return None


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--tiles", type=int, default=TILES_PER_GPU)
    ap.add_argument("--clusters-per-tile", type=int, default=CLUSTERS_PER_TILE)
    ap.add_argument("--ref-clusters", type=int, default=0,
                    help="clusters in the CPU sample (reference arm / cpu_baseline); 0 = one full tile")
    ap.add_argument("--workload", default="import", choices=["import", "dedup"],
                    help="import (default): the north-star path; dedup: the duplicate-collapsing row (tools/dedup_bench.py)")
    ap.add_argument("--layout", default="miseq", choices=["miseq", "hiseq"],
                    help="read layout of the synthetic run; miseq (default) is the configuration the metric is quoted on, "
                         "hiseq = BASELINE.json configs[2] (use --clusters-per-tile 3100000 for its tile size)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-cli", action="store_true", help="skip the files -> tailseq-import -> files leg (N = 1 only)")
    ap.add_argument("--no-extra", action="store_true", help="skip the HiSeq-lane and estimator-sweep legs (BASELINE configs 3-5)")
    ap.add_argument("--cli-tiles", type=int, default=4)
    if "--workload" in sys.argv and "dedup" in sys.argv:
        sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "tools"))
        import dedup_bench
        known = ap.parse_known_args()[0]
        return dedup_bench.main(["--steps", str(known.steps), "--warmup", str(known.warmup)] +
                                (["--no-cpu-baseline"] if known.no_cpu_baseline else []))
    args = ap.parse_args()
    if args.warmup < 3:
        args.warmup = 3
    if args.impl == "reference":
        return run_reference_arm(args)

    import torch
    import torch.distributed as dist
    from tailseeker_b200 import merge as M
    from tailseeker_b200.config import stock_config
    from tailseeker_b200.importer import TileProcessor, invert_colormatrix
    from tailseeker_b200.synth_torch import make_tile_device

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    local_world = int(os.environ.get("LOCAL_WORLD_SIZE", str(world)))
    host_cores = pin_rank_to_cores(local_rank, local_world) if world > 1 else (os.cpu_count() or 1)
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    comm = None
    if world > 1:
        M.use_torch_nccl()
        dist.init_process_group("nccl", device_id=dev)
        comm = M.NcclComm.from_torch_distributed(local_rank)       # the library's own collective uses it

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(x):
        t = torch.tensor([x], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    cfg = stock_config(args.layout, phix_match_name="PhiX")
    T, L3, bins = cfg.total_cycles, cfg.threep_length, cfg.dist_sampling_bins
    ntiles, N = args.tiles, args.clusters_per_tile

    proc = TileProcessor(cfg, device=local_rank)
    stream = torch.cuda.current_stream()
    proc.set_stream(stream.cuda_stream)
    S = proc.params.num_samples
    merge = M.RunMerge(proc)

    # ---- synthetic tile set, generated in HBM (one lane's worth per rank) -------------------
    tiles = []
    for t in range(ntiles):
        bcl, cif, mtext, pitch = make_tile_device(cfg, N, seed=1000 * (rank + 1) + t, device=dev, phix_frac=PHIX_FRAC)
        mtx = np.array([np.float32(x) for x in mtext.split()], dtype=np.float32)
        tiles.append((bcl, cif, invert_colormatrix(mtx), pitch, mtext))
    torch.cuda.synchronize()

    cut = default_cutoffs(T)
    packed = proc.cutoffs_to_packed(cut)

    def finish_step():
        """run-wide statistics: ONE all-reduce (N > 1) of the packed histograms + counters on the device,
        then the cutoff fit reads the summed histograms where they are (no trip through the host)"""
        merge.add_totals()
        merge.allreduce(comm)
        return merge.fit(positive=0, with_tiles=False)[0]

    def device_step():
        """One pass over the rank's tiles, queued back to back on the stream: nothing about a
        tile is read on the host until the end of the step (the fitted cutoffs and the number of
        measured tails)."""
        merge.reset(7)
        proc.reset_counters()
        proc.ruler_reset(T)
        for (bcl, cif, inv, pitch, _) in tiles:
            proc.upload(bcl.data_ptr(), cif.data_ptr(), inv, nclusters=N, on_device=True,
                        bcl_pitch=pitch, cif_pitch=pitch)
            proc.process_async()
            merge.add_tile()
            proc.ruler_tile_all_async(packed)
        fit = finish_step()
        return proc.ruler_called(), fit

    # ---- warm-up, then K timed steps (device-resident inputs) --------------------------------
    proc.ruler_reset(T)
    sampler = ClockSampler(local_rank)
    sampler.start()
    for _ in range(args.warmup):
        measured, fit = device_step()
    # byte accounting + per-stage timing from a representative pass (tile 0 calls)
    calls0 = proc.calls()       # last tile processed
    ab = algorithmic_bytes(cfg, proc.params, calls0, N, [proc.record_count(s) for s in range(S)])

    proc.profile(True)
    launches0 = proc.launch_count()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    t_begin = time.time()
    e0.record(stream)
    for _ in range(args.steps):
        measured, fit = device_step()
    e1.record(stream)
    barrier()
    clocks = sampler.stop(t_begin, time.time())
    ms_total = e0.elapsed_time(e1)
    launches = proc.launch_count() - launches0
    stage_ms, stage_runs = proc.profile_get()
    proc.profile(False)
    ms_step = max_over_ranks(ms_total) / args.steps
    clusters_step = ntiles * N * world
    value = clusters_step / (ms_step * 1e-3)

    # ---- roofline of the dominant kernel (normalise_score_pack) -------------------------------
    peak, peak_src = measured_peaks()
    k2_ms = stage_ms[2] / max(1, int(stage_runs[2]))
    k2_gbs = ab["k2"] / (k2_ms * 1e-3) / 1e9 if k2_ms > 0 else 0.0
    roof = {"bound": "hbm", "kernel": "k_normalise_score_pack_compact", "achieved": k2_gbs, "peak": peak,
            "unit": "GB/s", "frac": k2_gbs / peak, "traffic": None, "peak_source": peak_src,
            "algorithmic_bytes_per_launch": ab["k2"], "ms_per_launch": k2_ms,
            "stages_ms_per_tile": {
                "decode_demux_seed": stage_ms[0] / max(1, int(stage_runs[0])),
                "slot_scan_assign": stage_ms[1] / max(1, int(stage_runs[1])),
                "normalise_score_pack": k2_ms,
                "fair_sampling_fixups": stage_ms[3] / max(1, int(stage_runs[3])),
                "polya_ruler": stage_ms[4] / max(1, int(stage_runs[4])),
                "fit_cutoffs": stage_ms[5] / max(1, int(stage_runs[5]))},
            "pipeline_algorithmic_gbs_per_gpu": ab["total"] * ntiles / (ms_step * 1e-3) / 1e9,
            "processed_fraction": ab["processed"] / N, "record_slots_per_tile": ab["records"],
            "other_kernels": {}}
    for name, key, st in (("k_decode_demux_seed", "k1", 0), ("k_polya_ruler", "ruler", 4)):
        ms = stage_ms[st] / max(1, int(stage_runs[st]))
        gbs = ab[key] / (ms * 1e-3) / 1e9 if ms > 0 else 0.0
        roof["other_kernels"][name] = {"achieved": gbs, "frac": gbs / peak, "ms_per_launch": ms,
                                       "algorithmic_bytes_per_launch": ab[key]}
    tfile = os.path.join(ROOT, "profiles", "k2_traffic.json")
    if os.path.exists(tfile):
        try:
            tj = json.load(open(tfile))
            roof["traffic"] = tj.get("dram_bytes_per_launch")
            # the bound this kernel really runs against: instruction issue.  Warp-instructions of one
            # launch (ncu smsp__inst_executed.sum of the same build, scaled to this tile's work items)
            # over what 148 SMs x 4 schedulers can issue in the measured launch time.
            wi = tj.get("warp_instructions_per_launch")
            if wi and clocks.get("sm_mhz") and k2_ms > 0:
                wi = wi * ab["processed"] / max(1, tj.get("processed_clusters", ab["processed"]))
                cap = 148 * 4 * clocks["sm_mhz"] * 1e6 * k2_ms * 1e-3
                roof["issue"] = {"bound": "issue", "warp_instructions_per_launch": int(wi), "issue_slots_per_launch": int(cap),
                                 "frac": wi / cap, "source": tj.get("source")}
        except Exception:
            pass

    # ---- end-to-end: host (pinned) buffers through the C ABI ----------------------------------
    e2e = None
    if not args.no_e2e:
        nhost = min(4, ntiles)               # staging ring of this rank: four distinct pinned tiles (10 GB)
        host_tiles = []
        for i in range(nhost):
            bcl, cif, inv, pitch, _ = tiles[i]
            hb = torch.empty((T, N), dtype=torch.uint8, pin_memory=True)
            hc = torch.empty((L3, 4, N), dtype=torch.int16, pin_memory=True)
            hb.copy_(bcl[:, :N]); hc.copy_(cif[:, :, :N])
            host_tiles.append((hb.numpy(), hc.numpy(), inv))
        torch.cuda.synchronize()

        def e2e_step():
            merge.reset(7)
            proc.reset_counters()
            proc.ruler_reset(T)
            d2h = 0
            hb, hc, inv = host_tiles[0]
            proc.upload(hb, hc, inv)                         # H2D of tile 0 (pinned -> HBM)
            for t in range(ntiles):
                if t + 1 < ntiles:                           # H2D of tile t+1 overlaps the kernels of tile t
                    hb, hc, inv = host_tiles[(t + 1) % nhost]
                    proc.upload(hb, hc, inv)
                proc.process()
                merge.add_tile()
                calls = proc.calls()                          # D2H 16 B / cluster
                d2h += calls.nbytes
                lens, _ = proc.ruler_tile_all(packed)         # D2H 2 B / record slot
                d2h += lens.nbytes
            fit = finish_step()                               # D2H of the fitted cutoffs
            got = merge.get()                                 # and of the run-wide histograms / counters
            d2h += fit.nbytes + sum(v.nbytes for v in got.values())
            return d2h

        for _ in range(1):
            e2e_step()
        barrier()
        esteps = max(1, min(args.steps, 3))
        e2, e3 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e2.record(stream)
        for _ in range(esteps):
            d2h_bytes = e2e_step()
        e3.record(stream)
        barrier()
        dt = max_over_ranks(e2.elapsed_time(e3) * 1e-3 / esteps)
        h2d = int(ntiles * N * (T + 8 * L3))
        e2e = {"value": clusters_step / dt, "unit": UNIT,
               "h2d_bytes_per_step": h2d * world,      # whole job, like value
               "d2h_bytes_per_step": int(d2h_bytes) * world, "steps": esteps,
               "h2d_gbs_per_gpu": h2d / dt / 1e9, "host_cores_per_rank": host_cores,
               "note": "H2D of tile t+1 (copy stream) overlaps the kernels of tile t; results read back per tile; "
                       "%d distinct pinned host tiles per rank cycled; ranks pinned to disjoint core sets. The copies "
                       "run at the PCIe rate (54 GB/s per GPU) while a host can feed them: N x that is the ceiling" % nhost}
        del host_tiles

    # ---- CPU baseline beside it (rank 0, N=1 only) --------------------------------------------
    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        from oracle import refrun
        if refrun.available():
            ns = args.ref_clusters if args.ref_clusters > 0 else N
            bcl, cif, inv, pitch, mtext = tiles[0]
            hb = bcl[:, :ns].contiguous().cpu().numpy()
            hc = cif[:, :, :ns].contiguous().cpu().numpy()
            cores = os.cpu_count() or 1
            n, secs = reference_sample_run(cfg, hb, hc, mtext, cores, cut)
            cpu = {"value": n / secs, "unit": UNIT, "cores": cores, "kind": "reference",
                   "sample": "%d clusters of tile 0 (%s): reference tailseq-import (threads=%d, zlib level 1) "
                             "+ tailseq-polya-ruler per sample, files on /dev/shm, binaries' wall time"
                             % (ns, "one full tile" if ns == N else "a part of one tile", cores)}
            del hb, hc
        else:
            cpu = {"value": None, "unit": UNIT, "cores": 0, "kind": "reference",
                   "sample": "oracle/_ref not built"}

    # ---- device-side output encoding of one tile (what the drop-in program's write path runs) ----
    encode = None
    if rank == 0:
        proc.encode_configure()
        bcl, cif, inv, pitch, _ = tiles[0]
        proc.upload(bcl.data_ptr(), cif.data_ptr(), inv, nclusters=N, on_device=True, bcl_pitch=pitch, cif_pitch=pitch)
        proc.process()
        enc_ms = []
        for _ in range(4):
            t0 = time.perf_counter()
            proc.encode_only()
            wall = time.perf_counter() - t0
            enc_ms.append((proc.encode_ms(), wall * 1e3))
        streams = proc.encode()
        out_bytes = sum(len(v) for v in streams.values())
        encode = {"kernels_ms_per_tile": min(e[0] for e in enc_ms), "with_d2h_ms_per_tile": min(e[1] for e in enc_ms),
                  "compressed_bytes_per_tile": out_bytes,
                  "note": "taginfo + seqqual text formatted and all three streams deflated as BGZF blocks on the GPU; "
                          "second figure includes the copy of the compressed streams to pinned host memory"}
        del streams

    # free the MiSeq set before the other legs
    merge.close()
    proc.close()
    del tiles
    torch.cuda.empty_cache()

    # ---- BASELINE configs 3 - 5: HiSeq lane tiles (one lane per GPU), estimator sweep -----------
    extra = None
    if not args.no_extra:
        extra = extra_configs(args, torch, dist, M, dev, rank, world, local_rank, comm, barrier, max_over_ranks)

    cli = None
    if rank == 0 and world == 1 and not args.no_cli:
        try:
            cli = cli_leg(cfg, args.cli_tiles, N, os.cpu_count() or 1)
        except Exception as exc:          # the leg must not take the headline numbers down with it
            cli = {"error": repr(exc)[:300]}

    if rank == 0:
        config = workload_config(args.layout, ntiles, N, T)
        line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
                "warmup": args.warmup, "ms_per_step": ms_step, "higher_is_better": True,
                "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
                "config": config,
                "results": {"polya_called_per_step": int(measured) * world,
                            "cutoffs_fitted": int(np.isfinite(fit).sum()),
                            "collective": "one ncclAllReduce(uint64) of the packed run-wide histograms + counters per step"
                                          if world > 1 else "none (one rank)"},
                "roofline": roof, "cpu_baseline": cpu, "e2e": e2e, "e2e_cli": cli, "encode": encode,
                "other_configs": extra, "gpu_launches": int(launches),
                "clocks": {"sm_mhz": clocks["sm_mhz"], "sm_max_mhz": clocks["sm_max_mhz"],
                           "reasons": clocks["reasons"], "samples": clocks["samples"],
                           "window": clocks.get("window", "timed")}}
        print(json.dumps(line))
    if comm is not None:
        comm.destroy()
    if world > 1:
        dist.destroy_process_group()
    return 0


def extra_configs(args, torch, dist, M, dev, rank, world, local_rank, comm, barrier, max_over_ranks):
    """BASELINE.json configs 3-5 on the same GPUs: (3/4) HiSeq 2500 lane tiles, 309 cycles and 3.1 M
    clusters each, one lane per GPU -- a sample of `ht` tiles of the lane's 64 per rank; (5) the
    parameter-estimation sweep: run-wide cutoffs plus the per-tile cutoffs of a lane's 64 tiles."""
    from tailseeker_b200.config import stock_config
    from tailseeker_b200.importer import TileProcessor, invert_colormatrix
    from tailseeker_b200.synth_torch import make_tile_device
    cfg = stock_config("hiseq", phix_match_name="PhiX")
    T, HN, ht, lane_tiles = cfg.total_cycles, 3_100_000, 3, 64
    proc = TileProcessor(cfg, device=local_rank)
    stream = torch.cuda.current_stream()
    proc.set_stream(stream.cuda_stream)
    merge = M.RunMerge(proc, max_tiles=lane_tiles)
    tiles = []
    for t in range(ht):
        bcl, cif, mtext, pitch = make_tile_device(cfg, HN, seed=50000 + 100 * rank + t, device=dev, phix_frac=PHIX_FRAC)
        mtx = np.array([np.float32(x) for x in mtext.split()], dtype=np.float32)
        tiles.append((bcl, cif, invert_colormatrix(mtx), pitch))
    packed = proc.cutoffs_to_packed(default_cutoffs(T))

    def step(keep_tiles):
        merge.reset(7)
        proc.reset_counters()
        proc.ruler_reset(T)
        for k in range(keep_tiles):
            bcl, cif, inv, pitch = tiles[k % ht]
            proc.upload(bcl.data_ptr(), cif.data_ptr(), inv, nclusters=HN, on_device=True, bcl_pitch=pitch, cif_pitch=pitch)
            proc.process_async()
            merge.add_tile()
            proc.ruler_tile_all_async(packed)
        merge.add_totals()
        merge.allreduce(comm)
        return merge.fit(positive=0, with_tiles=False)

    step(ht)
    proc.profile(True)
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(stream)
    nsteps = 2
    for _ in range(nsteps):
        step(ht)
    e1.record(stream)
    barrier()
    ms = max_over_ranks(e0.elapsed_time(e1)) / nsteps
    stage_ms, stage_runs = proc.profile_get()
    proc.profile(False)
    hiseq = {"workload": "HiSeq 2500 lane tiles (BASELINE.json configs[2]/[3]): %d tiles x %d clusters per GPU of the lane's 64, "
                         "309 cycles, one lane per GPU, device-resident" % (ht, HN),
             "value": ht * HN * world / (ms * 1e-3), "unit": UNIT, "n_gpus": world, "ms_per_step": ms,
             "ms_per_tile": {"decode_demux_seed": stage_ms[0] / max(1, int(stage_runs[0])),
                             "normalise_score_pack": stage_ms[2] / max(1, int(stage_runs[2])),
                             "fair_sampling_fixups": stage_ms[3] / max(1, int(stage_runs[3])),
                             "polya_ruler": stage_ms[4] / max(1, int(stage_runs[4]))},
             "lane_seconds_at_this_rate": lane_tiles * HN / (ht * HN / (ms * 1e-3))}
    # estimator sweep: a lane's 64 tile histograms on every rank (the sample's tiles re-processed to fill
    # them), all-reduced run-wide sums, then run-wide + per-tile cutoffs in one call on device memory
    step(lane_tiles)
    barrier()
    t0 = torch.cuda.Event(enable_timing=True); t1 = torch.cuda.Event(enable_timing=True)
    reps = 3
    t0.record(stream)
    for _ in range(reps):
        cutoffs = merge.fit(positive=0, with_tiles=True)
    t1.record(stream)
    barrier()
    fit_ms = max_over_ranks(t0.elapsed_time(t1)) / reps
    sweep = {"workload": "parameter-estimation sweep (BASELINE.json configs[4]): run-wide + %d per-tile cutoff sets per GPU, "
                         "%d cycles x 1000 bins each, histograms resident on the device" % (lane_tiles, T),
             "ms_per_sweep": fit_ms, "tiles_fitted": lane_tiles * world, "n_gpus": world,
             "cutoff_sets_per_s": (1 + lane_tiles) * world / (fit_ms * 1e-3),
             "cutoffs_found_run_wide": int(np.isfinite(cutoffs[0]).sum())}
    merge.close()
    proc.close()
    del tiles
    torch.cuda.empty_cache()
    return {"hiseq_lane": hiseq, "estimator_sweep": sweep}


if __name__ == "__main__":
    sys.exit(main())
