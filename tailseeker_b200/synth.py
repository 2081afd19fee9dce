"""Seeded synthetic Illumina tiles of the flowcell shapes in BASELINE.json.

Produces the two file-native arrays the hot path consumes
  bcl  u8 [total_cycles][N]      (reference src/importer/bclreader.c:95-119,146-166:
                                  bits0-1 base ACGT, bits2-7 quality, byte 0 = no-call)
  cif  i16[threep_length][4][N]  (reference src/importer/cifreader.c:118-170: one file per
                                  cycle, four channel planes A,C,G,T)
plus the 4x4 crosstalk matrix text of BaseCalls/Matrix/s_{lane}_{read}_{tile}_matrix.txt
(reference signalproc.c:345-375), following the recipe of SURVEY.md section 8(d).

Reads are built so that every branch of process_spots() (spotanalyzer.c:617-755) is hit:
index substitutions, unknown barcodes, spike-in fingerprints, shifted / mismatched / missing
delimiters, terminal modifications, T-runs of many lengths (seed poly(A)), low-quality
5' ends (balancer QC quirk), biased balancers, sporadic dark cycles and fully dark clusters.
"""
from __future__ import annotations

import os
import struct
from dataclasses import dataclass
from typing import Optional

import numpy as np

from .config import ImporterConfig

BASE_CODE = {"A": 0, "C": 1, "G": 2, "T": 3}
TRUN_CHOICES = np.array([0, 0, 5, 12, 30, 60, 100, 150, 200])


@dataclass
class Tile:
    bcl: np.ndarray          # u8 [T][N]
    cif: np.ndarray          # i16 [L3][4][N]
    matrix_text: str         # 16 numbers, row major
    nclusters: int
    truth_sample: np.ndarray  # index into cfg.samples, -1 = unknown barcode

    @property
    def matrix(self) -> np.ndarray:
        """The 16 floats as fscanf("%f") reads them (signalproc.c:358-365)."""
        return np.array([np.float32(tok) for tok in self.matrix_text.split()],
                        dtype=np.float32)


def _codes(seq: str) -> np.ndarray:
    return np.array([BASE_CODE[c] for c in seq], dtype=np.uint8)


def make_tile(cfg: ImporterConfig, nclusters: int, seed: int,
              dark_cluster_frac: float = 0.004, dark_cycle_frac: float = 0.0008,
              nocall_frac: float = 0.002, unknown_frac: float = 0.03,
              lowqual_frac: float = 0.05, lowdiv_frac: float = 0.01,
              spikein_weight: float = 0.25, phix_frac: float = 0.0,
              chunk: int = 1 << 16) -> Tile:
    """Deterministic in (cfg, nclusters, seed).  Memory: ~2.3 KB per cluster.

    phix_frac: fraction of the unknown-barcode clusters whose 5' read is drawn from the PhiX174
    control genome (see overlay_phix_reads); 0 leaves the random stream of earlier fixtures
    untouched."""
    rng = np.random.default_rng(seed)
    T, L3 = cfg.total_cycles, cfg.threep_length
    ts = cfg.threep_start - 1
    istart, ilen = cfg.index_start - 1, cfg.index_length
    N = nclusters

    mtx = np.eye(4) + 0.1 * rng.random((4, 4))
    matrix_text = "\n".join(" ".join("%.6f" % v for v in row) for row in mtx) + "\n"
    mtx32 = np.array([float(t) for t in matrix_text.split()]).reshape(4, 4)

    bcl = np.empty((T, N), dtype=np.uint8)
    cif = np.empty((L3, 4, N), dtype=np.int16)
    truth = np.empty(N, dtype=np.int32)

    nsamp = len(cfg.samples)
    # spike-in standards are a small admixture: 7 x 0.25 against 4 x 3 (about 13 % of clusters)
    weights = np.array([spikein_weight if s.fingerprint_seq else 3.0 for s in cfg.samples])
    weights = weights / weights.sum() * (1.0 - unknown_frac)
    probs = np.concatenate([weights, [unknown_frac]])
    index_codes = np.stack([_codes(s.index) for s in cfg.samples]) if nsamp else \
        np.zeros((0, ilen), np.uint8)
    pos3 = np.arange(L3, dtype=np.int32)[:, None]

    for c0 in range(0, N, chunk):
        n = min(chunk, N - c0)
        bases = rng.integers(0, 4, size=(T, n), dtype=np.uint8)
        sidx = rng.choice(nsamp + 1, size=n, p=probs)
        sidx = np.where(sidx == nsamp, -1, sidx).astype(np.int32)
        truth[c0:c0 + n] = sidx

        # ---- index read, with 0-3 substitutions
        known = sidx >= 0
        idx = bases[istart:istart + ilen]
        idx[:, known] = index_codes[sidx[known]].T
        nsub = rng.choice(4, size=n, p=[0.93, 0.04, 0.02, 0.01])
        for k in range(1, 4):
            sel = np.where(known & (nsub >= k))[0]
            where = rng.integers(0, ilen, size=sel.size)
            idx[where, sel] = (idx[where, sel] + rng.integers(1, 4, size=sel.size)) & 3

        # ---- 3' read layout per cluster
        delim_rel = np.full(n, 15, dtype=np.int32)
        limit = np.full(n, L3, dtype=np.int32)
        design = np.full(n, -1, dtype=np.int32)
        fp_len = np.zeros(n, dtype=np.int32)
        r3 = bases[ts:ts + L3]
        for si, s in enumerate(cfg.samples):
            m = sidx == si
            if not m.any():
                continue
            if s.delimiter_seq is not None:
                delim_rel[m] = s.delimiter_start - 1
            if s.fingerprint_seq:
                fpc = _codes(s.fingerprint_seq)
                fp_len[m] = len(fpc)
                r3[:len(fpc), m] = fpc[:, None]
        # spike-in design lengths follow the sample order of SPIKEINS_STANDARD when present
        from .config import SPIKEINS_STANDARD
        design_by_name = {name: d for name, _, _, d in SPIKEINS_STANDARD}
        for si, s in enumerate(cfg.samples):
            if s.name in design_by_name and s.fingerprint_seq:
                design[sidx == si] = design_by_name[s.name]

        # fingerprint substitutions (spike-ins): 5 % one, 2 % two
        isspk = fp_len > 0
        fsub = rng.choice(3, size=n, p=[0.93, 0.05, 0.02])
        for k in range(1, 3):
            sel = np.where(isspk & (fsub >= k))[0]
            where = rng.integers(0, 10, size=sel.size)
            r3[where, sel] = (r3[where, sel] + rng.integers(1, 4, size=sel.size)) & 3

        # delimiter: shifted +-1 (3 % each), one substitution (3 %), destroyed (2 %)
        dmode = rng.choice(5, size=n, p=[0.89, 0.03, 0.03, 0.03, 0.02])
        shift = np.where(dmode == 1, 1, np.where(dmode == 2, -1, 0)).astype(np.int32)
        dstart = delim_rel + shift
        delim = _codes("GTCAG")
        dl = len(delim)
        rel = pos3 - dstart[None, :]
        indel = (rel >= 0) & (rel < dl) & (dmode != 4)[None, :]
        r3[...] = np.where(indel, delim[np.clip(rel, 0, dl - 1)], r3)
        sel = np.where(dmode == 3)[0]
        where = dstart[sel] + rng.integers(0, dl, size=sel.size)
        r3[where, sel] = (r3[where, sel] + rng.integers(1, 4, size=sel.size)) & 3

        # terminal modifications then the T-run (poly(A) read as T on the 3' read)
        mods = rng.choice(5, size=n, p=[0.70, 0.15, 0.08, 0.04, 0.03]).astype(np.int32)
        mods[isspk] = 0
        trun = TRUN_CHOICES[rng.integers(0, len(TRUN_CHOICES), size=n)].astype(np.int32)
        rnd = rng.random(n) < 0.15
        trun = np.where(rnd, rng.integers(0, 231, size=n), trun).astype(np.int32)
        trun = np.where(design >= 0, design, trun)
        tstart = dstart + dl + mods
        relt = pos3 - tstart[None, :]
        in_t = (relt >= 0) & (relt < trun[None, :])
        seqerr = rng.random((L3, n)) < 0.01
        r3[...] = np.where(in_t & ~seqerr, np.uint8(3), r3)
        # the base right after the T-run is never T, so the designed length is the run length
        endpos = tstart + trun
        ok = (endpos < L3) & (trun > 0)
        sel = np.where(ok)[0]
        r3[endpos[sel], sel] = rng.integers(0, 3, size=sel.size).astype(np.uint8)

        # low-diversity inserts (exercise the fair-sampling cap, spotanalyzer.c:460-489)
        lowdiv = (rng.random(n) < lowdiv_frac) & (trun == 0) & ~isspk
        sel = np.where(lowdiv)[0]
        if sel.size:
            motif = _codes("ACGGCATCAGGACTGACCGATGCAAGCTGACGATCGACTGCAGTCAGCTAGC")
            for j, code in enumerate(motif):
                p = dstart[sel] + dl + j
                good = p < L3
                r3[p[good], sel[good]] = code

        # ---- qualities and no-calls
        qual = rng.integers(23, 41, size=(T, n), dtype=np.uint8)
        lowq = rng.random(n) < lowqual_frac
        nlow = int(lowq.sum())
        if nlow:
            qual[:25, lowq] = rng.integers(2, 24, size=(25, nlow), dtype=np.uint8)
        byte = ((qual << 2) | bases).astype(np.uint8)
        byte[rng.random((T, n)) < nocall_frac] = 0
        bcl[:, c0:c0 + n] = byte

        # ---- intensities of the 3' read
        amp = np.clip(rng.normal(900.0, 150.0, size=n), 100.0, None).astype(np.float32)
        decay = np.exp(-np.arange(L3, dtype=np.float32) / 400.0)[:, None]
        level = amp[None, :] * decay                                   # [L3][n]
        pure = rng.normal(40.0, 25.0, size=(L3, 4, n)).astype(np.float32)
        onehot = (r3[:, None, :] == np.arange(4, dtype=np.uint8)[None, :, None])
        pure += onehot * level[:, None, :]
        # T signal bleeds a few cycles past the end of the run (phasing-like tail)
        for k in range(1, 6):
            bleed = (relt - trun[None, :] == (k - 1)) & (trun[None, :] > 0)
            pure[:, 3, :] += bleed * level * np.float32(0.55 ** k)
        darkc = rng.random(n) < dark_cluster_frac
        dark_from = rng.integers(0, L3, size=n)
        dead = darkc[None, :] & (pos3 >= dark_from[None, :])
        dead |= rng.random((L3, n)) < dark_cycle_frac
        pure = np.where(dead[:, None, :], rng.normal(0.0, 2.0, size=(L3, 4, n)), pure)
        raw = np.einsum("jk,ckn->cjn", mtx32.astype(np.float32), pure.astype(np.float32))
        cif[:, :, c0:c0 + n] = np.clip(np.rint(raw), -32768, 32767).astype(np.int16)
        del pure, raw, onehot

    if phix_frac > 0:
        overlay_phix_reads(bcl, np.where(truth < 0)[0], phix_frac, seed)
    return Tile(bcl=bcl, cif=cif, matrix_text=matrix_text, nclusters=N, truth_sample=truth)


def phix_genome() -> np.ndarray:
    """Forward strand of the PhiX174 control as base codes 0-3, unpacked from
    csrc/phix174.h (the table the library itself is built from)."""
    import re
    path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "csrc", "phix174.h")
    text = open(path).read()
    length = int(re.search(r"TSK_PHIX174_LENGTH (\d+)", text).group(1))
    words = np.array([int(w, 16) for w in re.findall(r"0x([0-9a-f]{8})u", text)], dtype=np.uint64)
    codes = (words[:, None] >> (2 * np.arange(16, dtype=np.uint64))[None, :]) & np.uint64(3)
    return codes.reshape(-1)[:length].astype(np.uint8)


def overlay_phix_reads(bcl: np.ndarray, candidates: np.ndarray, frac: float, seed: int,
                       span: int = 51) -> np.ndarray:
    """Rewrite the base bits of cycles [0, span) of a random `frac` of `candidates` with reads
    from either PhiX strand.  Half are verbatim (the exact-substring path of
    controlaligner.c:117-119); the rest carry 1-8 substitutions and, for a quarter, one 1-3 nt
    insertion or deletion, so that local-alignment scores land on both sides of the
    acceptance threshold (:135).  No-call bytes stay no-calls.  Returns the chosen clusters."""
    rng = np.random.default_rng([seed, 0x50486958])
    fwd = phix_genome()
    rev = (3 - fwd)[::-1]
    span = min(span, bcl.shape[0])
    chosen = candidates[rng.random(candidates.size) < frac]
    for n in chosen:
        strand = fwd if rng.random() < 0.5 else rev
        start = int(rng.integers(0, strand.size - span - 8))
        read = strand[start:start + span + 8].copy()
        if rng.random() >= 0.5:
            if rng.random() < 0.25:
                at, ln = int(rng.integers(8, 40)), int(rng.integers(1, 4))
                if rng.random() < 0.5:
                    read = np.concatenate([read[:at], rng.integers(0, 4, ln).astype(np.uint8), read[at:]])
                else:
                    read = np.concatenate([read[:at], read[at + ln:]])
            nsub = int(rng.integers(1, 9))
            where = rng.integers(0, 46, size=nsub)
            read[where] = (read[where] + rng.integers(1, 4, size=nsub)) & 3
        col = bcl[:span, n]
        bcl[:span, n] = np.where(col == 0, 0, (col & 0xfc) | read[:span]).astype(np.uint8)
    return chosen


# --------------------------------------------------------------------------------------
# file writers: the on-disk layout the reference binaries open
# --------------------------------------------------------------------------------------
def write_tile_files(tile: Tile, cfg: ImporterConfig, root: str) -> str:
    """Lay the tile out under `root` as an Illumina run folder fragment:
    L%03d/C%d.1/s_%d_%04d.cif (cifreader.c:235), BaseCalls/L%03d/C%d.1/s_%d_%04d.bcl
    (bclreader.c:206), BaseCalls/Matrix/s_{lane}_3_{tile}_matrix.txt.  Returns the matrix path."""
    lane, tno = cfg.lane, cfg.tile
    T, L3, ts = cfg.total_cycles, cfg.threep_length, cfg.threep_start - 1
    N = tile.nclusters
    for c in range(T):
        d = os.path.join(root, "BaseCalls", "L%03d" % lane, "C%d.1" % (c + 1))
        os.makedirs(d, exist_ok=True)
        with open(os.path.join(d, "s_%d_%04d.bcl" % (lane, tno)), "wb") as f:
            f.write(struct.pack("<I", N))
            f.write(np.ascontiguousarray(tile.bcl[c]).tobytes())
    for i in range(L3):
        cyc = ts + i + 1
        d = os.path.join(root, "L%03d" % lane, "C%d.1" % cyc)
        os.makedirs(d, exist_ok=True)
        with open(os.path.join(d, "s_%d_%04d.cif" % (lane, tno)), "wb") as f:
            f.write(b"CIF" + struct.pack("<BBHHI", 1, 2, cyc, 1, N))
            f.write(np.ascontiguousarray(tile.cif[i]).astype("<i2").tobytes())
    mdir = os.path.join(root, "BaseCalls", "Matrix")
    os.makedirs(mdir, exist_ok=True)
    mpath = os.path.join(mdir, "s_%d_3_%04d_matrix.txt" % (lane, tno))
    with open(mpath, "w") as f:
        f.write(tile.matrix_text)
    return mpath
