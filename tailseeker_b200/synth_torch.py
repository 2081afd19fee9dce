"""Device-side synthetic tile generator (bench.py only): the recipe of synth.make_tile()
 written with torch ops so that full-size tiles (~1 M clusters,
2.3 KB each) are produced directly in HBM in well under a second.

Same read construction and distributions as synth.make_tile(), different random stream.
torch is plumbing here (device memory + RNG); nothing in this file is on the measured path.
"""
from __future__ import annotations

import numpy as np
import torch

from .config import ImporterConfig, SPIKEINS_STANDARD

BASE_CODE = {"A": 0, "C": 1, "G": 2, "T": 3}


def _codes(seq: str, device) -> torch.Tensor:
    return torch.tensor([BASE_CODE[c] for c in seq], dtype=torch.uint8, device=device)


def make_tile_device(cfg: ImporterConfig, nclusters: int, seed: int, device="cuda",
                     pitch_align: int = 128, dark_cluster_frac: float = 0.004,
                     dark_cycle_frac: float = 0.0008, nocall_frac: float = 0.002,
                     unknown_frac: float = 0.03, lowqual_frac: float = 0.05,
                     lowdiv_frac: float = 0.01, phix_frac: float = 0.0, chunk: int = 1 << 18):
    """Returns (bcl u8[T][pitch], cif i16[L3][4][pitch], matrix_text, pitch).

    phix_frac: fraction of the unknown-barcode clusters whose 5' read comes from the PhiX174
    control genome (half verbatim, half with substitutions and an occasional indel), like
    synth.overlay_phix_reads()."""
    dev = torch.device(device)
    g = torch.Generator(device=dev)
    g.manual_seed(seed)
    rs = np.random.default_rng(seed)
    T, L3 = cfg.total_cycles, cfg.threep_length
    ts = cfg.threep_start - 1
    istart, ilen = cfg.index_start - 1, cfg.index_length
    N = nclusters
    pitch = (N + pitch_align - 1) // pitch_align * pitch_align

    mtx = np.eye(4) + 0.1 * rs.random((4, 4))
    matrix_text = "\n".join(" ".join("%.6f" % v for v in row) for row in mtx) + "\n"
    mtx32 = torch.tensor(np.array([float(t) for t in matrix_text.split()]).reshape(4, 4),
                         dtype=torch.float32, device=dev)

    bcl = torch.zeros((T, pitch), dtype=torch.uint8, device=dev)
    cif = torch.zeros((L3, 4, pitch), dtype=torch.int16, device=dev)

    nsamp = len(cfg.samples)
    weights = np.array([0.25 if s.fingerprint_seq else 3.0 for s in cfg.samples])
    weights = weights / weights.sum() * (1.0 - unknown_frac)
    probs = torch.tensor(np.concatenate([weights, [unknown_frac]]), dtype=torch.float32, device=dev)
    index_codes = torch.stack([_codes(s.index, dev) for s in cfg.samples])      # [S][ilen]
    design_by_name = {name: d for name, _, _, d in SPIKEINS_STANDARD}
    delim_rel_s = torch.tensor([(s.delimiter_start - 1) if s.delimiter_seq is not None else 15
                                for s in cfg.samples] + [15], dtype=torch.int32, device=dev)
    isspk_s = torch.tensor([bool(s.fingerprint_seq) for s in cfg.samples] + [False], device=dev)
    design_s = torch.tensor([design_by_name.get(s.name, -1) if s.fingerprint_seq else -1
                             for s in cfg.samples] + [-1], dtype=torch.int32, device=dev)
    fp = _codes("ACAGTAGCTC", dev)
    delim = _codes("GTCAG", dev)
    dl = delim.numel()
    motif = _codes("ACGGCATCAGGACTGACCGATGCAAGCTGACGATCGACTGCAGTCAGCTAGC", dev)
    truns = torch.tensor([0, 0, 5, 12, 30, 60, 100, 150, 200], dtype=torch.int32, device=dev)
    pos3 = torch.arange(L3, dtype=torch.int32, device=dev)[:, None]
    if phix_frac > 0:
        from .synth import phix_genome
        fwd = torch.tensor(phix_genome(), dtype=torch.uint8, device=dev)
        strands = torch.stack([fwd, (3 - fwd).flip(0)])                              # [2][5386]
        span = min(51, T)

    def rnd(*shape):
        return torch.rand(shape, generator=g, device=dev)

    def rint(lo, hi, *shape):
        return torch.randint(lo, hi, shape, generator=g, device=dev)

    for c0 in range(0, N, chunk):
        n = min(chunk, N - c0)
        bases = rint(0, 4, T, n).to(torch.uint8)
        sidx = torch.multinomial(probs, n, replacement=True, generator=g)          # nsamp = unknown
        known = sidx < nsamp
        sclamp = torch.clamp(sidx, max=nsamp - 1)

        # index read with 0-3 substitutions
        idx = bases[istart:istart + ilen]
        idx[:, known] = index_codes[sclamp[known]].T
        nsub = torch.multinomial(torch.tensor([0.93, 0.04, 0.02, 0.01], device=dev), n, True, generator=g)
        ar = torch.arange(n, device=dev)
        for k in range(1, 4):
            sel = ar[known & (nsub >= k)]
            where = rint(0, ilen, sel.numel())
            idx[where, sel] = (idx[where, sel] + rint(1, 4, sel.numel()).to(torch.uint8)) & 3

        r3 = bases[ts:ts + L3]
        delim_rel = delim_rel_s[sidx]
        isspk = isspk_s[sidx]
        design = design_s[sidx]
        r3[:fp.numel(), isspk] = fp[:, None]
        fsub = torch.multinomial(torch.tensor([0.93, 0.05, 0.02], device=dev), n, True, generator=g)
        for k in range(1, 3):
            sel = ar[isspk & (fsub >= k)]
            where = rint(0, 10, sel.numel())
            r3[where, sel] = (r3[where, sel] + rint(1, 4, sel.numel()).to(torch.uint8)) & 3

        dmode = torch.multinomial(torch.tensor([0.89, 0.03, 0.03, 0.03, 0.02], device=dev), n, True,
                                  generator=g)
        shift = torch.where(dmode == 1, 1, torch.where(dmode == 2, -1, 0)).to(torch.int32)
        dstart = delim_rel + shift
        rel = pos3 - dstart[None, :]
        indel = (rel >= 0) & (rel < dl) & (dmode != 4)[None, :]
        r3.copy_(torch.where(indel, delim[torch.clamp(rel, 0, dl - 1).long()], r3))
        sel = ar[dmode == 3]
        where = (dstart[sel] + rint(0, dl, sel.numel())).long()
        r3[where, sel] = (r3[where, sel] + rint(1, 4, sel.numel()).to(torch.uint8)) & 3

        mods = torch.multinomial(torch.tensor([0.70, 0.15, 0.08, 0.04, 0.03], device=dev), n, True,
                                 generator=g).to(torch.int32)
        mods[isspk] = 0
        trun = truns[rint(0, truns.numel(), n)]
        trun = torch.where(rnd(n) < 0.15, rint(0, 231, n).to(torch.int32), trun)
        trun = torch.where(design >= 0, design, trun)
        tstart = dstart + dl + mods
        relt = pos3 - tstart[None, :]
        in_t = (relt >= 0) & (relt < trun[None, :])
        seqerr = rnd(L3, n) < 0.01
        r3.copy_(torch.where(in_t & ~seqerr, torch.tensor(3, dtype=torch.uint8, device=dev), r3))
        endpos = tstart + trun
        sel = ar[(endpos < L3) & (trun > 0)]
        r3[endpos[sel].long(), sel] = rint(0, 3, sel.numel()).to(torch.uint8)

        lowdiv = (rnd(n) < lowdiv_frac) & (trun == 0) & ~isspk
        sel = ar[lowdiv]
        if sel.numel():
            for j in range(motif.numel()):
                p = dstart[sel] + dl + j
                good = p < L3
                r3[p[good].long(), sel[good]] = motif[j]

        qual = rint(23, 41, T, n).to(torch.uint8)
        lowq = rnd(n) < lowqual_frac
        nlow = int(lowq.sum())
        if nlow:
            qual[:25, lowq] = rint(2, 24, 25, nlow).to(torch.uint8)
        if phix_frac > 0:
            sel = ar[~known & (rnd(n) < phix_frac)]
            m = sel.numel()
            if m:
                strand = rint(0, 2, m)
                start = rint(0, strands.shape[1] - span - 8, m)
                col = torch.arange(span, device=dev)[None, :]                            # [1][span]
                # one deletion (shift > 0) or insertion-like slip (shift < 0) for 1/8 of the reads
                shift = torch.where(rnd(m) < 0.125, rint(-3, 4, m), torch.zeros(m, dtype=torch.long, device=dev))
                at = rint(8, 40, m)
                src = start[:, None] + col + (col >= at[:, None]) * shift[:, None]
                read = strands[strand[:, None], src.clamp(0, strands.shape[1] - 1)]     # [m][span]
                noisy = rnd(m) < 0.5
                nsub = rint(1, 9, m)
                hit = noisy[:, None] & (rnd(m, span) < (nsub[:, None].float() / 46.0)) & (col < 46)
                read = torch.where(hit, (read + rint(1, 4, m, span).to(torch.uint8)) & 3, read)
                read = torch.where(noisy[:, None] | (shift[:, None] == 0), read,
                                   strands[strand[:, None], start[:, None] + col])
                bases[:span, sel] = read.T.contiguous()
        byte = (qual << 2) | bases
        byte[rnd(T, n) < nocall_frac] = 0
        bcl[:, c0:c0 + n] = byte

        amp = torch.clamp(900.0 + 150.0 * torch.randn(n, generator=g, device=dev), min=100.0)
        decay = torch.exp(-torch.arange(L3, dtype=torch.float32, device=dev) / 400.0)[:, None]
        level = amp[None, :] * decay
        pure = 40.0 + 25.0 * torch.randn((L3, 4, n), generator=g, device=dev)
        onehot = r3[:, None, :] == torch.arange(4, dtype=torch.uint8, device=dev)[None, :, None]
        pure += onehot * level[:, None, :]
        for k in range(1, 6):
            bleed = ((relt - trun[None, :]) == (k - 1)) & (trun[None, :] > 0)
            pure[:, 3, :] += bleed * level * float(0.55 ** k)
        darkc = rnd(n) < dark_cluster_frac
        dark_from = rint(0, L3, n)
        dead = darkc[None, :] & (pos3 >= dark_from[None, :])
        dead |= rnd(L3, n) < dark_cycle_frac
        pure = torch.where(dead[:, None, :], 2.0 * torch.randn((L3, 4, n), generator=g, device=dev), pure)
        raw = torch.einsum("jk,ckn->cjn", mtx32, pure)
        cif[:, :, c0:c0 + n] = torch.clamp(torch.round(raw), -32768, 32767).to(torch.int16)
        del pure, raw, onehot, bases, byte, qual
    return bcl, cif, matrix_text, pitch
