"""TEST INFRASTRUCTURE ONLY -- CPU restatement of tailseq-dedup-perfect
(reference src/deduplicator/tailseq-dedup-perfect.c), used by tests/ as the parity checker.

Pinned: tests/test_dedup.py compares it with the reference binary (oracle/_ref/tailseq-dedup-perfect,
built from the unmodified source by oracle/Makefile) when that is present and with the committed
outputs of that binary in tests/golden/dedup_*.  It follows the reference statement by statement
(the queue of tqueue_append(), :168-214, is a Python list)."""
from __future__ import annotations

import gzip
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
REF_BIN = os.path.join(HERE, "_ref", "tailseq-dedup-perfect")


def priority(flags: int) -> int:
    """calculate_tag_prority(), :87-105; flag bits from src/sigproc-flags.h:30-41."""
    return ((flags & 0x0010 == 0) * 1024 + (flags & 0x0080 == 0) * 512 + (flags & 0x0002 == 0) * 256 +
            (flags & 0x0008 != 0) * 128 + (flags & 0x0800 == 0) * 64 + (flags & 0x0200 == 0) * 32 +
            (flags & 0x0400 == 0) * 16 + (flags & 0x0040 == 0) * 8 + (flags & 0x0020 == 0) * 4 +
            (flags & 0x0001 != 0) * 2)


def _roundf(x: np.float32) -> int:
    """(int)roundf(x): half away from zero."""
    x = float(x)
    return int(np.floor(abs(x) + 0.5)) * (1 if x >= 0 else -1)


def dedup_perfect(data: bytes):
    """stdin bytes -> (stdout bytes, trace bytes), main() :348-416."""
    out, trace = [], []
    queue = []                  # entries: (tile, cluster, flags, polya, mods, head)
    state = {"ndup": 0, "high": -1, "group": 0, "umi": None, "size": 2048}

    def flush():                # process_tag_duplicates(), :277-345
        if not queue:
            return
        if len(queue) == 1:
            t = queue[0]
            out.append(t[5] + b"\t%d\n" % state["ndup"])
            trace.append(b"%s\t%d\t%d\t%d\n" % (t[0], t[1], state["group"], t[3]))
            state["group"] += 1
        else:
            total = sum(t[3] for t in queue)
            mean = np.float32(total) / np.float32(len(queue))
            best, bestd = 0, abs(np.float32(queue[0][3]) - mean)
            for k in range(1, len(queue)):
                d = abs(np.float32(queue[k][3]) - mean)
                if d < bestd:
                    best, bestd = k, d
            rep = queue[best]
            final = _roundf(mean) if rep[3] >= 0 else -1
            out.append(b"%s\t%d\t%d\t%d\t%s\t%d\n" % (rep[0], rep[1], rep[2], final, rep[4], state["ndup"]))
            for k, t in enumerate(queue):
                trace.append(b"%s\t%d\t%d\t%d\n" % (t[0], t[1], state["group"], final if k == best else -2))
            state["group"] += 1
        queue.clear()

    for line in data.split(b"\n")[:-1] if data.endswith(b"\n") else data.split(b"\n"):
        f = line.split(b"\t")
        if len(f) != 6:
            break
        cur = (f[0], int(f[1]), int(f[2]), int(f[3]), f[4], line[:line.rindex(b"\t")])
        if queue and f[5] != state["umi"]:
            flush()
            state["ndup"], state["high"] = 0, -1
        state["umi"] = f[5]
        prio = priority(cur[2])
        state["ndup"] += 1
        if not queue or prio > state["high"]:                       # :178-194
            for t in queue:
                trace.append(b"%s\t%d\t%d\t-3\n" % (t[0], t[1], state["group"]))
            queue.clear()
            queue.append(cur)
            state["high"] = prio
        else:
            if len(queue) == state["size"] - 1:                     # :196-198
                # tqueue_expand() (:142-166) copies the queued entries but not the one being
                # appended (it sits at `front`, outside tqueue_foreach), so that record comes
                # out of the calloc'ed new array as zeros / empty strings; the size stays grown
                # for the rest of the run
                state["size"] = int(np.float32(state["size"]) * np.float32(1.5))
                cur = (b"", 0, 0, 0, b"", b"")
            if prio >= state["high"]:                               # :201-204
                queue.append(cur)
            else:                                                   # :205-211
                trace.append(b"%s\t%d\t%d\t-3\n" % (cur[0], cur[1], state["group"]))
    flush()
    return b"".join(out), b"".join(trace)


def ref_available() -> bool:
    return os.access(REF_BIN, os.X_OK)


def run_reference(data: bytes, workdir: str, binary: str = REF_BIN, extra_args=()):
    """The reference binary (or a drop-in with the same argv) on `data`: (stdout, trace)."""
    tr = os.path.join(workdir, "trace.gz")
    if os.path.exists(tr):
        os.unlink(tr)
    r = subprocess.run([binary, *extra_args, tr], input=data, capture_output=True)
    if r.returncode != 0:
        raise RuntimeError("%s failed (%d): %s" % (binary, r.returncode, r.stderr.decode()))
    with gzip.open(tr, "rb") as f:
        return r.stdout, f.read()


def synth_taginfo(n, seed=0, umi_len=30, n_umis=None, tiles=("a1101", "a1102", "b2101"), big_groups=(),
                  with_n=True, sort=True):
    """Synthetic taginfo lines.  sort=True: ordered like `sort -k6,6 -k1,1 -k2,2n` (C locale);
    otherwise concatenated tiles in (tile, cluster) order."""
    rng = np.random.default_rng(seed)
    n_umis = n_umis or max(1, int(n * 0.7))
    alphabet = np.frombuffer(b"ACGT", np.uint8)
    pool = alphabet[rng.integers(0, 4, (n_umis, umi_len))]
    if with_n:
        pool[rng.random(pool.shape) < 0.002] = ord("N")
    which = rng.integers(0, n_umis, n)
    at = 0
    for size in big_groups:                      # a few heavily duplicated UMIs
        which[at:at + size] = rng.integers(0, n_umis)
        at += size
    rng.shuffle(which)
    flag_bits = np.array([0x1, 0x2, 0x4, 0x8, 0x10, 0x20, 0x40, 0x80, 0x100, 0x200, 0x400, 0x800])
    probs = np.array([0.7, 0.1, 0.3, 0.4, 0.15, 0.05, 0.1, 0.02, 0.05, 0.03, 0.03, 0.02])
    flags = ((rng.random((n, 12)) < probs) * flag_bits).sum(1)
    polya = np.where(rng.random(n) < 0.1, -1, rng.integers(0, 231, n))
    polya = np.where(rng.random(n) < 0.3, rng.integers(0, 12, n), polya)
    tile_ix = np.sort(rng.integers(0, len(tiles), n))
    cluster = np.zeros(n, np.int64)
    for t in range(len(tiles)):
        m = tile_ix == t
        cluster[m] = np.sort(rng.choice(np.arange(1, 4 * m.sum() + 2), m.sum(), replace=False))
    mods = [b"", b"T", b"TT", b"G", b"GU".replace(b"U", b"T"), b"CA"]
    rows = []
    for i in range(n):
        rows.append((pool[which[i]].tobytes(), tiles[tile_ix[i]].encode(), int(cluster[i]), int(flags[i]),
                     int(polya[i]), mods[int(rng.integers(0, len(mods)))]))
    if sort:
        rows.sort(key=lambda r: (r[0], r[1], r[2]))
    return b"".join(b"%s\t%d\t%d\t%d\t%s\t%s\n" % (r[1], r[2], r[3], r[4], r[5], r[0]) for r in rows)
