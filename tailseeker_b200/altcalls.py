"""[alternative_calls]: FASTQ files that replace the base calls of a stretch of cycles.

Python mirror of the host reader in csrc/host/tileio.c (read_fastq_override), used by the tests and
by callers that assemble `bcl[T][N]` themselves before `TileProcessor.upload`.  Behaviour follows
the reference (src/importer/altcalls.c:35-44,131-196; tailseq-import.c:203-230; list order
parseconfig.c:181-182):

* one four-line record per cluster, in tile order; every sequence and quality line must have the
  read length of the first record;
* byte = base code (C 1, G 2, T/U 3, anything else -- A, N, ... -- 0) | (quality char - 33) << 2,
  truncated to eight bits.  An `N` therefore turns into an `A` unless its quality is `!` (byte 0);
* entries are applied newest INI entry first, so where two overlap the EARLIER INI entry wins;
* the BCL files of overridden cycles are never read.

This is input decoding on the host; the per-cluster work stays on the GPU.
"""
from __future__ import annotations

import gzip
from typing import Iterable, Sequence, Tuple

import numpy as np

_CODE = np.zeros(256, np.uint8)
for _ch, _v in (("C", 1), ("G", 2), ("T", 3), ("U", 3)):
    _CODE[ord(_ch)] = _CODE[ord(_ch.lower())] = _v


class AltCallError(ValueError):
    """The FASTQ is not usable; the message is the one the programs print."""


def _open(path: str):
    with open(path, "rb") as f:
        magic = f.read(2)
    return gzip.open(path, "rb") if magic == b"\x1f\x8b" else open(path, "rb")


def read_fastq_calls(path: str, nclusters: int | None = None) -> np.ndarray:
    """-> u8[readlength][N] base-call bytes of every record of `path` (gzip or plain).
    `nclusters` given: the file must hold exactly that many records."""
    with _open(path) as f:
        lines = f.read().split(b"\n")
    if lines and lines[-1] == b"":
        lines.pop()
    else:                       # no newline at the very end: the programs drop the last character
        if lines:
            lines[-1] = lines[-1][:-1]
    if not lines:
        raise AltCallError("No sequence available in %s." % path)
    if len(lines) % 4:
        if len(lines) % 4 == 1 and not lines[-1].startswith(b"@"):
            raise AltCallError("%s is not in FASTQ format." % path)
        raise AltCallError("Unexpected EOF while reading %s" % path)
    n = len(lines) // 4
    if any(not h.startswith(b"@") for h in lines[0::4]) or any(not h.startswith(b"+") for h in lines[2::4]):
        raise AltCallError("%s is not in FASTQ format." % path)
    L = len(lines[1])
    if any(len(x) != L for x in lines[1::4]) or any(len(x) != L for x in lines[3::4]):
        raise AltCallError("Sequence length in the FASTQ is inconsistent.")
    if nclusters is not None and n < nclusters:
        raise AltCallError("Not enough sequences from %s" % path)
    if nclusters is not None and n > nclusters:
        raise AltCallError("Extra sequences found in %s" % path)
    seq = np.frombuffer(b"".join(lines[1::4]), np.uint8).reshape(n, L)
    qual = np.frombuffer(b"".join(lines[3::4]), np.uint8).reshape(n, L)
    calls = _CODE[seq] | ((qual - np.uint8(33)) << np.uint8(2))
    return np.ascontiguousarray(calls.T.astype(np.uint8))


def overlay(bcl: np.ndarray, entries: Sequence[Tuple[int, str]]) -> np.ndarray:
    """`bcl` u8[T][N] with the cycles named by `entries` -- (first cycle 1-based, FASTQ path) in INI
    order -- replaced.  Returns a new array."""
    out = np.array(bcl, dtype=np.uint8, copy=True)
    T, N = out.shape
    for first, path in reversed(list(entries)):          # the reference's list is newest first
        if first <= 0:
            raise AltCallError("Alternative call position is out of range.")
        calls = read_fastq_calls(path, N)
        if first - 1 + calls.shape[0] > T:
            raise AltCallError("Alternative calls in %s run past the last cycle (%d)." % (path, T))
        out[first - 1:first - 1 + calls.shape[0]] = calls
    return out


def overridden_cycles(entries: Iterable[Tuple[int, int]], total_cycles: int) -> np.ndarray:
    """bool[T]: cycles whose BCL file is never opened; `entries` = (first cycle 1-based, read length)."""
    m = np.zeros(total_cycles, bool)
    for first, length in entries:
        m[first - 1:first - 1 + length] = True
    return m


def write_fastq(path: str, calls: np.ndarray, prefix: str = "r") -> None:
    """Inverse of read_fastq_calls for bytes whose base is not N (tests, synthetic data):
    `calls` u8[L][N] -> FASTQ, gzip'ed when `path` ends in .gz."""
    calls = np.asarray(calls, np.uint8)
    seq = np.frombuffer(b"ACGT", np.uint8)[calls & 3].T
    qual = ((calls >> 2) + 33).astype(np.uint8).T
    body = b"".join(b"@%s%d\n%s\n+\n%s\n" % (prefix.encode(), i, seq[i].tobytes(), qual[i].tobytes())
                    for i in range(seq.shape[0]))
    with (gzip.open(path, "wb", compresslevel=1) if path.endswith(".gz") else open(path, "wb")) as f:
        f.write(body)
