"""Host mirror of tailseq-dedup-perfect (reference src/deduplicator/tailseq-dedup-perfect.c) on
top of the C ABI (tsk_dedup_*): parsing of taginfo lines as parse_line() (:216-274) reads them,
UMI keys, and the two text outputs (stdout lines :286-287,324-326; duplicate trace
:184-186,208-210,290-291,332-335).  The grouping / collapsing itself runs on the GPU only."""
from __future__ import annotations

import ctypes

import numpy as np

from . import _lib

_CODE = np.zeros(256, np.uint64)
for _ch, _v in zip(b"ACGNT", (1, 2, 3, 4, 5)):
    _CODE[_ch] = _v
MAX_UMI = 42


def pack_umis(umis):
    """3 bits per base, first base most significant, 21 bases per word (include/tailseeker_b200.h)."""
    n = len(umis)
    hi = np.zeros(n, np.uint64)
    lo = np.zeros(n, np.uint64)
    for i, u in enumerate(umis):
        b = u if isinstance(u, bytes) else u.encode()
        if len(b) > MAX_UMI:
            raise ValueError("UMI longer than %d bases" % MAX_UMI)
        c = _CODE[np.frombuffer(b, np.uint8)]
        if (c == 0).any():
            raise ValueError("UMI %r has a character other than ACGNT" % (u,))
        for w, arr in ((0, hi), (1, lo)):
            part = c[21 * w:21 * (w + 1)]
            v = 0
            for r, x in enumerate(part):
                v |= int(x) << (60 - 3 * r)
            arr[i] = v
    return hi, lo


def parse_taginfo(data: bytes):
    """Lines `tile\\tcluster\\tflags\\tpolyA\\tmods\\tUMI\\n` -> dict of columns."""
    lines = data.split(b"\n")
    if lines and lines[-1] == b"":
        lines.pop()
    cols = [ln.split(b"\t") for ln in lines]
    for c in cols:
        if len(c) != 6:
            raise ValueError("Could not parse line: %r" % (b"\t".join(c),))
    return {
        "tile": [c[0] for c in cols],
        "cluster": np.array([int(c[1]) for c in cols], np.int64),
        "flags": np.array([int(c[2]) for c in cols], np.int32),
        "polya": np.array([int(c[3]) for c in cols], np.int32),
        "mods": [c[4] for c in cols],
        "umi": [c[5] for c in cols],
        "head": [ln[:ln.rindex(b"\t")] for ln in lines],       # the line without its UMI column
    }


class Deduplicator:
    """One tsk_dedup object (device buffers for `capacity` records)."""

    def __init__(self, capacity: int, device: int = 0):
        self.L = _lib.load()
        self.h = ctypes.c_void_p()
        if self.L.tsk_dedup_create(ctypes.byref(self.h), device, int(capacity)) != 0:
            raise _lib.TskError(self.L.tsk_last_error().decode())
        self.n = 0
        self.ngroups = 0

    def close(self):
        if self.h:
            self.L.tsk_dedup_destroy(self.h)
            self.h = ctypes.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _check(self, rc):
        if rc != 0:
            raise _lib.TskError(self.L.tsk_last_error().decode())

    def upload(self, umi_hi, umi_lo, flags, polya, n=None):
        """NumPy arrays (host) or integer device addresses with `n`."""
        if n is None:
            n = len(flags)
            keep = [np.ascontiguousarray(umi_hi, np.uint64), np.ascontiguousarray(umi_lo, np.uint64),
                    np.ascontiguousarray(flags, np.int32), np.ascontiguousarray(polya, np.int32)]
            ptrs = [a.ctypes.data for a in keep]
        else:
            ptrs = [int(umi_hi), int(umi_lo), int(flags), int(polya)]
        self._check(self.L.tsk_dedup_upload(self.h, int(n), *ptrs))
        self.n = int(n)

    def run(self, sort_by_umi=False):
        g = ctypes.c_uint64()
        self._check(self.L.tsk_dedup_run(self.h, 1 if sort_by_umi else 0, ctypes.byref(g)))
        self.ngroups = g.value
        return self.ngroups

    def records(self):
        n = self.n
        out = {"order": np.empty(n, np.uint32), "group": np.empty(n, np.uint32),
               "trace_value": np.empty(n, np.int32), "trace_slot": np.empty(n, np.uint32)}
        self._check(self.L.tsk_dedup_get_records(self.h, out["order"].ctypes.data, out["group"].ctypes.data,
                                                 out["trace_value"].ctypes.data, out["trace_slot"].ctypes.data))
        return out

    def groups(self):
        g = self.ngroups
        out = {"rep": np.empty(g, np.uint32), "polya": np.empty(g, np.int32),
               "ndup": np.empty(g, np.uint32), "kept": np.empty(g, np.uint32)}
        self._check(self.L.tsk_dedup_get_groups(self.h, out["rep"].ctypes.data, out["polya"].ctypes.data,
                                                out["ndup"].ctypes.data, out["kept"].ctypes.data))
        return out

    def lost(self):
        """Positions of the records the reference loses when its queue grows (see the header)."""
        cnt = ctypes.c_uint32()
        buf = np.zeros(64, np.uint32)
        self._check(self.L.tsk_dedup_get_lost(self.h, buf.ctypes.data, 64, ctypes.byref(cnt)))
        return buf[:cnt.value].copy()

    def elapsed_ms(self):
        a, b = ctypes.c_float(), ctypes.c_float()
        self._check(self.L.tsk_dedup_elapsed_ms(self.h, ctypes.byref(a), ctypes.byref(b)))
        return a.value, b.value

    def launches(self):
        return int(self.L.tsk_dedup_launch_count(self.h))


def format_outputs(cols, rec, grp, input_order=False, lost=()):
    """(stdout bytes, trace bytes) of the reference program from the device results.
    input_order: emit the stdout lines in the order of the representatives in the input (what
    the pipeline's second `sort -k1,1 -k2,2n` produces when the input came in that order)."""
    order = rec["order"]
    if len(lost):                         # records the reference carries on as zeros / empty strings
        cols = {k: (list(v) if isinstance(v, list) else v.copy()) for k, v in cols.items()}
        for p in lost:
            r = int(order[p])
            cols["tile"][r] = cols["mods"][r] = cols["head"][r] = b""
            cols["cluster"][r] = cols["flags"][r] = 0
    out = []
    gs = range(len(grp["rep"]))
    if input_order:
        gs = sorted(gs, key=lambda g: int(order[grp["rep"][g]]))
    for g in gs:
        r = int(order[grp["rep"][g]])
        if grp["kept"][g] == 1:
            out.append(cols["head"][r] + b"\t%d\n" % grp["ndup"][g])
        else:
            out.append(b"%s\t%d\t%d\t%d\t%s\t%d\n" % (cols["tile"][r], cols["cluster"][r], cols["flags"][r],
                                                     grp["polya"][g], cols["mods"][r], grp["ndup"][g]))
    n = len(order)
    at = np.empty(n, np.int64)
    at[rec["trace_slot"]] = np.arange(n)
    trace = []
    for slot in range(n):
        p = int(at[slot])
        r = int(order[p])
        trace.append(b"%s\t%d\t%d\t%d\n" % (cols["tile"][r], cols["cluster"][r], rec["group"][p],
                                            rec["trace_value"][p]))
    return b"".join(out), b"".join(trace)
