"""TEST INFRASTRUCTURE ONLY -- ctypes front end of oracle/libtsk_oracle.so (tsk_oracle.c).

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline leg may import this."""
from __future__ import annotations

import ctypes
import os
import subprocess
from typing import Dict, Optional

import numpy as np

from tailseeker_b200.config import (CALL_DTYPE, COUNTER_FIELDS, CParams, ImporterConfig)

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, "libtsk_oracle.so")
_lib = None


def build(force: bool = False) -> None:
    src = os.path.join(HERE, "tsk_oracle.c")
    if (force or not os.path.exists(LIB_PATH)
            or os.path.getmtime(LIB_PATH) < os.path.getmtime(src)):
        subprocess.check_call(["make", "-s", "-C", HERE, "oracle"])


def lib():
    global _lib
    if _lib is None:
        build()
        L = ctypes.CDLL(LIB_PATH)
        L.tsk_oracle_process.restype = ctypes.c_int
        L.tsk_oracle_ruler.restype = ctypes.c_int
        L.tsk_oracle_invert_matrix.restype = ctypes.c_int
        L.tsk_oracle_cutoff_to_packed.restype = ctypes.c_uint32
        L.tsk_oracle_cutoff_to_packed.argtypes = [ctypes.c_double]
        L.tsk_oracle_logf.restype = ctypes.c_float
        L.tsk_oracle_logf.argtypes = [ctypes.c_float]
        _lib = L
    return _lib


def _ptr(a: np.ndarray):
    return a.ctypes.data_as(ctypes.c_void_p)


def invert_matrix(m: np.ndarray) -> np.ndarray:
    m = np.ascontiguousarray(m, dtype=np.float32).reshape(16)
    out = np.empty(16, dtype=np.float32)
    if lib().tsk_oracle_invert_matrix(_ptr(m), _ptr(out)) != 0:
        raise ValueError("singular colour matrix")
    return out


def cutoffs_to_packed(cutoffs) -> np.ndarray:
    return np.array([lib().tsk_oracle_cutoff_to_packed(float(c)) for c in cutoffs],
                    dtype=np.uint32)


def process(cfg: ImporterConfig, bcl: np.ndarray, cif: np.ndarray, invmatrix: np.ndarray,
            firstclusterno: int = 0, want_scores: bool = False,
            params: Optional[CParams] = None) -> Dict:
    """Oracle run of one block.  Returns dict(calls, records{sample idx: [n][2+dump]},
    counters [S][6], pos, neg, scores?)."""
    p = params if params is not None else cfg.to_cparams()
    T, L3 = p.total_cycles, p.threep_length
    assert bcl.dtype == np.uint8 and bcl.ndim == 2 and bcl.shape[0] == T
    assert cif.dtype == np.int16 and cif.shape[:2] == (L3, 4)
    N = bcl.shape[1]
    bcl = np.ascontiguousarray(bcl)
    cif = np.ascontiguousarray(cif)
    S = p.num_samples
    calls = np.zeros(N, dtype=CALL_DTYPE)
    maxdump = max(p.samples[s].signal_dump_length for s in range(S))
    cap = max(1, N) * (2 + maxdump)
    records = np.zeros(cap, dtype=np.uint32)
    rec_offset = np.zeros(S, dtype=np.uint64)
    rec_count = np.zeros(S, dtype=np.uint32)
    counters = np.zeros((S, 6), dtype=np.uint32)
    bins = p.dist_sampling_bins
    pos = np.zeros((T, bins), dtype=np.uint32)
    neg = np.zeros((T, bins), dtype=np.uint32)
    scores = np.empty((N, L3), dtype=np.float32) if want_scores else None
    inv = np.ascontiguousarray(invmatrix, dtype=np.float32).reshape(16)
    rc = lib().tsk_oracle_process(
        ctypes.byref(p), _ptr(bcl), ctypes.c_size_t(bcl.shape[1]), _ptr(cif),
        ctypes.c_size_t(cif.shape[2]), _ptr(inv), ctypes.c_uint32(N),
        ctypes.c_uint32(firstclusterno), _ptr(calls), _ptr(records), ctypes.c_size_t(cap),
        _ptr(rec_offset), _ptr(rec_count), _ptr(counters), _ptr(pos), _ptr(neg),
        _ptr(scores) if scores is not None else None)
    if rc != 0:
        raise RuntimeError("tsk_oracle_process failed")
    recs = {}
    for s in range(S):
        w = 2 + p.samples[s].signal_dump_length
        o, c = int(rec_offset[s]), int(rec_count[s])
        recs[s] = records[o:o + c * w].reshape(c, w).copy()
    return dict(calls=calls, records=recs, counters=counters, pos=pos, neg=neg, scores=scores)


def ruler(records: np.ndarray, cutoffs_packed: np.ndarray, min_polya: int = 8,
          downhill_w: float = 0.49, bins: int = 1000, gap: float = 0.1,
          pos: Optional[np.ndarray] = None):
    """records: u32 [n][2+dump_len].  Returns (int16 lengths [n], pos hist [ncut][bins])."""
    records = np.ascontiguousarray(records, dtype=np.uint32)
    n, w = records.shape
    cut = np.ascontiguousarray(cutoffs_packed, dtype=np.uint32)
    if pos is None:
        pos = np.zeros((cut.size, bins), dtype=np.uint32)
    out = np.empty(n, dtype=np.int16)
    rc = lib().tsk_oracle_ruler(_ptr(records), ctypes.c_size_t(n), ctypes.c_int(w - 2),
                                _ptr(cut), ctypes.c_int(cut.size), ctypes.c_int(min_polya),
                                ctypes.c_float(downhill_w), ctypes.c_int(bins),
                                ctypes.c_float(gap), _ptr(out), _ptr(pos))
    if rc != 0:
        raise RuntimeError("tsk_oracle_ruler failed")
    return out, pos


def logf(x: float) -> np.float32:
    return np.float32(lib().tsk_oracle_logf(ctypes.c_float(float(x))))
