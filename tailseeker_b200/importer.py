"""Host-side mirror of the reference importer / ruler / estimator seam, on top of the C ABI.

`TileProcessor` plays the role of tailseq-import's process() loop body
(src/importer/tailseq-import.c:568-682): upload one block of CIF/BCL data, run the
per-cluster path on the GPU (distribute_processing, :516-565), hand back the same products
the reference writes to disk.  `ruler()` is tailseq-polya-ruler's process_polya_ruling()
(src/polyaruler/polyaruler.c:201-314); `fit_cutoffs()` is find_all_eqodds_point()
(scripts/calculate-optimal-parameters.py:73-78).

PyTorch is not needed here; device memory is owned by the library.  Callers that already
hold the tile in HBM (bench.py, torch tensors) pass raw device pointers with on_device=True.
"""
from __future__ import annotations

import ctypes
from typing import Dict, Optional, Sequence

import numpy as np

from . import _lib
from .config import CALL_DTYPE, CParams, ImporterConfig


def _ptr(a):
    if a is None:
        return None
    if isinstance(a, int):
        return ctypes.c_void_p(a)
    return a.ctypes.data_as(ctypes.c_void_p)


def invert_colormatrix(m: np.ndarray) -> np.ndarray:
    """inverse_4x4_matrix() (src/contrib/misc.c:11-58): float cofactor expansion, same term
    order, so the inverse is bit-identical to the reference's load_color_matrix()."""
    m = np.asarray(m, dtype=np.float32).reshape(16)
    terms = ((+1, -1, 0, 0, -1, +1), (+1, +1, 0, -1, -1, 0), (-1, -1, +1, 0, 0, +1),
             (-1, -1, 0, 0, +1, +1), (-1, +1, 0, -1, +1, 0), (+1, -1, -1, 0, 0, +1))
    adj = np.zeros(16, dtype=np.float32)
    f32 = np.float32
    for i in range(4):
        for j in range(4):
            o = 2 + (j - i)
            ci, rj = i + 4 + o, j + 4 - o
            acc = f32(0)
            for t, (a1, b1, a2, b2, a3, b3) in enumerate(terms):
                e1 = m[((rj + b1) % 4) * 4 + ((ci + a1) % 4)]
                e2 = m[((rj + b2) % 4) * 4 + ((ci + a2) % 4)]
                e3 = m[((rj + b3) % 4) * 4 + ((ci + a3) % 4)]
                prod = f32(f32(e1 * e2) * e3)
                acc = prod if t == 0 else (f32(acc + prod) if t < 3 else f32(acc - prod))
            adj[j * 4 + i] = acc if (o % 2) else f32(-acc)
    det = f32(0)
    for k in range(4):
        det = f32(det + f32(m[k] * adj[k * 4]))
    if det == 0:
        raise ValueError("Color matrix could not be inverted.")
    det = f32(1.0 / float(det))
    return (adj * det).astype(np.float32)


class TileProcessor:
    def __init__(self, cfg: ImporterConfig, device: int = 0, params: Optional[CParams] = None):
        self.cfg = cfg
        self.params = params if params is not None else cfg.to_cparams()
        self.lib = _lib.load()
        self.device = device
        self._ctx = ctypes.c_void_p()
        _lib.check(self.lib.tsk_ctx_create(ctypes.byref(self._ctx), device, ctypes.byref(self.params)))
        self.nclusters = 0          # clusters of the tile last taken by process()
        self._pending = []          # [(nclusters, host arrays kept alive)] uploaded, not yet processed
        self._keep = None

    def close(self):
        if self._ctx:
            self.lib.tsk_ctx_destroy(self._ctx)
            self._ctx = ctypes.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # ------------------------------------------------------------------ importer path
    def upload(self, bcl, cif, invmatrix: np.ndarray, nclusters: Optional[int] = None,
               firstclusterno: int = 0, on_device: bool = False,
               bcl_pitch: Optional[int] = None, cif_pitch: Optional[int] = None):
        """bcl u8[T][pitch], cif i16[L3][4][pitch]: numpy arrays (host) or raw device
        addresses (ints) with on_device=True."""
        inv = np.ascontiguousarray(invmatrix, dtype=np.float32).reshape(16)
        if not on_device:
            bcl = np.ascontiguousarray(bcl, dtype=np.uint8)
            cif = np.ascontiguousarray(cif, dtype=np.int16)
            nclusters = bcl.shape[1] if nclusters is None else nclusters
            bcl_pitch = bcl.shape[1]
            cif_pitch = cif.shape[2]
        _lib.check(self.lib.tsk_tile_upload(self._ctx, _ptr(bcl), bcl_pitch, _ptr(cif), cif_pitch,
                                            _ptr(inv), nclusters, firstclusterno, int(on_device)))
        self._pending.append((int(nclusters), (bcl, cif)))

    def upload_rows(self, bcl: np.ndarray, cif: np.ndarray, invmatrix: np.ndarray, firstclusterno: int = 0,
                    gz_rows: Optional[Dict[int, tuple]] = None, staged: bool = False):
        """The tile cycle file by cycle file (tsk_tile_begin / put / end).  gz_rows: {cycle: (gzip file
        bytes, inflated size, skip)} for base-call rows that travel compressed and are inflated on the
        device; every other row of `bcl` and every cycle of `cif` is sent from the arrays -- through the
        library's page-locked staging slots when `staged` (what the file readers of tailseq-import do)."""
        inv = np.ascontiguousarray(invmatrix, dtype=np.float32).reshape(16)
        bcl = np.ascontiguousarray(bcl, dtype=np.uint8)
        cif = np.ascontiguousarray(cif, dtype=np.int16)
        n = bcl.shape[1]
        gz_rows = gz_rows or {}
        _lib.check(self.lib.tsk_tile_begin(self._ctx, n, firstclusterno, _ptr(inv)))
        keep = []

        def put(row: np.ndarray, send):
            if not staged:
                return send(_ptr(row))
            p, slot = ctypes.c_void_p(), ctypes.c_int32()
            _lib.check(self.lib.tsk_tile_stage_acquire(self._ctx, max(1, row.nbytes), ctypes.byref(p), ctypes.byref(slot)))
            ctypes.memmove(p, _ptr(row), row.nbytes)
            try:
                send(p)
            finally:
                _lib.check(self.lib.tsk_tile_stage_release(self._ctx, slot.value))

        for c in range(bcl.shape[0]):
            if c in gz_rows:
                raw, size, skip = gz_rows[c]
                keep.append(raw)
                _lib.check(self.lib.tsk_tile_put_bcl_gz(self._ctx, c, raw, len(raw), size, skip))
            else:
                put(bcl[c], lambda p, c=c: _lib.check(self.lib.tsk_tile_put_bcl(self._ctx, c, p)))
        for i in range(cif.shape[0]):
            put(cif[i], lambda p, i=i: _lib.check(self.lib.tsk_tile_put_cif(self._ctx, i, p, cif.shape[2])))
        _lib.check(self.lib.tsk_tile_end(self._ctx))
        self._pending.append((int(n), (bcl, cif, keep)))

    def _take(self):
        if self._pending:
            self.nclusters, self._keep = self._pending.pop(0)

    def process(self):
        """Process the oldest uploaded tile (up to two may be uploaded ahead: the copy of the next
        tile then overlaps this tile's kernels)."""
        _lib.check(self.lib.tsk_tile_process(self._ctx))
        self._take()

    def process_async(self):
        """process() without waiting for the device: queue tile after tile (upload / process_async
        / ruler_tile_all_async) and read results later; every getter waits for the device."""
        _lib.check(self.lib.tsk_tile_process_async(self._ctx))
        self._take()

    def process_timed(self, iters: int = 1) -> np.ndarray:
        ms = np.zeros(5, dtype=np.float32)
        _lib.check(self.lib.tsk_tile_process_timed(self._ctx, iters, _ptr(ms)))
        self._take()
        return ms

    def calls(self) -> np.ndarray:
        out = np.zeros(self.nclusters, dtype=CALL_DTYPE)
        _lib.check(self.lib.tsk_tile_get_calls(self._ctx, _ptr(out)))
        return out

    def record_count(self, sample: int) -> int:
        n = ctypes.c_uint32()
        _lib.check(self.lib.tsk_tile_get_record_count(self._ctx, sample, ctypes.byref(n)))
        return n.value

    def records(self, sample: int, drop_withdrawn: bool = True) -> np.ndarray:
        """u32 [nslots][2 + signal_dump_length] record slots of one sample, cluster order."""
        w = 2 + self.params.samples[sample].signal_dump_length
        n = self.record_count(sample)
        out = np.zeros((n, w), dtype=np.uint32)
        _lib.check(self.lib.tsk_tile_get_records(self._ctx, sample, _ptr(out), out.size))
        if drop_withdrawn and n:
            valid = (out[:, 1] >> 16).astype(np.uint16).view(np.int16) >= 0
            out = out[valid]
        return out

    def counters(self) -> np.ndarray:
        out = np.zeros((self.params.num_samples, 6), dtype=np.uint32)
        _lib.check(self.lib.tsk_get_counters(self._ctx, _ptr(out)))
        return out

    def reset_counters(self):
        _lib.check(self.lib.tsk_counters_reset(self._ctx))

    def hist(self):
        shape = (self.params.total_cycles, self.params.dist_sampling_bins)
        pos = np.zeros(shape, dtype=np.uint32)
        neg = np.zeros(shape, dtype=np.uint32)
        _lib.check(self.lib.tsk_get_hist(self._ctx, _ptr(pos), _ptr(neg)))
        return pos, neg

    def device_ptrs(self) -> Dict[str, int]:
        p = [ctypes.c_void_p() for _ in range(4)]
        _lib.check(self.lib.tsk_get_device_ptrs(self._ctx, *[ctypes.byref(x) for x in p]))
        return dict(pos=p[0].value, neg=p[1].value, counters=p[2].value, ruler_pos=p[3].value)

    def synchronize(self):
        _lib.check(self.lib.tsk_synchronize(self._ctx))

    def set_stream(self, cuda_stream: int):
        """Run on the caller's stream (e.g. torch.cuda.current_stream().cuda_stream)."""
        _lib.check(self.lib.tsk_set_stream(self._ctx, ctypes.c_void_p(cuda_stream)))

    def profile(self, on: bool = True):
        _lib.check(self.lib.tsk_profile_enable(self._ctx, int(on)))

    def profile_get(self):
        ms = np.zeros(6, dtype=np.float64)
        runs = np.zeros(6, dtype=np.uint64)
        _lib.check(self.lib.tsk_profile_get(self._ctx, _ptr(ms), _ptr(runs)))
        return ms, runs

    def launch_count(self) -> int:
        return int(self.lib.tsk_launch_count(self._ctx))

    # ------------------------------------------------------------------ ruler path
    def cutoffs_to_packed(self, cutoffs: Sequence[float]) -> np.ndarray:
        return np.array([self.lib.tsk_cutoff_to_packed(float(c)) for c in cutoffs], dtype=np.uint32)

    def ruler_reset(self, ncutoffs: int, bins: Optional[int] = None):
        bins = self.cfg.dist_sampling_bins if bins is None else bins
        _lib.check(self.lib.tsk_ruler_hist_reset(self._ctx, ncutoffs, bins))
        self._ruler_shape = (ncutoffs, bins)

    def ruler_records(self, records: np.ndarray, cutoffs_packed: np.ndarray, min_polya: int = 8,
                      downhill_w: float = 0.49, bins: int = 1000, gap: float = 0.1) -> np.ndarray:
        records = np.ascontiguousarray(records, dtype=np.uint32)
        n, w = records.shape
        cut = np.ascontiguousarray(cutoffs_packed, dtype=np.uint32)
        out = np.full(n, -1, dtype=np.int16)
        _lib.check(self.lib.tsk_ruler_records(self._ctx, _ptr(records), n, w - 2, _ptr(cut), cut.size,
                                              min_polya, downhill_w, bins, gap, _ptr(out), 0))
        return out

    def ruler_tile(self, sample: int, cutoffs_packed: np.ndarray, min_polya: int = 8,
                   downhill_w: float = 0.49, bins: int = 1000, gap: float = 0.1) -> np.ndarray:
        cut = np.ascontiguousarray(cutoffs_packed, dtype=np.uint32)
        out = np.full(self.record_count(sample), -1, dtype=np.int16)
        _lib.check(self.lib.tsk_ruler_tile(self._ctx, sample, _ptr(cut), cut.size, min_polya,
                                           downhill_w, bins, gap, _ptr(out)))
        return out

    def ruler_tile_all(self, cutoffs_packed: np.ndarray, min_polya: int = 8, downhill_w: float = 0.49,
                       bins: int = 1000, gap: float = 0.1):
        """All samples of the resident tile in one launch.  Returns (lengths int16 [sum of slots],
        offsets u32 [num_samples + 1])."""
        cut = np.ascontiguousarray(cutoffs_packed, dtype=np.uint32)
        S = self.params.num_samples
        total = sum(self.record_count(s) for s in range(S))
        out = np.full(max(total, 1), -1, dtype=np.int16)
        offs = np.zeros(S + 1, dtype=np.uint32)
        _lib.check(self.lib.tsk_ruler_tile_all(self._ctx, _ptr(cut), cut.size, min_polya, downhill_w, bins, gap,
                                               _ptr(out), _ptr(offs)))
        return out[:total], offs

    def ruler_tile_all_async(self, cutoffs_packed: np.ndarray, min_polya: int = 8, downhill_w: float = 0.49,
                             bins: int = 1000, gap: float = 0.1) -> None:
        """Enqueue the measurement of every record slot of the resident tile; lengths stay on the
        device (ruler_lengths()), histogram and called counter accumulate."""
        cut = np.ascontiguousarray(cutoffs_packed, dtype=np.uint32)
        _lib.check(self.lib.tsk_ruler_tile_all_async(self._ctx, _ptr(cut), cut.size, min_polya, downhill_w,
                                                     bins, gap))

    def ruler_lengths(self):
        """(lengths int16 [sum of slots], offsets u32 [num_samples + 1]) of the last
        ruler_tile_all_async(); waits for the device."""
        S = self.params.num_samples
        total = sum(self.record_count(s) for s in range(S))
        out = np.full(max(total, 1), -1, dtype=np.int16)
        offs = np.zeros(S + 1, dtype=np.uint32)
        _lib.check(self.lib.tsk_ruler_get_lengths(self._ctx, _ptr(out), _ptr(offs)))
        return out[:total], offs

    def ruler_called(self) -> int:
        """Records measured with a length >= min_polya since the last ruler_reset()."""
        v = ctypes.c_uint64(0)
        _lib.check(self.lib.tsk_ruler_get_called(self._ctx, ctypes.byref(v)))
        return int(v.value)

    def ruler_hist(self) -> np.ndarray:
        out = np.zeros(self._ruler_shape, dtype=np.uint32)
        _lib.check(self.lib.tsk_ruler_get_hist(self._ctx, _ptr(out)))
        return out

    # ---- output encoding on the device (tsk_encode.cu) ----
    def encode_configure(self, layout=None):
        lay = layout if layout is not None else self.cfg.to_encode_layout()
        _lib.check(self.lib.tsk_encode_configure(self._ctx, ctypes.byref(lay)))

    def encode(self):
        """Formats and deflates the outputs of the tile last processed.  Returns
        {(sample index, stream): bytes} with stream 0 seqqual, 1 taginfo, 2 signal records; each value
        is a run of whole BGZF blocks."""
        self.encode_only()
        out = {}
        for s in range(self.params.num_samples):
            for kind in range(3):
                ptr, n = ctypes.c_void_p(), ctypes.c_size_t()
                _lib.check(self.lib.tsk_tile_get_encoded(self._ctx, s, kind, ctypes.byref(ptr), ctypes.byref(n)))
                out[(s, kind)] = ctypes.string_at(ptr.value, n.value) if n.value else b""
        return out

    def encode_only(self):
        """tsk_tile_encode() alone: the streams stay in the library's pinned host buffer."""
        _lib.check(self.lib.tsk_tile_encode(self._ctx))

    def encode_ms(self) -> float:
        ms = ctypes.c_double()
        _lib.check(self.lib.tsk_encode_elapsed_ms(self._ctx, ctypes.byref(ms)))
        return ms.value

    def selftest_deflate(self, data: bytes) -> bytes:
        out = ctypes.create_string_buffer(len(data) + len(data) // 500 + 4096 + 64 * (len(data) // 32768 + 1))
        n = ctypes.c_size_t()
        _lib.check(self.lib.tsk_selftest_deflate(self._ctx, data, len(data), out, len(out), ctypes.byref(n)))
        return out.raw[:n.value]

    def bgzf_eof(self) -> bytes:
        ptr, n = ctypes.c_void_p(), ctypes.c_size_t()
        _lib.check(self.lib.tsk_bgzf_eof_block(ctypes.byref(ptr), ctypes.byref(n)))
        return ctypes.string_at(ptr.value, n.value)

    def selftest_logf(self, x: np.ndarray) -> np.ndarray:
        x = np.ascontiguousarray(x, dtype=np.float32)
        out = np.empty_like(x)
        _lib.check(self.lib.tsk_selftest_logf(self._ctx, _ptr(x), _ptr(out), x.size))
        return out

    def selftest_sweeps(self, div_pairs: int = 1 << 31):
        a = ctypes.c_uint64()
        b = ctypes.c_uint64()
        _lib.check(self.lib.tsk_selftest_sweeps(self._ctx, ctypes.c_uint64(div_pairs),
                                                ctypes.byref(a), ctypes.byref(b)))
        return a.value, b.value

    def force_generic(self, on: bool = True):
        _lib.check(self.lib.tsk_debug_force_generic(self._ctx, int(on)))

    # ------------------------------------------------------------------ estimator
    def fit_cutoffs(self, pos: np.ndarray, neg: np.ndarray, min_spots: Optional[int] = None) -> np.ndarray:
        """Per-cycle equal-odds cutoffs (NaN where undetermined) for one histogram pair."""
        pos = np.ascontiguousarray(pos, dtype=np.uint64)
        neg = np.ascontiguousarray(neg, dtype=np.uint64)
        cycles, bins = pos.shape
        highs = np.array(self.cfg.cutoff_score_search_high, dtype=np.float64)
        out = np.zeros(cycles, dtype=np.float64)
        _lib.check(self.lib.tsk_fit_cutoffs(
            self._ctx, _ptr(pos), _ptr(neg), cycles, bins,
            ctypes.c_uint64(self.cfg.minimum_spots_for_dist_sampling if min_spots is None else min_spots),
            float(self.cfg.kde_bandwidth_for_dist), float(self.cfg.cutoff_score_search_low),
            _ptr(highs), highs.size, _ptr(out)))
        return out
