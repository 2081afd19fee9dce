"""Run-wide sums on the devices: ctypes front end of the tsk_merge_* / tsk_nccl_* entry points
(csrc/tsk_merge.cu).  One `RunMerge` per process (= per GPU): it accumulates the histograms of the
tiles its TileProcessor handles, all-reduces them over NCCL in one collective and fits the cutoffs
on device memory -- the sums the reference forms over per-tile files in
scripts/calculate-optimal-parameters.py:139-140, scripts/measure-polya-lengths.py:64-70 and
scripts/stats-merge-demultiplexing-counts.py:33-53."""
from __future__ import annotations

import ctypes
import os
from typing import Optional, Sequence

import numpy as np

from . import _lib


def use_torch_nccl() -> None:
    """Point the library at the NCCL that ships with torch (TSK_NCCL_LIB) unless the caller chose
    one: both work, this merely avoids two NCCL versions in one process."""
    if os.environ.get("TSK_NCCL_LIB"):
        return
    try:
        import nvidia.nccl  # type: ignore
        cand = os.path.join(list(nvidia.nccl.__path__)[0], "lib", "libnccl.so.2")
        if os.path.exists(cand):
            os.environ["TSK_NCCL_LIB"] = cand
    except Exception:
        pass


class NcclComm:
    """An ncclComm_t made by the library itself.  Multi-process: rank 0 calls unique_id() and hands
    the 128 bytes to every rank (torch.distributed.broadcast_object_list, a file, MPI ...)."""

    def __init__(self, handle: int):
        self.handle = ctypes.c_void_p(handle)

    @staticmethod
    def unique_id() -> bytes:
        buf = ctypes.create_string_buffer(128)
        _lib.check(_lib.load().tsk_nccl_unique_id(buf))
        return buf.raw

    @classmethod
    def init_rank(cls, device: int, nranks: int, rank: int, uid: bytes) -> "NcclComm":
        h = ctypes.c_void_p()
        _lib.check(_lib.load().tsk_nccl_init_rank(ctypes.byref(h), device, nranks, rank, uid))
        return cls(h.value)

    @classmethod
    def init_all(cls, devices: Sequence[int]):
        n = len(devices)
        arr = (ctypes.c_void_p * n)()
        devs = (ctypes.c_int * n)(*devices)
        _lib.check(_lib.load().tsk_nccl_init_all(arr, n, devs))
        return [cls(arr[i]) for i in range(n)]

    @classmethod
    def from_torch_distributed(cls, device: int) -> Optional["NcclComm"]:
        """Inside an initialised torch.distributed job: a communicator over the same ranks (the
        unique id travels through the existing process group).  None for a single rank."""
        import torch.distributed as dist
        if not dist.is_available() or not dist.is_initialized() or dist.get_world_size() == 1:
            return None
        box = [cls.unique_id() if dist.get_rank() == 0 else None]
        dist.broadcast_object_list(box, src=0)
        return cls.init_rank(device, dist.get_world_size(), dist.get_rank(), box[0])

    def destroy(self):
        if self.handle:
            _lib.load().tsk_nccl_destroy(self.handle)
            self.handle = None


def group_start():
    _lib.check(_lib.load().tsk_nccl_group_start())


def group_end():
    _lib.check(_lib.load().tsk_nccl_group_end())


class RunMerge:
    def __init__(self, proc, max_tiles: int = 0):
        """proc: the TileProcessor whose tiles are summed (same device, same shape)."""
        self.lib = _lib.load()
        self.proc = proc
        p = proc.params
        self.cycles, self.bins, self.nsamples = p.total_cycles, p.dist_sampling_bins, p.num_samples
        self._m = ctypes.c_void_p()
        _lib.check(self.lib.tsk_merge_create(ctypes.byref(self._m), proc.device, self.cycles, self.bins,
                                             self.nsamples, max_tiles))

    def close(self):
        if self._m:
            self.lib.tsk_merge_destroy(self._m)
            self._m = ctypes.c_void_p()

    def add_tile(self):
        _lib.check(self.lib.tsk_merge_add_tile(self._m, self.proc._ctx))

    def add_totals(self):
        _lib.check(self.lib.tsk_merge_add_totals(self._m, self.proc._ctx))

    def allreduce(self, comm: Optional[NcclComm]):
        if comm is not None:
            _lib.check(self.lib.tsk_merge_allreduce(self._m, self.proc._ctx, comm.handle))

    def reset(self, what: int = 7):
        _lib.check(self.lib.tsk_merge_reset(self._m, self.proc._ctx, what))

    def get(self):
        hw = (self.cycles, self.bins)
        pos, neg, rul = (np.empty(hw, dtype=np.uint64) for _ in range(3))
        cnt = np.empty((self.nsamples, 6), dtype=np.uint64)
        _lib.check(self.lib.tsk_merge_get(self._m, self.proc._ctx, pos.ctypes.data, neg.ctypes.data,
                                          rul.ctypes.data, cnt.ctypes.data))
        return dict(pos=pos, neg=neg, ruler_pos=rul, counters=cnt)

    def tile_count(self) -> int:
        return self.lib.tsk_merge_tile_count(self._m)

    def fit(self, positive: int = 0, min_spots: int = 300, bandwidth: float = 0.1, low: float = 0.0,
            highs: Sequence[float] = (0.4, 0.6), with_tiles: bool = True) -> np.ndarray:
        """Cutoffs [1 + kept tiles][cycles] (row 0 run-wide); NaN = undetermined."""
        rows = 1 + self.tile_count()
        out = np.empty(rows * self.cycles, dtype=np.float64)
        hs = np.asarray(highs, dtype=np.float64)
        _lib.check(self.lib.tsk_merge_fit(self._m, self.proc._ctx, positive, min_spots, bandwidth, low,
                                          hs.ctypes.data, len(hs), out.ctypes.data))
        return out.reshape(-1, self.cycles)[: rows if with_tiles else 1]
