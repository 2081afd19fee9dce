"""ctypes binding of libtailseeker_b200.so -- exactly the symbols of include/tailseeker_b200.h.

There is no fallback: if the shared library (CUDA kernels for sm_100a) is missing or no
B200-class device is usable, calls raise."""
from __future__ import annotations

import ctypes
import os

from .config import CCall, CEncodeLayout, CParams

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, "libtailseeker_b200.so")

# every symbol include/tailseeker_b200.h declares (tests check the library exports them all)
SYMBOLS = [
    "tsk_last_error", "tsk_probe", "tsk_ctx_create", "tsk_ctx_destroy", "tsk_tile_upload",
    "tsk_tile_process", "tsk_tile_get_calls", "tsk_tile_get_record_count", "tsk_tile_get_records",
    "tsk_get_counters", "tsk_counters_reset", "tsk_get_hist", "tsk_get_device_ptrs",
    "tsk_cutoff_to_packed", "tsk_ruler_records", "tsk_ruler_tile", "tsk_ruler_tile_all", "tsk_ruler_hist_reset",
    "tsk_ruler_get_hist", "tsk_fit_cutoffs", "tsk_synchronize", "tsk_tile_process_timed",
    "tsk_launch_count", "tsk_selftest_logf", "tsk_set_stream", "tsk_profile_enable", "tsk_profile_get",
    "tsk_selftest_sweeps", "tsk_debug_force_generic", "tsk_tile_process_async",
    "tsk_ruler_tile_all_async", "tsk_ruler_get_lengths", "tsk_ruler_get_called",
    "tsk_dedup_pack_umi", "tsk_dedup_create", "tsk_dedup_destroy", "tsk_dedup_upload", "tsk_dedup_run",
    "tsk_dedup_get_records", "tsk_dedup_get_groups", "tsk_dedup_get_lost", "tsk_dedup_elapsed_ms", "tsk_dedup_launch_count",
    "tsk_encode_configure", "tsk_encode_reserve", "tsk_tile_encode", "tsk_tile_get_encoded", "tsk_encode_elapsed_ms", "tsk_bgzf_eof_block", "tsk_selftest_deflate", "tsk_bgzf_create", "tsk_bgzf_destroy", "tsk_bgzf_deflate", "tsk_bgzf_launch_count", "tsk_tile_begin", "tsk_tile_put_bcl", "tsk_tile_put_bcl_gz", "tsk_tile_put_cif", "tsk_tile_end", "tsk_tile_stage_acquire", "tsk_tile_stage_release",
    "tsk_nccl_unique_id", "tsk_nccl_init_rank", "tsk_nccl_init_all", "tsk_nccl_destroy", "tsk_nccl_group_start",
    "tsk_nccl_group_end", "tsk_hist_allreduce", "tsk_merge_create", "tsk_merge_destroy", "tsk_merge_add_tile",
    "tsk_merge_add_totals", "tsk_merge_allreduce", "tsk_merge_get", "tsk_merge_reset", "tsk_merge_tile_count", "tsk_merge_fit",
]

_lib = None


class TskError(RuntimeError):
    pass


def load():
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise TskError("libtailseeker_b200.so is not built (run `python -m tailseeker_b200.build` "
                       "or __graft_entry__.build()); there is no CPU fallback")
    L = ctypes.CDLL(LIB_PATH)
    vp, i32, u32, sz, f32, f64 = (ctypes.c_void_p, ctypes.c_int, ctypes.c_uint32, ctypes.c_size_t,
                                  ctypes.c_float, ctypes.c_double)
    L.tsk_last_error.restype = ctypes.c_char_p
    L.tsk_probe.argtypes = [i32]
    L.tsk_ctx_create.argtypes = [ctypes.POINTER(vp), i32, ctypes.POINTER(CParams)]
    L.tsk_ctx_destroy.argtypes = [vp]
    L.tsk_tile_upload.argtypes = [vp, vp, sz, vp, sz, vp, u32, u32, i32]
    L.tsk_tile_process.argtypes = [vp]
    L.tsk_tile_process_async.argtypes = [vp]
    L.tsk_tile_get_calls.argtypes = [vp, vp]
    L.tsk_tile_get_record_count.argtypes = [vp, i32, ctypes.POINTER(u32)]
    L.tsk_tile_get_records.argtypes = [vp, i32, vp, sz]
    L.tsk_get_counters.argtypes = [vp, vp]
    L.tsk_counters_reset.argtypes = [vp]
    L.tsk_get_hist.argtypes = [vp, vp, vp]
    L.tsk_get_device_ptrs.argtypes = [vp, ctypes.POINTER(vp), ctypes.POINTER(vp), ctypes.POINTER(vp),
                                      ctypes.POINTER(vp)]
    L.tsk_cutoff_to_packed.argtypes = [f64]
    L.tsk_cutoff_to_packed.restype = u32
    L.tsk_ruler_records.argtypes = [vp, vp, sz, i32, vp, i32, i32, f32, i32, f32, vp, i32]
    L.tsk_ruler_tile.argtypes = [vp, i32, vp, i32, i32, f32, i32, f32, vp]
    L.tsk_ruler_tile_all.argtypes = [vp, vp, i32, i32, f32, i32, f32, vp, vp]
    L.tsk_ruler_hist_reset.argtypes = [vp, i32, i32]
    L.tsk_ruler_tile_all_async.argtypes = [vp, vp, i32, i32, f32, i32, f32]
    L.tsk_ruler_get_lengths.argtypes = [vp, vp, vp]
    L.tsk_ruler_get_called.argtypes = [vp, ctypes.POINTER(ctypes.c_uint64)]
    L.tsk_ruler_get_hist.argtypes = [vp, vp]
    u64 = ctypes.c_uint64
    L.tsk_dedup_pack_umi.argtypes = [ctypes.c_char_p, sz, ctypes.POINTER(u64)]
    L.tsk_dedup_create.argtypes = [ctypes.POINTER(vp), i32, u64]
    L.tsk_dedup_destroy.argtypes = [vp]
    L.tsk_dedup_destroy.restype = None
    L.tsk_dedup_upload.argtypes = [vp, u64, vp, vp, vp, vp]
    L.tsk_dedup_run.argtypes = [vp, i32, ctypes.POINTER(u64)]
    L.tsk_dedup_get_records.argtypes = [vp, vp, vp, vp, vp]
    L.tsk_dedup_get_groups.argtypes = [vp, vp, vp, vp, vp]
    L.tsk_dedup_get_lost.argtypes = [vp, vp, u32, ctypes.POINTER(u32)]
    L.tsk_dedup_elapsed_ms.argtypes = [vp, ctypes.POINTER(f32), ctypes.POINTER(f32)]
    L.tsk_dedup_launch_count.argtypes = [vp]
    L.tsk_dedup_launch_count.restype = u64
    L.tsk_encode_configure.argtypes = [vp, ctypes.POINTER(CEncodeLayout)]
    L.tsk_tile_encode.argtypes = [vp]
    L.tsk_tile_get_encoded.argtypes = [vp, i32, i32, ctypes.POINTER(vp), ctypes.POINTER(sz)]
    L.tsk_encode_elapsed_ms.argtypes = [vp, ctypes.POINTER(f64)]
    L.tsk_selftest_deflate.argtypes = [vp, vp, sz, vp, sz, ctypes.POINTER(sz)]
    L.tsk_nccl_unique_id.argtypes = [vp]
    L.tsk_nccl_init_rank.argtypes = [ctypes.POINTER(vp), i32, i32, i32, vp]
    L.tsk_nccl_init_all.argtypes = [ctypes.POINTER(vp), i32, ctypes.POINTER(i32)]
    L.tsk_nccl_destroy.argtypes = [vp]
    L.tsk_hist_allreduce.argtypes = [vp, vp]
    L.tsk_merge_create.argtypes = [ctypes.POINTER(vp), i32, i32, i32, i32, i32]
    L.tsk_merge_destroy.argtypes = [vp]
    L.tsk_merge_destroy.restype = None
    L.tsk_merge_add_tile.argtypes = [vp, vp]
    L.tsk_merge_add_totals.argtypes = [vp, vp]
    L.tsk_merge_allreduce.argtypes = [vp, vp, vp]
    L.tsk_merge_get.argtypes = [vp, vp, vp, vp, vp, vp]
    L.tsk_merge_reset.argtypes = [vp, vp, i32]
    L.tsk_merge_tile_count.argtypes = [vp]
    L.tsk_merge_fit.argtypes = [vp, vp, i32, ctypes.c_uint64, f64, f64, vp, i32, vp]
    L.tsk_tile_begin.argtypes = [vp, u32, u32, vp]
    L.tsk_tile_put_bcl.argtypes = [vp, i32, vp]
    L.tsk_tile_put_bcl_gz.argtypes = [vp, i32, vp, sz, ctypes.c_uint64, ctypes.c_uint64]
    L.tsk_tile_put_cif.argtypes = [vp, i32, vp, sz]
    L.tsk_tile_end.argtypes = [vp]
    L.tsk_encode_reserve.argtypes = [vp, u32]
    L.tsk_tile_stage_acquire.argtypes = [vp, sz, ctypes.POINTER(vp), ctypes.POINTER(i32)]
    L.tsk_tile_stage_release.argtypes = [vp, i32]
    L.tsk_bgzf_create.argtypes = [ctypes.POINTER(vp), i32]
    L.tsk_bgzf_destroy.argtypes = [vp]
    L.tsk_bgzf_destroy.restype = None
    L.tsk_bgzf_deflate.argtypes = [vp, vp, sz, ctypes.POINTER(vp), ctypes.POINTER(sz)]
    L.tsk_bgzf_launch_count.argtypes = [vp]
    L.tsk_bgzf_launch_count.restype = ctypes.c_uint64
    L.tsk_bgzf_eof_block.argtypes = [ctypes.POINTER(vp), ctypes.POINTER(sz)]
    L.tsk_fit_cutoffs.argtypes = [vp, vp, vp, i32, i32, ctypes.c_uint64, f64, f64, vp, i32, vp]
    L.tsk_synchronize.argtypes = [vp]
    L.tsk_tile_process_timed.argtypes = [vp, i32, vp]
    L.tsk_launch_count.argtypes = [vp]
    L.tsk_launch_count.restype = ctypes.c_uint64
    L.tsk_selftest_logf.argtypes = [vp, vp, vp, sz]
    L.tsk_selftest_sweeps.argtypes = [vp, ctypes.c_uint64, vp, vp]
    L.tsk_debug_force_generic.argtypes = [vp, i32]
    L.tsk_set_stream.argtypes = [vp, vp]
    L.tsk_profile_enable.argtypes = [vp, i32]
    L.tsk_profile_get.argtypes = [vp, vp, vp]
    for name in SYMBOLS:
        fn = getattr(L, name)
        if fn.restype is ctypes.c_int and name not in ("tsk_probe",):
            pass
    _lib = L
    return L


def check(rc: int) -> None:
    if rc != 0:
        raise TskError(load().tsk_last_error().decode() or "tailseeker_b200 call failed")
