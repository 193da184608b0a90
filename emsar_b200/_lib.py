"""ctypes binding of libemsar_cuda.so (include/emsar_cuda.h). Fails loudly: no CPU fallback exists."""
from __future__ import annotations

import ctypes as C
import os

_PKG = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_PKG, "libemsar_cuda.so")
_lib = None


class EmsarError(RuntimeError):
    def __init__(self, status, what, detail):
        super().__init__(f"{what}: {detail} [status {status}]")
        self.status = status


class IndexDesc(C.Structure):
    _fields_ = [("T", C.c_int32), ("C", C.c_int64), ("class_ptr", C.c_void_p), ("class_tid", C.c_void_p),
                ("nF", C.c_int32), ("euma", C.c_void_p), ("has_node", C.c_void_p),
                ("min_fraglength", C.c_int32), ("max_fraglength", C.c_int32), ("readlength", C.c_int32),
                ("max_t_size", C.c_int32), ("aux", C.c_void_p)]


class IndexInfo(C.Structure):
    _fields_ = [("T", C.c_int32), ("C", C.c_int64), ("nnz", C.c_int64), ("n_multi", C.c_int64), ("nnz_multi", C.c_int64),
                ("n_kseg", C.c_int32), ("max_card", C.c_int32), ("hash_slots", C.c_int64), ("hash_inserted", C.c_int64),
                ("n_sets_nocut", C.c_int32), ("max_set_tids", C.c_int32), ("device_bytes", C.c_int64),
                ("frag_min", C.c_int32), ("frag_max", C.c_int32)]


class DeviceInfo(C.Structure):
    _fields_ = [("sm_count", C.c_int32), ("cc_major", C.c_int32), ("cc_minor", C.c_int32), ("l2_bytes", C.c_int64),
                ("hbm_bytes", C.c_int64), ("em_blocks_per_sm", C.c_int32), ("em_block_threads", C.c_int32),
                ("name", C.c_char * 64)]


class SolveOpts(C.Structure):
    _fields_ = [("eps_abs", C.c_double), ("eps_rel", C.c_double), ("max_iter", C.c_int32), ("delta", C.c_double),
                ("eumacut", C.c_double), ("max_ntid_per_sid", C.c_int32), ("in_model", C.c_void_p), ("sharded", C.c_int32)]


class SolveOut(C.Structure):
    _fields_ = [("fpkm", C.c_void_p), ("efflen", C.c_void_p), ("ireadcount", C.c_void_p), ("ireadcount_int", C.c_void_p),
                ("tpm", C.c_void_p), ("n_iter", C.c_int32), ("final_delta", C.c_double), ("loglik", C.c_double),
                ("total_ireadcount", C.c_int64), ("total_readcount", C.c_int64), ("eumacut", C.c_double),
                ("max_sid", C.c_int32), ("em_ms", C.c_double), ("prep_ms", C.c_double)]


class ModelStats(C.Structure):
    _fields_ = [("T", C.c_int32), ("C_a", C.c_int64), ("nnz_a", C.c_int64), ("rows_short", C.c_int64), ("rows_long", C.c_int64),
                ("rows_hub", C.c_int64), ("rows_fixed", C.c_int64), ("e_tiles", C.c_int64), ("m_tiles", C.c_int64),
                ("bytes_per_iter", C.c_int64), ("stream_bytes_per_iter", C.c_int64),
                ("em_variant", C.c_int32), ("all_local", C.c_int32), ("halo_rows", C.c_int64), ("halo_classes", C.c_int64),
                ("resident_index_bytes", C.c_int64), ("index_bytes", C.c_int64), ("peer_bytes_per_iter", C.c_int64)]


# every symbol include/emsar_cuda.h declares (tests check that the library exports all of them)
SYMBOLS = [
    "emsar_cuda_open", "emsar_cuda_close", "emsar_cuda_strerror", "emsar_cuda_last_error", "emsar_cuda_launch_count",
    "emsar_cuda_synchronize", "emsar_cuda_device_info", "emsar_index_create", "emsar_index_info_get", "emsar_index_aux_get", "emsar_index_destroy",
    "emsar_sample_begin", "emsar_sample_count", "emsar_sample_count_compact", "emsar_sample_count_device", "emsar_sample_counts_set", "emsar_sample_counts_get",
    "emsar_sample_solve", "emsar_sample_segments_get", "emsar_sample_wf_get", "emsar_sample_end", "emsar_sample_prepare",
    "emsar_sample_model_stats", "emsar_sample_em_run", "emsar_sample_theta_get", "emsar_sample_theta_randomize", "emsar_sample_finalize",
    "emsar_host_alloc", "emsar_host_free", "emsar_sample_count_wait", "emsar_comm_unique_id", "emsar_comm_init", "emsar_comm_destroy", "emsar_comm_info", "emsar_sample_counts_allreduce", "emsar_shard_ranges", "emsar_locality_order", "emsar_cuda_timer_start", "emsar_cuda_timer_stop", "emsar_sample_time_adjeuma",
    "emsar_build_classes_run", "emsar_build_classes_free",
]


def lib():
    """Load libemsar_cuda.so (built in-tree by emsar_b200/build.py). Raises if it is missing."""
    global _lib
    if _lib is None:
        if not os.path.exists(_SO):
            raise EmsarError(-1, "libemsar_cuda.so", f"{_SO} is missing: run `python -m emsar_b200.build` (there is no CPU fallback)")
        L = C.CDLL(_SO)
        L.emsar_cuda_strerror.restype = C.c_char_p
        L.emsar_cuda_last_error.restype = C.c_char_p
        _lib = L
    return _lib


def check(status, what):
    if status != 0:
        L = lib()
        raise EmsarError(status, what, f"{L.emsar_cuda_strerror(status).decode()}: {L.emsar_cuda_last_error().decode()}")
