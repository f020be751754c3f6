"""ctypes binding of libsparkfm_b200.so -- exactly the symbols include/sparkfm_b200.h declares.

The library is built in-tree (sparkfm_b200/libsparkfm_b200.so) by `make -C sparkfm_b200/csrc`
(see __graft_entry__.build).  There is no fallback: if the shared object is missing, loading
raises, and every compute entry point fails with SFM_ERR_CUDA when no GPU is visible.
"""
from __future__ import annotations

import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
SO_PATH = os.environ.get("SFM_LIB") or os.path.join(_HERE, "libsparkfm_b200.so")  # SFM_LIB: A/B builds

SFM_ABI_VERSION = 1
SFM_OK = 0
SFM_ERR_ARG, SFM_ERR_CUDA, SFM_ERR_NCCL, SFM_ERR_OOM = -1, -2, -3, -4
SFM_ERR_INDEX, SFM_ERR_IO, SFM_ERR_STATE = -5, -6, -7
SFM_UNIQUE_ID_BYTES = 128


class SfmConfig(C.Structure):
    _fields_ = [
        ("abi_version", C.c_int32), ("task", C.c_int32), ("k", C.c_int32), ("k0", C.c_int32),
        ("k1", C.c_int32), ("device", C.c_int32), ("n_slots", C.c_int64),
        ("reg0", C.c_float), ("regw", C.c_float), ("regv", C.c_float),
        ("step_size", C.c_float), ("mini_batch_fraction", C.c_float), ("sampler_mode", C.c_int32),
        ("sampler_seed", C.c_uint64),
    ]


class SfmStats(C.Structure):
    _fields_ = [
        ("train_steps", C.c_int64), ("train_rows", C.c_int64), ("train_nnz", C.c_int64),
        ("predict_rows", C.c_int64), ("predict_nnz", C.c_int64), ("kernel_launches", C.c_int64),
        ("h2d_bytes", C.c_int64), ("d2h_bytes", C.c_int64),
        ("ms_forward", C.c_double), ("ms_sort", C.c_double), ("ms_reduce", C.c_double),
        ("ms_allreduce", C.c_double), ("ms_update", C.c_double), ("ms_predict", C.c_double),
        ("ms_total_train", C.c_double),
    ]


_vp = C.c_void_p
_H = C.c_void_p  # sfm_handle*
_i64p, _i32p = C.POINTER(C.c_int64), C.POINTER(C.c_int32)
_f32p, _f64p = C.POINTER(C.c_float), C.POINTER(C.c_double)
_u8p, _u32p = C.POINTER(C.c_uint8), C.POINTER(C.c_uint32)

# name -> (restype, argtypes); mirrors include/sparkfm_b200.h one to one
SIGNATURES = {
    "sfm_abi_version": (C.c_int32, []),
    "sfm_status_string": (C.c_char_p, [C.c_int32]),
    "sfm_device_count": (C.c_int32, []),
    "sfm_host_alloc": (C.c_int32, [C.POINTER(_vp), C.c_uint64]),
    "sfm_host_free": (C.c_int32, [_vp]),
    "sfm_create": (C.c_int32, [C.POINTER(SfmConfig), C.POINTER(_H)]),
    "sfm_destroy": (C.c_int32, [_H]),
    "sfm_last_error": (C.c_char_p, [_H]),
    "sfm_get_config": (C.c_int32, [_H, C.POINTER(SfmConfig)]),
    "sfm_set_hyper": (C.c_int32, [_H, C.c_float, C.c_float, C.c_float, C.c_float, C.c_float]),
    "sfm_init_model": (C.c_int32, [_H, C.c_double, C.c_double, C.c_uint64]),
    "sfm_set_model": (C.c_int32, [_H, C.c_float, _f32p, _f32p]),
    "sfm_get_model": (C.c_int32, [_H, _f32p, _f32p, _f32p]),
    "sfm_set_model_f64": (C.c_int32, [_H, C.c_double, _f64p, _f64p]),
    "sfm_get_model_f64": (C.c_int32, [_H, _f64p, _f64p, _f64p]),
    "sfm_save": (C.c_int32, [_H, C.c_char_p]),
    "sfm_load": (C.c_int32, [C.c_char_p, C.c_int32, C.POINTER(_H)]),
    "sfm_predict": (C.c_int32, [_H, _i64p, _i32p, _f32p, C.c_int64, _f32p]),
    "sfm_load_dataset": (C.c_int32, [_H, _i64p, _i32p, _f32p, _f32p, C.c_int64, C.c_int64]),
    "sfm_unload_dataset": (C.c_int32, [_H]),
    "sfm_synth_ctr_dataset": (C.c_int32, [_H, C.c_int64, C.c_int64, C.c_int32, _i32p, _u32p,
                                          _i64p, C.c_uint64]),
    "sfm_get_dataset_rows": (C.c_int32, [_H, C.c_int64, C.c_int64, _i64p, _i32p, _f32p, _f32p]),
    "sfm_dataset_info": (C.c_int32, [_H, _i64p, _i64p, _i32p]),
    "sfm_predict_resident": (C.c_int32, [_H, C.c_int64, C.c_int64, _f32p]),
    "sfm_evaluate": (C.c_int32, [_H, _f64p]),
    "sfm_evaluate_auc": (C.c_int32, [_H, _f64p]),
    "sfm_train_step": (C.c_int32, [_H, _i64p, C.c_int64, C.c_int64, _f64p, _i64p]),
    "sfm_train_step_csr": (C.c_int32, [_H, _i64p, _i32p, _f32p, _f32p, C.c_int64, C.c_int64,
                                       _f64p, _i64p]),
    "sfm_stage_csr": (C.c_int32, [_H, C.c_int32, _i64p, _i32p, _f32p, _f32p, C.c_int64]),
    "sfm_train_step_staged": (C.c_int32, [_H, C.c_int32, C.c_int64, _f64p, _i64p]),
    "sfm_stage_onehot": (C.c_int32, [_H, C.c_int32, _u32p, _u32p, _f32p, C.c_int64, C.c_int32,
                                     C.c_int32]),
    "sfm_pack_onehot": (C.c_int32, [_i32p, _f32p, C.c_int64, C.c_int32, C.c_int32, _u32p, _u32p]),
    "sfm_train": (C.c_int32, [_H, C.c_int64, C.c_int64, _f64p]),
    "sfm_sample_rows": (C.c_int32, [C.c_uint64, C.c_int64, C.c_double, C.c_int64, C.c_int64,
                                    _i64p, _i64p]),
    "sfm_partition_rows": (C.c_int32, [C.c_uint64, C.c_int64, C.c_int64, C.c_int64, C.c_int64,
                                       _i64p, _i64p]),
    "sfm_gradient": (C.c_int32, [_H, _i64p, C.c_int64, _f32p, _f32p, _f32p, _f64p, _i64p]),
    "sfm_comm_unique_id": (C.c_int32, [_u8p]),
    "sfm_comm_init": (C.c_int32, [_H, _u8p, C.c_int32, C.c_int32]),
    "sfm_comm_info": (C.c_int32, [_H, _i32p, _i32p]),
    "sfm_comm_mode": (C.c_int32, [_H, _i32p]),
    "sfm_als_sweep": (C.c_int32, [_H, C.c_int32, _f64p]),
    "sfm_als_residuals": (C.c_int32, [_H, _f64p, C.c_int64]),
    "sfm_comm_broadcast_model": (C.c_int32, [_H]),
    "sfm_parse_libfm": (C.c_int32, [C.c_char_p, C.c_uint64, C.c_int32, _i64p, _i64p, _i32p,
                                    _f64p, _i64p, _i32p, _f64p, _i64p]),
    "sfm_format_libfm": (C.c_int32, [_f64p, _i64p, _i32p, _f64p, C.c_int64, C.c_char_p,
                                     C.c_uint64, C.POINTER(C.c_uint64)]),
    "sfm_stats_get": (C.c_int32, [_H, C.POINTER(SfmStats)]),
    "sfm_stats_reset": (C.c_int32, [_H]),
    "sfm_set_phase_timing": (C.c_int32, [_H, C.c_int32]),
    "sfm_synchronize": (C.c_int32, [_H]),
    "sfm_timer_start": (C.c_int32, [_H]),
    "sfm_timer_stop": (C.c_int32, [_H, _f32p]),
}

_lib = None


def load():
    """Loads the shared object (raises OSError if it has not been built) and types every export."""
    global _lib
    if _lib is None:
        if not os.path.exists(SO_PATH):
            raise OSError(
                f"{SO_PATH} is missing: run `python -c 'import __graft_entry__ as g; g.build()'` "
                "or `make -C sparkfm_b200/csrc`; there is no CPU fallback")
        L = C.CDLL(SO_PATH)
        for name, (res, args) in SIGNATURES.items():
            fn = getattr(L, name)  # AttributeError if the export is missing
            fn.restype = res
            fn.argtypes = args
        _lib = L
    return _lib


class SfmError(RuntimeError):
    def __init__(self, status, message):
        super().__init__(f"[{status}] {message}")
        self.status = status


def check(status, handle=None):
    if status == SFM_OK:
        return
    L = load()
    msg = L.sfm_status_string(status).decode()
    if handle:
        detail = L.sfm_last_error(handle).decode()
        if detail:
            msg = f"{msg}: {detail}"
    raise SfmError(status, msg)
