"""ctypes binding of libscirs2_fft_cuda.so (include/scirs2_fft_cuda.h).

The library is the product; this module only declares its C ABI.  Loading fails
loudly when the shared object has not been built (``python -c 'import
__graft_entry__ as g; g.build()'`` or ``make``) — there is no Python fallback.
"""
from __future__ import annotations

import ctypes as C
import os

SFC_MAX_DIMS = 8

# sfc_status -> FFTError variant (scirs2-fft/src/error.rs:7-46)
SFC_OK = 0
SFC_ERR_COMPUTATION = -1
SFC_ERR_DIMENSION = -2
SFC_ERR_VALUE = -3
SFC_ERR_NOT_IMPLEMENTED = -4
SFC_ERR_IO = -5
SFC_ERR_BACKEND = -6
SFC_ERR_PLAN = -7
SFC_ERR_COMMUNICATION = -8
SFC_ERR_MEMORY = -9

SFC_F32, SFC_F64, SFC_C64, SFC_C128 = 0, 1, 2, 3
SFC_C2C, SFC_R2C, SFC_C2R = 0, 1, 2
SFC_PREC_F32, SFC_PREC_F64 = 0, 1
SFC_FORWARD, SFC_INVERSE = 0, 1
SFC_DESC_CUSTOM_IN_SHAPE = 1
SFC_DESC_REAL_INPUT = 2
SFC_DESC_AXIS_LEN = 4
SFC_DESC_AUX_MUL = 8
SFC_DESC_REAL_OUTPUT = 16
SFC_DESC_DCT2 = 32
SFC_DESC_DCT2_ORTHO0 = 64
SFC_DESC_DCT3 = 128
SFC_DESC_TRIG_SINE = 256
SFC_DESC_DCT4 = 512


class sfc_desc(C.Structure):
    _fields_ = [
        ("ndim", C.c_int32),
        ("shape", C.c_int64 * SFC_MAX_DIMS),
        ("naxes", C.c_int32),
        ("axes", C.c_int32 * SFC_MAX_DIMS),
        ("kind", C.c_int32),
        ("prec", C.c_int32),
        ("direction", C.c_int32),
        ("flags", C.c_int32),
        ("scale", C.c_double),
        ("in_shape", C.c_int64 * SFC_MAX_DIMS),
        ("scatter_parts", C.c_int32),
        ("reserved", C.c_int32),
        ("scatter_pitch", C.c_int64),
        ("axis_in_len", C.c_int64),
        ("axis_out_len", C.c_int64),
        ("aux_in", C.c_void_p),
        ("aux_out", C.c_void_p),
        ("scale_dc", C.c_double),
    ]


class sfc_plan_info(C.Structure):
    _fields_ = [
        ("in_bytes", C.c_int64),
        ("out_bytes", C.c_int64),
        ("scratch_bytes", C.c_int64),
        ("algorithmic_bytes", C.c_int64),
        ("device_bytes", C.c_int64),
        ("nominal_flops", C.c_double),
        ("num_launches", C.c_int32),
        ("num_passes", C.c_int32),
    ]


class sfc_cache_stats(C.Structure):
    _fields_ = [
        ("hit_count", C.c_uint64),
        ("miss_count", C.c_uint64),
        ("hit_rate", C.c_double),
        ("size", C.c_uint64),
        ("max_size", C.c_uint64),
    ]


SFC_MAX_GPUS = 16
SFC_DECOMP_REPLICATED, SFC_DECOMP_BATCH_SPLIT, SFC_DECOMP_SLAB = 0, 1, 2
SFC_SLAB_TRANSPOSED, SFC_SLAB_NATURAL = 0, 1


class sfc_dist_desc(C.Structure):
    _fields_ = [
        ("base", sfc_desc),
        ("decomposition", C.c_int32),
        ("layout", C.c_int32),
        ("chunks", C.c_int32),
        ("reserved", C.c_int32),
    ]


class sfc_dist_info(C.Structure):
    _fields_ = [
        ("world", C.c_int32),
        ("rank", C.c_int32),
        ("decomposition", C.c_int32),
        ("layout", C.c_int32),
        ("chunks", C.c_int32),
        ("reserved", C.c_int32),
        ("local_in_elems", C.c_int64),
        ("local_out_elems", C.c_int64),
        ("local_in_shape", C.c_int64 * SFC_MAX_DIMS),
        ("local_out_shape", C.c_int64 * SFC_MAX_DIMS),
        ("exchange_bytes_sent", C.c_int64),
        ("num_exchanges", C.c_int32),
        ("num_launches", C.c_int32),
        ("algorithmic_bytes", C.c_int64),
        ("nominal_flops", C.c_double),
    ]


LIB_NAME = "libscirs2_fft_cuda.so"
LIB_PATH = os.path.join(os.path.dirname(os.path.abspath(__file__)), "lib", LIB_NAME)

_vp, _i64, _i32, _int, _dbl = C.c_void_p, C.c_int64, C.c_int32, C.c_int, C.c_double
_pi64, _pi32, _pd = C.POINTER(C.c_int64), C.POINTER(C.c_int32), C.POINTER(C.c_double)
_str = C.c_char_p

# every symbol include/scirs2_fft_cuda.h declares: name -> (restype, argtypes)
SIGNATURES = {
    "sfc_init": (_int, [_int]),
    "sfc_device_count": (_int, []),
    "sfc_last_error": (_str, []),
    "sfc_abi_version": (_int, []),
    "sfc_is_available": (_int, []),
    "sfc_plan_create": (_int, [C.POINTER(_vp), C.POINTER(sfc_desc)]),
    "sfc_plan_destroy": (_int, [_vp]),
    "sfc_plan_get_info": (_int, [_vp, C.POINTER(sfc_plan_info)]),
    "sfc_plan_describe": (_int, [_vp, C.c_char_p, C.c_size_t]),
    "sfc_exec_device": (_int, [_vp, _vp, _vp, _vp]),
    "sfc_exec_host": (_int, [_vp, _vp, _vp]),
    "sfc_exec_device_scatter": (_int, [_vp, _vp, C.POINTER(_vp), _i32, _vp]),
    "sfc_dev_malloc": (_int, [C.POINTER(_vp), C.c_size_t]),
    "sfc_dev_free": (_int, [_vp]),
    "sfc_ipc_get_handle": (_int, [_vp, _vp]),
    "sfc_ipc_open_handle": (_int, [_vp, C.POINTER(_vp)]),
    "sfc_ipc_close_handle": (_int, [_vp]),
    "sfc_stream_synchronize": (_int, [_vp]),
    "sfc_comm_init_local": (_int, [C.POINTER(_vp), _i32, _pi32]),
    "sfc_comm_init_rank": (_int, [C.POINTER(_vp), _str, _i32, _i32, _i32]),
    "sfc_comm_destroy": (_int, [_vp]),
    "sfc_comm_size": (_int, [_vp]),
    "sfc_comm_rank": (_int, [_vp]),
    "sfc_comm_barrier": (_int, [_vp]),
    "sfc_comm_allgather": (_int, [_vp, _vp, _vp, C.c_size_t]),
    "sfc_comm_alloc": (_int, [_vp, C.c_size_t, C.POINTER(_vp)]),
    "sfc_comm_free": (_int, [_vp, _vp]),
    "sfc_dist_plan_create": (_int, [C.POINTER(_vp), _vp, C.POINTER(sfc_dist_desc)]),
    "sfc_dist_plan_destroy": (_int, [_vp]),
    "sfc_dist_plan_get_info": (_int, [_vp, C.POINTER(sfc_dist_info)]),
    "sfc_dist_plan_profile": (_int, [_vp, _i32]),
    "sfc_dist_plan_stage_ms": (_int, [_vp, _pd, _i32]),
    "sfc_dist_exec_device": (_int, [_vp, _vp, _vp, _vp]),
    "sfc_dist_exec_device_multi": (_int, [_vp, C.POINTER(_vp), C.POINTER(_vp), C.POINTER(_vp)]),
    "sfc_dist_synchronize": (_int, [_vp]),
    "sfc_dist_exec_host": (_int, [_vp, _vp, _vp]),
    "sfc_set_num_gpus": (_int, [_i32]),
    "sfc_get_num_gpus": (_int, []),
    "sfc_cache_get_stats": (_int, [C.POINTER(sfc_cache_stats)]),
    "sfc_cache_set_enabled": (_int, [_int]),
    "sfc_cache_is_enabled": (_int, []),
    "sfc_cache_clear": (_int, []),
    "sfc_cache_configure": (_int, [C.c_uint64, _dbl]),
    "sfc_planner_set_option": (_int, [_str, _str]),
    "sfc_planner_get_option": (_int, [_str, C.c_char_p, C.c_size_t]),
    "sfc_fft": (_int, [_vp, _i64, _int, _i64, _vp, _i64, _pi64]),
    "sfc_ifft": (_int, [_vp, _i64, _int, _i64, _vp, _i64, _pi64]),
    "sfc_rfft": (_int, [_vp, _i64, _int, _i64, _vp, _i64, _pi64]),
    "sfc_irfft": (_int, [_vp, _i64, _int, _i64, _vp, _i64, _pi64]),
    "sfc_fft2": (_int, [_vp, _i64, _i64, _int, _pi64, _pi32, _str, _vp, _i64, _pi64]),
    "sfc_ifft2": (_int, [_vp, _i64, _i64, _int, _pi64, _pi32, _str, _vp, _i64, _pi64]),
    "sfc_rfft2": (_int, [_vp, _i64, _i64, _int, _pi64, _vp, _i64, _pi64]),
    "sfc_irfft2": (_int, [_vp, _i64, _i64, _int, _pi64, _vp, _i64, _pi64]),
    "sfc_fftn": (_int, [_vp, _i32, _pi64, _int, _pi64, _pi64, _i32, _str, _vp, _i64, _pi64]),
    "sfc_ifftn": (_int, [_vp, _i32, _pi64, _int, _pi64, _pi64, _i32, _str, _vp, _i64, _pi64]),
    "sfc_rfftn": (_int, [_vp, _i32, _pi64, _int, _pi64, _pi64, _i32, _str, _vp, _i64, _pi64]),
    "sfc_irfftn": (_int, [_vp, _i32, _pi64, _int, _pi64, _i32, _pi64, _i32, _str, _vp, _i64, _pi64]),
    "sfc_fft_strided": (_int, [_vp, _i32, _pi64, _int, _i64, _int, _vp, _i64]),
    "sfc_backend_fft": (_int, [_vp, _i64, _vp, _i64]),
    "sfc_backend_ifft": (_int, [_vp, _i64, _vp, _i64]),
    "sfc_backend_fft_sized": (_int, [_vp, _i64, _vp, _i64, _i64]),
    "sfc_backend_ifft_sized": (_int, [_vp, _i64, _vp, _i64, _i64]),
    "sfc_backend_supports_feature": (_int, [_str]),
    "sfc_backend_name": (_str, []),
    "sfc_backend_description": (_str, []),
    "sfc_execute_batch": (_int, [_vp, _vp, _i64, _i64, _int]),
    "sfc_rfft_batch": (_int, [_vp, _i64, _i64, _int, _vp]),
    "sfc_irfft_batch": (_int, [_vp, _i64, _i64, _int, _vp]),
    "sfc_dct": (_int, [_vp, _i32, _vp, _vp, _i32, _i32, _i32, _str, _vp]),
    "sfc_dst": (_int, [_vp, _i32, _vp, _vp, _i32, _i32, _i32, _str, _vp]),
    "sfc_dht": (_int, [_vp, _i64, _vp]),
    "sfc_idht": (_int, [_vp, _i64, _vp]),
    "sfc_dht2": (_int, [_vp, _i64, _i64, _i32, _i32, _vp]),
    "sfc_hfft": (_int, [_vp, _i64, _int, _i64, _vp, _i64, _vp]),
    "sfc_ihfft": (_int, [_vp, _i64, _i64, _vp, _i64, _vp]),
    "sfc_hilbert": (_int, [_vp, _i64, _vp]),
    "sfc_fft_inplace": (_int, [_vp, _i64, _vp, _i64, _i32, _i32]),
    "sfc_fft2_efficient": (_int, [_vp, _i64, _i64, _int, _i64, _i64, _i32, _i32, _vp]),
    "sfc_fft_streaming": (_int, [_vp, _i64, _int, _i64, _i32, _i64, _vp]),
    "sfc_fftn_optimized": (_int, [_vp, _i32, _vp, _vp, _i32, _vp]),
    "sfc_czt": (_int, [_vp, _i64, _i64, _i64, _i32, C.c_double, C.c_double, C.c_double, C.c_double, _vp]),
    "sfc_signal_spectra": (_int, [_vp, _i64, _vp, _i64, _i64, _i64, _i64, _i32, _i32, _i64, C.c_double, _vp]),
    "sfc_stft": (_int, [_vp, _i64, _vp, _i64, _i64, _i64, _i32, _i32, _i32, _i32, C.c_double, _vp, _i64, _vp, _vp]),
}

_lib = None


def load() -> C.CDLL:
    """Load the shared library (once) and attach the prototypes."""
    global _lib
    if _lib is not None:
        return _lib
    path = os.environ.get("SFC_LIB_PATH", LIB_PATH)  # developer knob: try another build of the same library
    if not os.path.exists(path):
        raise ImportError(
            f"{path} is missing: build it with `make` (or __graft_entry__.build()); "
            "scirs_b200 has no CPU fallback"
        )
    lib = C.CDLL(path)
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(lib, name)  # AttributeError if the ABI lost a symbol
        fn.restype = res
        fn.argtypes = args
    _lib = lib
    return lib
