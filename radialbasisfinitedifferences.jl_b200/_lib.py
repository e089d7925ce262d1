"""ctypes loader for librbffd.so (the C ABI declared in include/rbffd.h).

There is deliberately no fallback: if the CUDA library is missing or no device is present every
compute entry point raises RbffdError.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

_HERE = os.path.dirname(os.path.abspath(__file__))
# RBFFD_LIB selects a development build of the SAME library (kernel experiments, e.g. built with EXTRA=-DNSW_TIMING)
LIB_PATH = os.environ.get("RBFFD_LIB") or os.path.join(_HERE, "librbffd.so")
MAX_OPS = 12

OK, ERR_INVALID, ERR_K_TOO_LARGE, ERR_SINGULAR, ERR_CUDA, ERR_UNSUPPORTED, ERR_HALO = range(7)
OP_DERIV, OP_LAPLACE = 0, 1
ADVDIFF_COLLOCATED = 1


class RbffdError(RuntimeError):
    def __init__(self, code, msg):
        super().__init__(f"librbffd error {code}: {msg}")
        self.code = code


class Options(C.Structure):
    _fields_ = [("dim", C.c_int32), ("p", C.c_int32), ("polydeg", C.c_int32), ("n", C.c_int32), ("nops", C.c_int32),
                ("ops", (C.c_int32 * 4) * MAX_OPS), ("index_base", C.c_int32), ("sort_columns", C.c_int32),
                ("kernel", C.c_int32), ("variant", C.c_int32), ("index_width", C.c_int32), ("reserved", C.c_int32 * 3)]


class Halo(C.Structure):
    _fields_ = [("u", C.c_void_p), ("n_lo", C.c_int64), ("n_owned", C.c_int64), ("n_hi", C.c_int64), ("flags", C.c_void_p),
                ("peer_lo_u", C.c_void_p), ("peer_lo_flags", C.c_void_p), ("peer_lo_offset", C.c_int64), ("count_to_lo", C.c_int64),
                ("peer_hi_u", C.c_void_p), ("peer_hi_flags", C.c_void_p), ("peer_hi_offset", C.c_int64), ("count_to_hi", C.c_int64)]


class AdvDiffParams(C.Structure):
    _fields_ = [("iE", C.c_int32), ("iDx", C.c_int32), ("iDy", C.c_int32), ("iDxx", C.c_int32), ("iDyy", C.c_int32),
                ("iDxk", C.c_int32), ("iDyk", C.c_int32), ("flags", C.c_int32),
                ("alpha", C.c_double), ("ux", C.c_double), ("uy", C.c_double), ("gamma", C.c_double)]


def build(force: bool = False) -> str:
    """Compile librbffd.so in-tree for sm_100a (nvcc cross-compiles without a GPU)."""
    csrc = os.path.join(_HERE, "csrc")
    if force:
        subprocess.check_call(["make", "-s", "-C", csrc, "clean"])
    subprocess.check_call(["make", "-s", "-j8", "-C", csrc])
    return LIB_PATH


_lib = None
_vp = C.c_void_p
_i64 = C.c_int64
_i32 = C.c_int32
_dbl = C.c_double

_SIGNATURES = {
    "rbffd_version": ([], C.c_int),
    "rbffd_create": ([C.c_int, C.POINTER(_vp)], C.c_int),
    "rbffd_destroy": ([_vp], C.c_int),
    "rbffd_last_error": ([_vp], C.c_char_p),
    "rbffd_set_stream": ([_vp, _vp], C.c_int),
    "rbffd_get_stream": ([_vp, C.POINTER(_vp)], C.c_int),
    "rbffd_reset_stream": ([_vp], C.c_int),
    "rbffd_synchronize": ([_vp], C.c_int),
    "rbffd_timings": ([_vp, C.POINTER(_dbl), C.c_int], C.c_int),
    "rbffd_knn_device": ([_vp, _vp, _i64, _i32, _vp, _i64, _i32, _vp, _vp, _vp, _vp], C.c_int),
    "rbffd_calculateneighbors_host": ([_vp, _vp, _i64, _vp, _i64, _i32, _i32, _vp, _i32, _vp, _vp, _vp, _vp], C.c_int),
    "rbffd_generate_operator_host": ([_vp, C.POINTER(Options), _vp, _i64, _vp, _i64, _vp, _vp, _vp], C.c_int),
    "rbffd_weights_device": ([_vp, C.POINTER(Options), _vp, _i64, _vp, _i64, _vp, _i64, _vp, _vp, _vp], C.c_int),
    "rbffd_launch_count": ([_vp], C.c_longlong),
    "rbffd_measure_fp64_peak": ([_vp, C.POINTER(_dbl), C.POINTER(_dbl)], C.c_int),
    "rbffd_operator_generate": ([_vp, C.POINTER(Options), _vp, _i64, _vp, _i64, _vp, C.POINTER(_vp)], C.c_int),
    "rbffd_operator_generate_host": ([_vp, C.POINTER(Options), _vp, _i64, _vp, _i64, _vp, C.POINTER(_vp)], C.c_int),
    "rbffd_device_malloc": ([_vp, _i64, C.POINTER(_vp)], C.c_int),
    "rbffd_device_free": ([_vp, _vp], C.c_int),
    "rbffd_device_upload": ([_vp, _vp, _vp, _i64], C.c_int),
    "rbffd_device_download": ([_vp, _vp, _vp, _i64], C.c_int),
    "rbffd_operator_from_host": ([_vp, _i64, _i64, _i32, _i32, _vp, _i32, _vp, C.POINTER(_vp)], C.c_int),
    "rbffd_operator_from_device": ([_vp, _i64, _i64, _i32, _i32, _vp, _vp, C.POINTER(_vp)], C.c_int),
    "rbffd_stencils_device": ([_vp, _vp, _i64, _i32, _vp, _i64, _i32, _vp, _vp, _vp], C.c_int),
    "rbffd_operator_destroy": ([_vp], C.c_int),
    "rbffd_operator_info": ([_vp, C.POINTER(_i64), C.POINTER(_i64), C.POINTER(_i32), C.POINTER(_i32)], C.c_int),
    "rbffd_operator_pointers": ([_vp, _i32, C.POINTER(_vp), C.POINTER(_vp)], C.c_int),
    "rbffd_operator_to_host": ([_vp, _i32, _vp, _vp], C.c_int),
    "rbffd_spmv_device": ([_vp, _i32, _dbl, _vp, _dbl, _vp], C.c_int),
    "rbffd_spmv_t_device": ([_vp, _i32, _dbl, _vp, _dbl, _vp], C.c_int),
    "rbffd_spmv_multi_device": ([_vp, _i32, C.POINTER(_i32), C.POINTER(_dbl), _vp, _vp], C.c_int),
    "rbffd_operator_combine_device": ([_vp, _i32, C.POINTER(_i32), C.POINTER(_dbl), _vp], C.c_int),
    "rbffd_spmv_host": ([_vp, _i32, _dbl, _vp, _dbl, _vp], C.c_int),
    "rbffd_spmv_t_host": ([_vp, _i32, _dbl, _vp, _dbl, _vp], C.c_int),
    "rbffd_rhs_advdiff_device": ([_vp, C.POINTER(AdvDiffParams), _vp, _vp], C.c_int),
    "rbffd_rhs_advdiff_stage_device": ([_vp, C.POINTER(AdvDiffParams), _vp, _dbl, _vp, _dbl, _dbl, _vp], C.c_int),
    "rbffd_stage_update_device": ([_vp, _i64, _dbl, _vp, _dbl, _vp, _dbl, _vp, _vp], C.c_int),
    "rbffd_spmv_stage_device": ([_vp, _i32, C.POINTER(_i32), C.POINTER(_dbl), _vp, _dbl, _vp, _dbl, _dbl, _vp], C.c_int),
    "rbffd_shard_spmv_stage_device": ([_vp, _vp, _i32, C.POINTER(_i32), C.POINTER(_dbl), _vp, _dbl, _vp, _dbl, _dbl, _vp], C.c_int),
    "rbffd_rhs_advdiff_host": ([_vp, C.POINTER(AdvDiffParams), _vp, _vp], C.c_int),
    "rbffd_bc_create": ([_vp, _i32, _vp, _vp, _vp, _vp, _vp, _vp, _i32, C.POINTER(_vp)], C.c_int),
    "rbffd_bc_apply_device": ([_vp, _vp], C.c_int),
    "rbffd_bc_apply_host": ([_vp, _vp], C.c_int),
    "rbffd_bc_destroy": ([_vp], C.c_int),
    "rbffd_ipc_alloc": ([_vp, _i64, C.POINTER(_vp), C.c_char_p], C.c_int),
    "rbffd_ipc_free": ([_vp, _vp], C.c_int),
    "rbffd_ipc_open": ([_vp, C.c_char_p, C.POINTER(_vp)], C.c_int),
    "rbffd_ipc_close": ([_vp, _vp], C.c_int),
    "rbffd_halo_push_device": ([_vp, C.POINTER(Halo), C.c_uint32], C.c_int),
    "rbffd_halo_wait_device": ([_vp, C.POINTER(Halo), C.c_uint32], C.c_int),
    "rbffd_halo_ack_device": ([_vp, C.POINTER(Halo), C.c_uint32], C.c_int),
    "rbffd_gather_device": ([_vp, _vp, _vp, _i64, _vp], C.c_int),
    "rbffd_scatter_add_device": ([_vp, _vp, _vp, _i64, _vp], C.c_int),
    "rbffd_jittered_lattice_device": ([_vp, _i32, _i64, C.c_uint64, _i64, _i64, _vp], C.c_int),
    "rbffd_jittered_lattice_box_device": ([_vp, _i32, _i64, C.c_uint64, _vp, _vp, _vp, _vp, _vp, _vp], C.c_int),
    "rbffd_shard_plan_host": ([_vp, _i64, _i32, _i32, _vp, _vp], C.c_int),
    "rbffd_shard_create_host": ([_vp, _vp, _i64, _i32, _vp, _i32, _i32, _i32, C.POINTER(_vp)], C.c_int),
    "rbffd_shard_create_device": ([_vp, _i32, _i32, _vp, _vp, _vp, _i64, _vp, _vp, _i32, _i32, C.POINTER(_vp)], C.c_int),
    "rbffd_shard_destroy": ([_vp], C.c_int),
    "rbffd_shard_status": ([_vp], C.c_int),
    "rbffd_shard_info": ([_vp, C.POINTER(_i64), C.POINTER(_i64), C.POINTER(_i64)], C.c_int),
    "rbffd_shard_global_ids_host": ([_vp, _i32, _vp], C.c_int),
    "rbffd_shard_device_arrays": ([_vp, C.POINTER(_vp), C.POINTER(_vp)], C.c_int),
    "rbffd_shard_recv_count": ([_vp, _i32, C.POINTER(_i64)], C.c_int),
    "rbffd_shard_send_count": ([_vp, _i32, C.POINTER(_i64)], C.c_int),
    "rbffd_shard_recv_ids_host": ([_vp, _i32, _i32, _vp], C.c_int),
    "rbffd_shard_set_send_ids_host": ([_vp, _i32, _i32, _vp, _i64], C.c_int),
    "rbffd_shard_finalize": ([_vp, _vp, _vp], C.c_int),
    "rbffd_shard_connect": ([_vp, _i32, _vp, _i64, _i64, _i64], C.c_int),
    "rbffd_shard_spmv_device": ([_vp, _vp, _i32, C.POINTER(_i32), C.POINTER(_dbl), _vp, _vp], C.c_int),
    "rbffd_shard_spmv_local_device": ([_vp, _vp, _i32, C.POINTER(_i32), C.POINTER(_dbl), _vp, _vp], C.c_int),
    "rbffd_shard_spmv_t_device": ([_vp, _vp, _i32, _dbl, _vp, _dbl, _vp], C.c_int),
    "rbffd_shard_spmv_t_local_device": ([_vp, _vp, _i32, _dbl, _vp, _dbl, _vp], C.c_int),
    "rbffd_shard_pack_device": ([_vp, _i32, _vp, _vp], C.c_int),
    "rbffd_shard_unpack_device": ([_vp, _i32, _vp], C.c_int),
    "rbffd_shard_tpack_device": ([_vp, _i32, _vp], C.c_int),
    "rbffd_shard_tunpack_add_device": ([_vp, _i32, _vp, _vp], C.c_int),
}


def lib():
    """Load librbffd.so; raises RbffdError (never falls back) if it is missing."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise RbffdError(ERR_CUDA, f"{LIB_PATH} is missing: run `python -c 'import __graft_entry__ as g; g.build()'`"
                                       " (there is no CPU fallback)")
        L = C.CDLL(LIB_PATH)
        for name, (args, res) in _SIGNATURES.items():
            fn = getattr(L, name)
            fn.argtypes = args
            fn.restype = res
        _lib = L
    return _lib


def exported_symbols():
    return sorted(_SIGNATURES)
