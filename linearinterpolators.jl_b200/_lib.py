"""ctypes binding of ``libli_b200.so`` (the C ABI declared in ``include/li_b200.h``).

The product has no CPU compute path: if the CUDA library is missing this module raises at
import time, and every compute entry point fails with ``LI_ERR_CUDA`` without a device.
"""
from __future__ import annotations

import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("LI_B200_LIB") or os.path.join(_HERE, "libli_b200.so")   # (override: A/B builds in development)

OK, ERR_ARG, ERR_DIM, ERR_INEXACT, ERR_BOUNDS, ERR_UNSUPPORTED, ERR_CUDA, ERR_NOMEM, ERR_NCCL = range(9)
COMM_ID_BYTES = 128
CG_MAXITER, CG_FTEST, CG_GTEST, CG_NOT_POSITIVE = range(4)
F32, F64 = 0, 1
FLAT, SAFEFLAT = 0, 1
VARIANT_AUTO, VARIANT_GATHER, VARIANT_ATOMIC = 0, 1, 2
FLAG_ASYNC = 1
MAX_S = 16
PRIME = 16          # LI_PRIME: family | PRIME = the derivative kernel


class KernelDesc(C.Structure):
    """``li_kernel_desc``."""
    _fields_ = [("family", C.c_int32), ("S", C.c_int32), ("boundary", C.c_int32), ("dtype", C.c_int32),
                ("params", C.c_double * 4)]


class TableInfo(C.Structure):
    """``li_table_info``."""
    _fields_ = [("dtype", C.c_int32), ("S", C.c_int32), ("nrows", C.c_int64), ("ncols", C.c_int64),
                ("windowed", C.c_int32), ("monotone", C.c_int32)]


class CgOptions(C.Structure):
    """``li_cg_options``."""
    _fields_ = [("maxiter", C.c_int32), ("check_every", C.c_int32), ("ftol", C.c_double), ("gtol_abs", C.c_double),
                ("gtol_rel", C.c_double)]


class CgInfo(C.Structure):
    """``li_cg_info``."""
    _fields_ = [("iters", C.c_int32), ("status", C.c_int32), ("rho0", C.c_double), ("rho", C.c_double)]


APPLY_FN = C.CFUNCTYPE(C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p)   # li_apply_fn


# ---- exceptions mirroring the reference's (SURVEY.md §5 "Argument validation") ----------------
class LIError(RuntimeError):
    """Julia ``ErrorException``."""


class ErrorException(LIError):
    pass


class DimensionMismatch(ValueError):
    pass


class ArgumentError(ValueError):
    pass


class InexactError(ArithmeticError):
    pass


class CudaError(LIError):
    pass


if not os.path.exists(LIB_PATH):
    raise ImportError(
        f"{LIB_PATH} is missing: build it with `python linearinterpolators.jl_b200/build.py` "
        "(nvcc, sm_100a).  There is no CPU fallback.")

lib = C.CDLL(LIB_PATH)

_vp, _i64, _dbl, _int, _sz = C.c_void_p, C.c_int64, C.c_double, C.c_int, C.c_size_t
_kd = C.POINTER(KernelDesc)
_PROTOS = {
    "li_version": (C.c_int, []),
    "li_last_error": (C.c_char_p, []),
    "li_status_string": (C.c_char_p, [_int]),
    "li_device_count": (_int, [C.POINTER(_int)]),
    "li_set_device": (_int, [_int]),
    "li_get_device": (_int, [C.POINTER(_int)]),
    "li_num_sms": (_int, [C.POINTER(_int)]),
    "li_malloc": (_int, [C.POINTER(_vp), _sz]),
    "li_free": (_int, [_vp]),
    "li_host_alloc": (_int, [C.POINTER(_vp), _sz]),
    "li_host_free": (_int, [_vp]),
    "li_memcpy_h2d": (_int, [_vp, _vp, _sz, _vp]),
    "li_memcpy_d2h": (_int, [_vp, _vp, _sz, _vp]),
    "li_memcpy_d2d": (_int, [_vp, _vp, _sz, _vp]),
    "li_memset_zero": (_int, [_vp, _sz, _vp]),
    "li_stream_create": (_int, [C.POINTER(_vp)]),
    "li_stream_destroy": (_int, [_vp]),
    "li_stream_sync": (_int, [_vp]),
    "li_vscale": (_int, [_int, _vp, _i64, _dbl, _vp]),
    "li_table_create": (_int, [_kd, _vp, _int, _i64, _i64, _vp, C.POINTER(_vp)]),
    "li_table_import": (_int, [_int, _int, _i64, _i64, _vp, _vp, _vp, C.POINTER(_vp)]),
    "li_table_export": (_int, [_vp, _vp, _vp, _vp]),
    "li_table_get_info": (_int, [_vp, C.POINTER(TableInfo)]),
    "li_table_destroy": (_int, [_vp]),
    "li_tab_apply": (_int, [_int, _vp, _i64, _i64, _dbl, _vp, _dbl, _vp, _int, _vp]),
    "li_sep_workspace_bytes": (_int, [_int, _int, C.POINTER(_vp), C.POINTER(_i64), _i64, C.POINTER(_sz)]),
    "li_sep_apply": (_int, [_int, _int, C.POINTER(_vp), C.POINTER(_i64), _i64, _dbl, _vp, _dbl, _vp, _vp, _sz,
                            _int, _vp]),
    "li_sep_apply_batch": (_int, [_int, _int, _i64, C.POINTER(_vp), C.POINTER(_i64), _i64, _dbl, C.POINTER(_vp), _dbl,
                                  C.POINTER(_vp), _int, _vp]),
    "li_otf1d_apply": (_int, [_int, _kd, _vp, _i64, _i64, _dbl, _vp, _dbl, _vp, _int, _vp]),
    "li_sep2d_otf_apply": (_int, [_int, _kd, _vp, _i64, _kd, _vp, _i64, _i64, _i64, _i64, _dbl, _vp, _dbl, _vp,
                                  _vp]),
    "li_affine2d_apply": (_int, [_int, _kd, _kd, C.POINTER(_dbl), _i64, _i64, _i64, _i64, _i64, _dbl, _vp, _dbl,
                                 _vp, _int, _int, _vp]),
    "li_tab_apply_host": (_int, [_int, _vp, _i64, _i64, _dbl, _vp, _dbl, _vp, _int, _vp]),
    "li_sep_apply_host": (_int, [_int, _int, C.POINTER(_vp), C.POINTER(_i64), _i64, _dbl, _vp, _dbl, _vp, _int,
                                 _vp]),
    "li_tab_apply_host_async": (_int, [_int, _vp, _i64, _i64, _dbl, _vp, _dbl, _vp, _int, _vp]),
    "li_sep_apply_host_async": (_int, [_int, _int, C.POINTER(_vp), C.POINTER(_i64), _i64, _dbl, _vp, _dbl, _vp, _int,
                                       _vp]),
    "li_kernel_launch_count": (_i64, []),
    "li_set_option": (_int, [C.c_char_p, _int]),
    "li_get_option": (_int, [C.c_char_p, C.POINTER(_int)]),
    "li_comm_unique_id": (_int, [_vp, _sz]),
    "li_comm_init_rank": (_int, [_vp, _int, _int, C.POINTER(_vp)]),
    "li_comm_init_all": (_int, [_int, C.POINTER(_int), C.POINTER(_vp)]),
    "li_comm_info": (_int, [_vp, C.POINTER(_int), C.POINTER(_int), C.POINTER(_int)]),
    "li_comm_destroy": (_int, [_vp]),
    "li_comm_group_start": (_int, []),
    "li_comm_group_end": (_int, []),
    "li_nccl_version": (_int, [C.POINTER(_int)]),
    "li_allreduce_sum": (_int, [_vp, _int, _vp, _i64, _vp]),
    "li_adjoint_accumulate_allreduce": (_int, [_vp, _int, _i64, _i64, _i64, C.POINTER(_vp), C.POINTER(_vp), _dbl, _dbl, _vp,
                                               _int, _vp]),
    "li_sep_otf_apply": (_int, [_int, _int, C.POINTER(_vp), C.POINTER(_vp), C.POINTER(_i64), C.POINTER(_i64), _i64, _dbl, _vp, _dbl,
                         _vp, _vp]),
    "li_table_atwa": (_int, [_vp, _vp, _vp, _vp]),
    "li_regularize": (_int, [_int, _vp, _i64, _dbl, _dbl, _vp]),
    "li_vdot": (_int, [_int, _vp, _vp, _i64, C.POINTER(_dbl), _vp]),
    "li_vdot_async": (_int, [_int, _vp, _vp, _i64, _vp, _vp]),
    "li_vupdate": (_int, [_int, _vp, _dbl, _vp, _i64, _vp]),
    "li_vcombine": (_int, [_int, _vp, _dbl, _vp, _dbl, _vp, _i64, _vp]),
    "li_conjgrad": (_int, [APPLY_FN, _vp, _int, _i64, _vp, _vp, _int, C.POINTER(CgOptions), C.POINTER(CgInfo), _vp]),
    "li_conjgrad_sep": (_int, [_int, C.POINTER(_vp), C.POINTER(_i64), _i64, _dbl, _vp, _vp, _int, C.POINTER(CgOptions),
                               C.POINTER(CgInfo), _vp]),
}
for _name, (_res, _args) in _PROTOS.items():
    _f = getattr(lib, _name)  # AttributeError here == the .so does not export the declared ABI
    _f.restype = _res
    _f.argtypes = _args

EXPORTED = tuple(_PROTOS)


def last_error() -> str:
    return lib.li_last_error().decode("utf-8", "replace")


def check(status: int) -> None:
    """Turn an ``li_status`` into the exception the reference would throw."""
    if status == OK:
        return
    msg = last_error()
    if status == ERR_DIM:
        raise DimensionMismatch(msg)
    if status == ERR_INEXACT:
        raise InexactError(msg)
    if status in (ERR_ARG, ERR_BOUNDS):
        raise ErrorException(msg)
    if status == ERR_UNSUPPORTED:
        raise NotImplementedError(msg)
    if status == ERR_NOMEM:
        raise MemoryError(msg)
    if status == ERR_NCCL:
        raise CudaError(f"LI_ERR_NCCL: {msg}")
    raise CudaError(f"{lib.li_status_string(status).decode()}: {msg}")


def set_option(name: str, value: int) -> None:
    """``li_set_option``: process-wide dispatch switch (e.g. ``set_option("no_fused", 1)`` for the cross-check path)."""
    check(lib.li_set_option(name.encode(), int(value)))


def get_option(name: str) -> int:
    v = C.c_int()
    check(lib.li_get_option(name.encode(), C.byref(v)))
    return int(v.value)


def launch_count() -> int:
    return int(lib.li_kernel_launch_count())
