"""ctypes binding of the C ABI declared in ``include/lumacu.h``.

The shared library is built in-tree (``lumahdrv_b200/liblumacu.so``) by
``lumahdrv_b200/csrc/Makefile``.  There is no fallback: if the library is
missing, or a compute call is made without a CUDA device, this raises.
"""
from __future__ import annotations

import ctypes as C
import subprocess
from pathlib import Path

PKG_DIR = Path(__file__).resolve().parent
LIB_PATH = PKG_DIR / "liblumacu.so"
CSRC_DIR = PKG_DIR / "csrc"

OK = 0
STATUS_NAMES = {
    0: "LUMACU_OK",
    1: "LUMACU_ERR_INVALID_ARGUMENT",
    2: "LUMACU_ERR_CUDA",
    3: "LUMACU_ERR_NOT_CONFIGURED",
    4: "LUMACU_ERR_UNSUPPORTED",
    5: "LUMACU_ERR_OUT_OF_MEMORY",
    6: "LUMACU_ERR_NO_DEVICE",
}


class LumaException(RuntimeError):
    """Mirror of the reference's LumaException (include/luma/luma_exception.h:53-71)."""

    def __init__(self, msg: str, status: int = -1):
        super().__init__(msg)
        self.status = status


class FrameStats(C.Structure):
    _fields_ = [("sum", C.c_double), ("max", C.c_float), ("min", C.c_float)]


class DisplayParams(C.Structure):
    """lumacu_display_params (include/lumacu.h)."""
    _fields_ = [("exposure", C.c_float), ("gamma", C.c_float), ("user_scaling", C.c_float), ("do_tmo", C.c_int), ("ldr_sim", C.c_int),
                ("filter", C.c_int)]


class Metadata(C.Structure):
    """lumacu_metadata: the scalars of Matroska attachments 430..433, 435, 436 (include/lumacu.h)."""
    _fields_ = [("ptf_bit_depth", C.c_uint32), ("color_bit_depth", C.c_uint32), ("ptf", C.c_int32), ("color_space", C.c_int32),
                ("pre_scaling", C.c_float), ("max_lum", C.c_float), ("min_lum", C.c_float)]


def build_library(force: bool = False) -> Path:
    """Compile liblumacu.so for sm_100a (nvcc cross-compiles without a GPU)."""
    if force or not LIB_PATH.exists():
        subprocess.run(["make", "-j", "8", "-C", str(CSRC_DIR)] + (["-B"] if force else []), check=True, stdout=subprocess.DEVNULL)
    return LIB_PATH


_LIB = None

# name -> (restype, argtypes); must list every symbol include/lumacu.h declares
_P = C.c_void_p
_PP3 = C.POINTER(C.c_void_p)
_PI3 = C.POINTER(C.c_int32)
_PS3 = C.POINTER(C.c_size_t)
SIGNATURES = {
    "lumacu_version": (C.c_int, []),
    "lumacu_status_name": (C.c_char_p, [C.c_int]),
    "lumacu_device_count": (C.c_int, [C.POINTER(C.c_int)]),
    "lumacu_have_ptf_tables": (C.c_int, []),
    "lumacu_create": (C.c_int, [C.c_int, C.POINTER(_P)]),
    "lumacu_destroy": (C.c_int, [_P]),
    "lumacu_last_error": (C.c_char_p, [_P]),
    "lumacu_device": (C.c_int, [_P]),
    "lumacu_synchronize": (C.c_int, [_P]),
    "lumacu_stream": (_P, [_P]),
    "lumacu_host_alloc": (C.c_int, [C.c_size_t, C.POINTER(_P)]),
    "lumacu_host_free": (C.c_int, [_P]),
    "lumacu_metadata_pack": (C.c_int, [C.POINTER(Metadata), _P, C.c_uint32, _P, C.c_size_t, C.POINTER(C.c_size_t)]),
    "lumacu_metadata_unpack": (C.c_int, [_P, C.c_size_t, C.POINTER(Metadata), _P, C.c_size_t, C.POINTER(C.c_uint32)]),
    "lumacu_build_lut": (C.c_int, [C.c_int, C.c_uint, C.c_float, C.c_float, _P, C.c_size_t]),
    "lumacu_derive_thresholds": (C.c_int, [_P, C.c_uint32, _P]),
    "lumacu_plan_buckets": (C.c_int, [_P, C.c_uint32, C.POINTER(C.c_uint32), C.POINTER(C.c_uint32),
                                      C.POINTER(C.c_uint32), C.POINTER(C.c_uint32)]),
    "lumacu_set_quantizer": (C.c_int, [_P, _P, C.c_uint32, C.c_uint32, C.c_int, C.c_float]),
    "lumacu_broadcast_quantizer": (C.c_int, [C.POINTER(_P), C.c_int, C.c_int]),
    "lumacu_get_quantizer": (C.c_int, [_P, _P, C.c_size_t, C.POINTER(C.c_uint32), C.POINTER(C.c_uint32), C.POINTER(C.c_int),
                                       C.POINTER(C.c_float)]),
    "lumacu_encode_async": (C.c_int, [_P, _P, C.c_uint32, C.c_uint32, C.c_int, C.c_float, _PP3, _PI3, C.c_int,
                                      C.POINTER(FrameStats)]),
    "lumacu_decode_async": (C.c_int, [_P, _PP3, _PI3, C.c_uint32, C.c_uint32, C.c_int, C.c_float, _P]),
    "lumacu_encode_half_rgba": (C.c_int, [_P, _P, C.c_uint32, C.c_uint32, C.c_int, C.c_int, C.c_float, _PP3, _PI3,
                                          C.POINTER(FrameStats)]),
    "lumacu_decode_half_rgba": (C.c_int, [_P, _PP3, _PI3, C.c_uint32, C.c_uint32, C.c_int, C.c_float, _P]),
    "lumacu_wait_input": (C.c_int, [_P]),
    "lumacu_wait": (C.c_int, [_P]),
    "lumacu_pending": (C.c_int, [_P]),
    "lumacu_encode": (C.c_int, [_P, _P, C.c_uint32, C.c_uint32, C.c_int, C.c_float, _PP3, _PI3, C.c_int,
                                C.POINTER(FrameStats)]),
    "lumacu_decode": (C.c_int, [_P, _PP3, _PI3, C.c_uint32, C.c_uint32, C.c_int, C.c_float, _P]),
    "lumacu_display": (C.c_int, [_P, _PP3, _PI3, C.c_uint32, C.c_uint32, C.c_int, C.c_float, C.POINTER(DisplayParams), _P, C.c_int32]),
    "lumacu_display_dev": (C.c_int, [_P, _PP3, _PI3, C.c_uint32, C.c_uint32, C.c_int, C.c_float, C.POINTER(DisplayParams), _P,
                                     C.c_int32, C.c_uint32, _P, C.c_size_t, _P]),
    "lumacu_test_frame_dev": (C.c_int, [_P, _P, C.c_uint32, C.c_uint32, _P]),
    "lumacu_half_rgba_to_frame_dev": (C.c_int, [_P, _P, C.c_uint32, C.c_uint32, C.c_int, _P, _P]),
    "lumacu_frame_to_half_rgba_dev": (C.c_int, [_P, _P, C.c_uint32, C.c_uint32, _P, _P]),
    "lumacu_pfs_xyz_to_frame_dev": (C.c_int, [_P, _P, _P, _P, C.c_uint32, C.c_uint32, _P, _P]),
    "lumacu_frame_to_pfs_xyz_dev": (C.c_int, [_P, _P, C.c_uint32, C.c_uint32, _P, _P, _P, _P]),
    "lumacu_set_host_bands": (C.c_int, [_P, C.c_int]),
    "lumacu_host_register": (C.c_int, [_P, C.c_size_t]),
    "lumacu_host_unregister": (C.c_int, [_P]),
    "lumacu_set_tuning": (C.c_int, [_P, C.c_int, C.c_int, C.c_int]),
    "lumacu_quantize_planes": (C.c_int, [_P, _P, C.c_uint32, C.c_uint32, C.c_int, _PP3, _PI3, C.POINTER(FrameStats)]),
    "lumacu_dequantize_planes": (C.c_int, [_P, _PP3, _PI3, C.c_uint32, C.c_uint32, C.c_int, _P]),
    "lumacu_transform_color_space": (C.c_int, [_P, _P, C.c_uint32, C.c_uint32, C.c_int, C.c_float]),
    "lumacu_quantize": (C.c_int, [_P, _P, _P, C.c_size_t, C.c_uint]),
    "lumacu_dequantize": (C.c_int, [_P, _P, _P, C.c_size_t, C.c_uint]),
    "lumacu_encode_dev": (C.c_int, [_P, _P, _P, C.c_uint32, C.c_uint32, C.c_int, C.c_float, _PP3, _PI3, C.c_uint32,
                                    C.c_size_t, _PS3, _P, _P]),
    "lumacu_decode_dev": (C.c_int, [_P, _PP3, _PI3, C.c_uint32, C.c_uint32, C.c_int, C.c_float, _P, C.c_uint32,
                                    C.c_size_t, _PS3, _P]),
    "lumacu_transform_color_space_dev": (C.c_int, [_P, _P, C.c_uint32, C.c_uint32, C.c_int, C.c_float, _P]),
    "lumacu_quantize_dev": (C.c_int, [_P, _P, _P, C.c_size_t, C.c_uint, _P]),
    "lumacu_dequantize_dev": (C.c_int, [_P, _P, _P, C.c_size_t, C.c_uint, _P]),
    "lumacu_set_kernel_path": (C.c_int, [_P, C.c_int]),
    "lumacu_set_pq_tables": (C.c_int, [_P, C.c_int]),
    "lumacu_last_kernel_path": (C.c_int, [_P]),
    "lumacu_launch_count": (C.c_uint64, [_P]),
    "lumacu_search_info": (C.c_int, [_P, C.POINTER(C.c_int), C.POINTER(C.c_uint32), C.POINTER(C.c_uint32),
                                     C.POINTER(C.c_uint32)]),
}


def lib() -> C.CDLL:
    """The loaded C-ABI library; raises if it has not been built."""
    global _LIB
    if _LIB is None:
        if not LIB_PATH.exists():
            raise LumaException(
                f"{LIB_PATH} is missing: run `python -c 'import __graft_entry__ as g; g.build()'` "
                "(there is no CPU fallback for the transform)")
        handle = C.CDLL(str(LIB_PATH))
        for name, (res, args) in SIGNATURES.items():
            fn = getattr(handle, name)  # AttributeError if the ABI lost a symbol
            fn.restype = res
            fn.argtypes = args
        _LIB = handle
    return _LIB


def check(status: int, ctx=None, what: str = "") -> None:
    if status != OK:
        msg = lib().lumacu_last_error(ctx).decode(errors="replace")
        name = STATUS_NAMES.get(status, f"status {status}")
        raise LumaException(f"{what + ': ' if what else ''}{name}: {msg}", status)
