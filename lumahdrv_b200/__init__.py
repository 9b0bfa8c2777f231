"""lumahdrv_b200 -- B200 (sm_100a) implementation of Luma HDRv's per-pixel
HDR<->integer transform (LumaQuantizer + the plane loops of LumaEncoder /
LumaDecoder), behind the reference's own interface.

    liblumacu.so   C ABI (include/lumacu.h): hand-written CUDA kernels + host glue
    luma.py        host mirror of LumaQuantizer / LumaEncoder / LumaDecoder (ctypes)
    device.py      device-resident batch API on torch tensors (streams, frame shards)
"""
from ._lib import LumaException, build_library, lib  # noqa: F401
from .luma import (CS_LUV, CS_RGB, CS_XYZ, CS_YCBCR, PTF_JND_HDRVDP, PTF_LINEAR, PTF_LOG, PTF_PQ, PTF_PSI,  # noqa: F401
                   Context, LumaDecoder, LumaDecoderParams, LumaEncoder, LumaEncoderParams, LumaQuantizer,
                   alloc_planes, build_lut, plane_dims, vpx_strides)

__version__ = "0.1.0"
