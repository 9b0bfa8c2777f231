"""Host-side mirror of the reference's interface for the per-pixel transform stage.

Same names, argument meaning and error behaviour as the reference classes, over
the C ABI in ``include/lumacu.h``:

* :class:`LumaQuantizer`  -- include/luma/luma_quantizer.h:89-126
* :class:`LumaEncoder`    -- include/luma/luma_encoder.h:110-176 (``encode`` minus ``run()``:
  the VP9/Matroska half stays with libvpx on the host and is out of scope)
* :class:`LumaDecoder`    -- include/luma/luma_decoder.h:112-175 (``decode`` minus ``run()``)

Frames are numpy ``float32`` arrays of shape ``[3, h, w]`` (LumaFrame layout,
include/luma/luma_frame.h:83-86); planes are pitched ``uint8`` arrays
``[rows, stride]`` like ``vpx_image_t`` planes.  All arithmetic runs on the GPU;
nothing here computes pixels on the CPU.
"""
from __future__ import annotations

import ctypes as C
from dataclasses import dataclass

import numpy as np

from . import _lib
from ._lib import FrameStats, LumaException, check

# enum order is wire format (include/luma/luma_quantizer.h:95-96)
PTF_PSI, PTF_PQ, PTF_LOG, PTF_JND_HDRVDP, PTF_LINEAR = range(5)
CS_LUV, CS_RGB, CS_YCBCR, CS_XYZ = range(4)
_PTF_NAMES = {"PSI": 0, "PQ": 1, "LOG": 2, "JND_HDRVDP": 3, "LINEAR": 4}
_CS_NAMES = {"LUV": 0, "RGB": 1, "YCBCR": 2, "XYZ": 3}


def _ptf(v) -> int:
    return _PTF_NAMES[v.upper()] if isinstance(v, str) else int(v)


def _cs(v) -> int:
    return _CS_NAMES[v.upper()] if isinstance(v, str) else int(v)


def plane_dims(w: int, h: int, profile: int):
    """Plane sizes of profiles 0..3 (src/luma_encoder.cpp:121-128,265-269)."""
    sub = profile in (0, 2)
    cw, ch = ((w + 1) >> 1, (h + 1) >> 1) if sub else (w, h)
    return [(w, h), (cw, ch), (cw, ch)]


def vpx_strides(w: int, profile: int, align: int = 32):
    """Byte pitches of ``vpx_img_alloc(..., align)`` as the reference encoder allocates them."""
    sub = profile in (0, 2)
    nbytes = 2 if profile > 1 else 1
    aw = (w + 1) & ~1 if sub else w
    s = ((aw + align - 1) & ~(align - 1)) * nbytes
    return [s, s >> 1 if sub else s, s >> 1 if sub else s]


def alloc_planes(w: int, h: int, profile: int, strides=None, fill: int = 0):
    strides = strides or vpx_strides(w, profile)
    return [np.full((ph, st), fill, dtype=np.uint8) for (pw, ph), st in zip(plane_dims(w, h, profile), strides)]


def _plane_args(planes):
    ptrs = (C.c_void_p * 3)(*[int(p.ctypes.data) for p in planes])
    strides = (C.c_int32 * 3)(*[int(p.strides[0]) for p in planes])
    return ptrs, strides


class Context:
    """Owns one ``lumacu_ctx`` (a CUDA device + stream + device-side quantizer state)."""

    def __init__(self, device: int = 0):
        self._lib = _lib.lib()
        self._h = C.c_void_p()
        check(self._lib.lumacu_create(int(device), C.byref(self._h)), None, "lumacu_create")
        self.device = int(device)

    def close(self):
        if getattr(self, "_h", None):
            self._lib.lumacu_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    @property
    def handle(self):
        if not self._h:
            raise LumaException("context is closed")
        return self._h

    def synchronize(self):
        check(self._lib.lumacu_synchronize(self.handle), self.handle, "lumacu_synchronize")

    @property
    def launch_count(self) -> int:
        return int(self._lib.lumacu_launch_count(self.handle))

    def set_kernel_path(self, path: int) -> None:
        """0 = tuned kernels when their preconditions hold (default), 1 = generic kernels only."""
        check(self._lib.lumacu_set_kernel_path(self.handle, int(path)), self.handle, "lumacu_set_kernel_path")

    def set_tuning(self, enc_variant: int = 0, dec_variant: int = 0, blocks_per_sm_cap: int = 0) -> None:
        """lumacu_set_tuning: pick an instantiation of the tuned kernels (0 = default; 1000 + v = bucket/threshold
        luma search instead of the direct table, 2000 + v = direct table read from global memory).  Every variant produces identical bits."""
        check(self._lib.lumacu_set_tuning(self.handle, int(enc_variant), int(dec_variant), int(blocks_per_sm_cap)),
              self.handle, "lumacu_set_tuning")

    def set_pq_tables(self, enable: bool) -> None:
        """CS_YCBCR: tuned kernels with (default) or without the exhaustive PQ tables; identical bits."""
        check(self._lib.lumacu_set_pq_tables(self.handle, int(bool(enable))), self.handle, "lumacu_set_pq_tables")

    def set_host_bands(self, bands: int) -> None:
        """Row bands per host-pointer call (0 = automatic)."""
        check(self._lib.lumacu_set_host_bands(self.handle, int(bands)), self.handle, "lumacu_set_host_bands")

    @property
    def last_kernel_path(self) -> int:
        """1 if the last encode/decode launch ran a tuned kernel, 0 if it ran a generic one."""
        return int(self._lib.lumacu_last_kernel_path(self.handle))


def build_lut(ptf, bitdepth: int, max_lum: float = 10000.0, min_lum: float = 0.005) -> np.ndarray:
    """The table half of LumaQuantizer::setQuantizer (host libm, reference formulas)."""
    n = (1 << int(bitdepth))
    out = np.empty(n, dtype=np.float32)
    check(_lib.lib().lumacu_build_lut(_ptf(ptf), int(bitdepth), float(max_lum), float(min_lum),
                                      out.ctypes.data, out.size), None, "lumacu_build_lut")
    return out


class LumaQuantizer:
    """LumaQuantizer (include/luma/luma_quantizer.h:89-126) backed by the CUDA layer."""

    PTF_PSI, PTF_PQ, PTF_LOG, PTF_JND_HDRVDP, PTF_LINEAR = range(5)
    CS_LUV, CS_RGB, CS_YCBCR, CS_XYZ = range(4)

    def __init__(self, device: int = 0, context: Context | None = None):
        self.ctx = context or Context(device)
        self._lib = self.ctx._lib
        # reference constructor defaults (src/luma_quantizer.cpp:45-52)
        self.m_Lmax, self.m_Lmin = 10000.0, 0.005
        self.m_colorSpace = CS_LUV
        self.m_mapping: np.ndarray | None = None
        self.m_maxVal = self.m_maxValColor = 0
        self.m_bitdepth = self.m_bitdepthColor = 0
        self._dirty = True

    @staticmethod
    def name(v, kind: str = "ptf") -> str:
        """LumaQuantizer::name (src/luma_quantizer.cpp:60-111)."""
        if kind == "ptf":
            return {PTF_PQ: "Perceptual quantizer (PQ, SMPTE ST 2084)", PTF_LOG: "Logarithmic",
                    PTF_JND_HDRVDP: "JND HDR-VDP", PTF_PSI: "Perceptual - Ferwerda's t.v.i.",
                    PTF_LINEAR: "Linear scaling"}.get(int(v), "Undefined")
        return {CS_LUV: "Lu'v'", CS_RGB: "RGB", CS_YCBCR: "YCbCr (ITU-R BT.2020)", CS_XYZ: "XYZ"}.get(
            int(v), "Undefined")

    def setQuantizer(self, ptf, bitdepth: int, cs, bitdepthC: int, maxLum: float = 10000.0, minLum: float = 0.005):
        """src/luma_quantizer.cpp:172-212."""
        self.m_ptf = _ptf(ptf)
        self.m_colorSpace = _cs(cs)
        self.m_bitdepth, self.m_bitdepthColor = int(bitdepth), int(bitdepthC)
        self.m_maxVal = (1 << self.m_bitdepth) - 1
        self.m_maxValColor = (1 << self.m_bitdepthColor) - 1
        self.m_Lmax, self.m_Lmin = float(maxLum), float(minLum)
        self.m_mapping = build_lut(self.m_ptf, self.m_bitdepth, self.m_Lmax, self.m_Lmin)
        self._dirty = True
        self._upload()
        return self

    # The reference hands out the internal LUT pointer and the decoder writes through
    # it (src/luma_decoder.cpp:122); the mirror re-uploads lazily after such a write.
    def getMapping(self) -> np.ndarray:
        self._dirty = True
        return self.m_mapping

    def setMapping(self, lut) -> None:
        lut = np.ascontiguousarray(lut, dtype=np.float32).reshape(-1)
        n = min(lut.size, self.m_mapping.size)
        self.m_mapping[:n] = lut[:n]
        self._dirty = True

    def getSize(self) -> int:
        return self.m_maxVal  # maxVal, not the entry count (include/luma/luma_quantizer.h:109)

    def getMaxLum(self) -> float:
        return self.m_Lmax

    def getMinLum(self) -> float:
        return self.m_Lmin

    def _upload(self):
        if self.m_mapping is None:
            raise LumaException("LumaQuantizer: setQuantizer has not been called", 3)
        if self._dirty:
            h = self.ctx.handle
            check(self._lib.lumacu_set_quantizer(h, self.m_mapping.ctypes.data, self.m_mapping.size,
                                                 self.m_maxValColor, self.m_colorSpace, self.m_Lmax), h,
                  "lumacu_set_quantizer")
            self._dirty = False

    def search_info(self) -> dict:
        self._upload()
        mode, nb, sh, wk = C.c_int(), C.c_uint32(), C.c_uint32(), C.c_uint32()
        h = self.ctx.handle
        check(self._lib.lumacu_search_info(h, C.byref(mode), C.byref(nb), C.byref(sh), C.byref(wk)), h)
        return {"mode": mode.value, "n_buckets": nb.value, "shift": sh.value, "walk": wk.value}

    def _elementwise(self, fn, val, ch: int):
        self._upload()
        scalar = np.isscalar(val)
        a = np.ascontiguousarray(np.atleast_1d(val), dtype=np.float32)
        out = np.empty_like(a)
        h = self.ctx.handle
        check(fn(h, a.ctypes.data, out.ctypes.data, a.size, int(ch)), h)
        return float(out[0]) if scalar else out

    def quantize(self, val, ch: int):
        """src/luma_quantizer.cpp:215-244 (codes are returned as floats, like the reference)."""
        return self._elementwise(self._lib.lumacu_quantize, val, ch)

    def dequantize(self, val, ch: int):
        """src/luma_quantizer.cpp:247-264."""
        return self._elementwise(self._lib.lumacu_dequantize, val, ch)

    def transformColorSpace(self, frame: np.ndarray, toCs: bool, sc: float = 1.0) -> bool:
        """src/luma_quantizer.cpp:267-482: in place; False on an unknown colour space."""
        if self.m_colorSpace not in (CS_LUV, CS_RGB, CS_YCBCR, CS_XYZ):
            return False
        _check_frame(frame)
        self._upload()
        _, hgt, wid = frame.shape
        h = self.ctx.handle
        check(self._lib.lumacu_transform_color_space(h, frame.ctypes.data, wid, hgt, int(bool(toCs)), float(sc)), h,
              "lumacu_transform_color_space")
        return True


def _check_frame(frame: np.ndarray):
    if not (isinstance(frame, np.ndarray) and frame.dtype == np.float32 and frame.ndim == 3 and frame.shape[0] == 3
            and frame.flags.c_contiguous):
        raise LumaException("frame must be a C-contiguous float32 array of shape [3, h, w]", 1)


@dataclass
class LumaEncoderParams:
    """LumaEncoderParamsBase + LumaEncoderParams defaults (include/luma/luma_encoder.h:59-70,110-118)."""
    quantizerScale: int = 2
    ptfBitDepth: int = 11
    colorBitDepth: int = 8
    preScaling: float = 1.0
    minLum: float = 0.005
    maxLum: float = 1e4
    fps: float = 25.0
    ptf: int = PTF_PQ
    colorSpace: int = CS_LUV
    bitrate: int = 10000
    profile: int = 2
    keyframeInterval: int = 0
    bitDepth: int = 12
    lossLess: bool = False


class LumaEncoder:
    """The per-pixel stage of LumaEncoder (include/luma/luma_encoder.h:121-176)."""

    def __init__(self, device: int = 0):
        self.m_params = LumaEncoderParams()
        self.m_quant = LumaQuantizer(device)
        self.m_initialized = False
        self.strict_side_effect = False  # write the colour-transformed frame back like the reference does
        self.last_stats: dict | None = None
        self.warnings: list[str] = []

    def getParams(self) -> LumaEncoderParams:
        return self.m_params

    def setParams(self, params: LumaEncoderParams):
        self.m_params = params

    def initialized(self) -> bool:
        return self.m_initialized

    def initialize(self, outputFile, w: int, h: int, verbose: bool = False) -> bool:
        """src/luma_encoder.cpp:63-129 without the Matroska/VP9 set-up."""
        p = self.m_params
        if p.profile > 1 and p.bitDepth == 8:
            p.profile -= 2
        if p.profile < 2 and p.bitDepth > 8:
            p.profile += 2
        p.ptf, p.colorSpace = _ptf(p.ptf), _cs(p.colorSpace)
        self.m_quant.setQuantizer(p.ptf, p.ptfBitDepth, p.colorSpace, p.colorBitDepth, p.maxLum, p.minLum)
        if w <= 0 or h <= 0 or w % 2 or h % 2:
            raise LumaException("Invalid frame size")
        self.width, self.height = int(w), int(h)
        self.m_rawFrame = alloc_planes(w, h, p.profile)  # vpx_img_alloc(..., 32) geometry
        self.m_initialized = True
        return True

    def setChannels(self, frame: np.ndarray, planes=None):
        """LumaEncoder::setChannels (src/luma_encoder.cpp:196-201): `frame` is ALREADY colour-transformed
        (LumaQuantizer.transformColorSpace(frame, True, sc)); [2x2 mean,] quantize and pack only."""
        if not self.m_initialized:
            raise LumaException("LumaEncoder: not initialized", 3)
        _check_frame(frame)
        _, h, w = frame.shape
        if (w, h) != (self.width, self.height):
            raise LumaException("Invalid frame size")
        planes = planes if planes is not None else self.m_rawFrame
        q = self.m_quant
        q._upload()
        ptrs, strides = _plane_args(planes)
        st = FrameStats()
        hnd = q.ctx.handle
        check(q._lib.lumacu_quantize_planes(hnd, frame.ctypes.data, w, h, self.m_params.profile, ptrs, strides, C.byref(st)),
              hnd, "lumacu_quantize_planes")
        mean = st.sum / (w * h)
        self.last_stats = {"sum": st.sum, "mean": mean, "max": st.max, "min": st.min}
        if mean <= 1.0:  # src/luma_encoder.cpp:314-316
            self.warnings.append("Warning! Mean luminance is %f cd/m2. Is the input calibrated to physical units?" % mean)
        return planes

    def encode(self, frame: np.ndarray, planes=None):
        """LumaEncoder::encode (include/luma/luma_encoder.h:142-148) minus run():
        returns the three integer planes the reference hands to libvpx."""
        if not self.m_initialized:
            raise LumaException("LumaEncoder: not initialized", 3)
        _check_frame(frame)
        _, h, w = frame.shape
        if (w, h) != (self.width, self.height):
            raise LumaException("Invalid frame size")
        planes = planes if planes is not None else self.m_rawFrame
        q = self.m_quant
        q._upload()
        ptrs, strides = _plane_args(planes)
        st = FrameStats()
        hnd = q.ctx.handle
        check(q._lib.lumacu_encode(hnd, frame.ctypes.data, w, h, self.m_params.profile, float(self.m_params.preScaling),
                                   ptrs, strides, int(self.strict_side_effect), C.byref(st)), hnd, "lumacu_encode")
        mean = st.sum / (w * h)
        self.last_stats = {"sum": st.sum, "mean": mean, "max": st.max, "min": st.min}
        if mean <= 1.0:  # src/luma_encoder.cpp:314-316
            msg = ("Warning! Mean luminance is %f cd/m2. Is the input calibrated to physical units?" % mean)
            self.warnings.append(msg)
        return planes

    def encode_half_rgba(self, rgba: np.ndarray, channels: int = 7, planes=None):
        """ExrInterface::readFrame's pixel loop + LumaEncoder::encode in one call: `rgba` is the [h, w, 4] float16 array
        of Imf::Rgba pixels as read from the file (8 B/px cross the bus instead of 12); same planes as expanding on
        the host and calling encode()."""
        if not self.m_initialized:
            raise LumaException("LumaEncoder: not initialized", 3)
        if not (isinstance(rgba, np.ndarray) and rgba.dtype == np.float16 and rgba.ndim == 3 and rgba.shape[2] == 4
                and rgba.flags.c_contiguous):
            raise LumaException("rgba must be a C-contiguous float16 [h, w, 4] array", 1)
        h, w, _ = rgba.shape
        if (w, h) != (self.width, self.height):
            raise LumaException("Invalid frame size")
        planes = planes if planes is not None else self.m_rawFrame
        q = self.m_quant
        q._upload()
        ptrs, strides = _plane_args(planes)
        st = FrameStats()
        hnd = q.ctx.handle
        check(q._lib.lumacu_encode_half_rgba(hnd, rgba.ctypes.data, w, h, int(channels), self.m_params.profile,
                                             float(self.m_params.preScaling), ptrs, strides, C.byref(st)), hnd,
              "lumacu_encode_half_rgba")
        mean = st.sum / (w * h)
        self.last_stats = {"sum": st.sum, "mean": mean, "max": st.max, "min": st.min}
        if mean <= 1.0:  # src/luma_encoder.cpp:314-316
            self.warnings.append("Warning! Mean luminance is %f cd/m2. Is the input calibrated to physical units?" % mean)
        return planes


@dataclass
class LumaDecoderParams:
    """include/luma/luma_decoder.h:60-70,112-120."""
    ptf: int = PTF_PSI
    colorSpace: int = CS_LUV
    preScaling: float = 1.0
    minLum: float = 0.005
    maxLum: float = 1e4
    ptfBitDepth: int = 11
    colorBitDepth: int = 8
    highBitDepth: bool = True
    profile: int = 2


class LumaDecoder:
    """The per-pixel stage of LumaDecoder (include/luma/luma_decoder.h:122-175)."""

    def __init__(self, device: int = 0):
        self.m_params = LumaDecoderParams()
        self.m_quant = LumaQuantizer(device)
        self.m_initialized = False
        self.m_frame: np.ndarray | None = None

    def getParams(self) -> LumaDecoderParams:
        return self.m_params

    def setParams(self, params: LumaDecoderParams):
        self.m_params = params

    def getQuantizer(self) -> LumaQuantizer:
        return self.m_quant

    def initialized(self) -> bool:
        return self.m_initialized

    def initialize(self, attachments: dict | None = None, verbose: bool = False) -> bool:
        """src/luma_decoder.cpp:63-166 with the Matroska attachments 430..436 passed as a dict
        ({430: ptfBitDepth, 431: colorBitDepth, 432: ptf, 433: colorSpace, 434: LUT floats,
        435: preScaling, 436: (maxLum, minLum)})."""
        p = self.m_params
        if attachments is not None:
            missing = [k for k in (430, 431, 432, 433, 434) if k not in attachments]
            if missing:
                raise LumaException("Failed to locate Luma HDRv meta data")
            p.ptfBitDepth, p.colorBitDepth = int(attachments[430]), int(attachments[431])
            p.ptf, p.colorSpace = _ptf(attachments[432]), _cs(attachments[433])
            if 435 in attachments:
                p.preScaling = float(attachments[435])
            if 436 in attachments:
                p.maxLum, p.minLum = (float(v) for v in attachments[436])
        self.m_quant.setQuantizer(p.ptf, p.ptfBitDepth, p.colorSpace, p.colorBitDepth, p.maxLum, p.minLum)
        if attachments is not None:
            self.m_quant.setMapping(attachments[434])  # memcpy over getMapping(), :122
        self.m_initialized = True
        return True

    def decode(self, planes, w: int, h: int, profile: int | None = None) -> np.ndarray:
        """LumaDecoder::decode (include/luma/luma_decoder.h:143-161) minus run(): returns the
        decoder-owned frame, overwritten by the next call."""
        if not self.m_initialized:
            raise LumaException("LumaDecoder: not initialized", 3)
        profile = self.m_params.profile if profile is None else int(profile)
        if self.m_frame is None or self.m_frame.shape != (3, h, w):
            self.m_frame = np.empty((3, h, w), dtype=np.float32)
        q = self.m_quant
        q._upload()
        ptrs, strides = _plane_args(planes)
        hnd = q.ctx.handle
        check(q._lib.lumacu_decode(hnd, ptrs, strides, w, h, profile, float(self.m_params.preScaling),
                                   self.m_frame.ctypes.data), hnd, "lumacu_decode")
        return self.m_frame

    def decode_half_rgba(self, planes, w: int, h: int, profile: int | None = None, out: np.ndarray | None = None) -> np.ndarray:
        """LumaDecoder::decode + ExrInterface::writeFrame's pixel loop in one call: returns the [h, w, 4] float16 array
        of Imf::Rgba pixels lumadec would hand to the EXR writer (alpha 0); 8 B/px cross the bus instead of 12."""
        if not self.m_initialized:
            raise LumaException("LumaDecoder: not initialized", 3)
        profile = self.m_params.profile if profile is None else int(profile)
        if out is None:
            out = np.empty((h, w, 4), dtype=np.float16)
        elif not (out.dtype == np.float16 and out.shape == (h, w, 4) and out.flags.c_contiguous):
            raise LumaException("out must be a C-contiguous float16 [h, w, 4] array", 1)
        q = self.m_quant
        q._upload()
        ptrs, strides = _plane_args(planes)
        hnd = q.ctx.handle
        check(q._lib.lumacu_decode_half_rgba(hnd, ptrs, strides, w, h, profile, float(self.m_params.preScaling),
                                             out.ctypes.data), hnd, "lumacu_decode_half_rgba")
        return out

    def getVpxChannels(self, planes, w: int, h: int, profile: int | None = None) -> np.ndarray:
        """LumaDecoder::getVpxChannels (src/luma_decoder.cpp:205-240): unpack, dequantize, [2x2 replicate]; the
        result still has to go through LumaQuantizer.transformColorSpace(frame, False, sc)."""
        if not self.m_initialized:
            raise LumaException("LumaDecoder: not initialized", 3)
        profile = self.m_params.profile if profile is None else int(profile)
        if self.m_frame is None or self.m_frame.shape != (3, h, w):
            self.m_frame = np.empty((3, h, w), dtype=np.float32)
        q = self.m_quant
        q._upload()
        ptrs, strides = _plane_args(planes)
        hnd = q.ctx.handle
        check(q._lib.lumacu_dequantize_planes(hnd, ptrs, strides, w, h, profile, self.m_frame.ctypes.data), hnd,
              "lumacu_dequantize_planes")
        return self.m_frame

    def display(self, planes, w: int, h: int, exposure: float = 1.0, gamma: float = 2.2, user_scaling: float = 1.0,
                do_tmo: bool = False, ldr_sim: bool = False, profile: int | None = None, linear: bool = False) -> np.ndarray:
        """The player's display path (src/lumaplay_dequantizer.frag:70-157): planes -> 8-bit RGBA [h, w, 4].
        linear=True samples like the player's GL_LINEAR textures (bilinear chroma, half-code LUT fetch)."""
        if not self.m_initialized:
            raise LumaException("LumaDecoder: not initialized", 3)
        profile = self.m_params.profile if profile is None else int(profile)
        q = self.m_quant
        q._upload()
        ptrs, strides = _plane_args(planes)
        out = np.empty((h, w, 4), dtype=np.uint8)
        p = _lib.DisplayParams(float(exposure), float(gamma), float(user_scaling), int(do_tmo), int(ldr_sim), int(linear))
        hnd = q.ctx.handle
        check(q._lib.lumacu_display(hnd, ptrs, strides, w, h, profile, float(self.m_params.preScaling), C.byref(p),
                                    out.ctypes.data, w * 4), hnd, "lumacu_display")
        return out
