"""ctypes bindings for the TEST-ONLY checkers under oracle/.

* :class:`Oracle`  -- the plain-C restatement (``libluma_oracle.so``).
* :class:`Reference` -- the unmodified reference sources compiled in place
  (``oracle/_ref/libluma_ref.so``), driven through their own public classes.

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s CPU-baseline
legs may import this module.  The product package never does.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess
from pathlib import Path

import numpy as np

HERE = Path(__file__).resolve().parent
ORACLE_SO = HERE / "libluma_oracle.so"
REF_SO = HERE / "_ref" / "libluma_ref.so"
REF_O0_SO = HERE / "_ref" / "libluma_ref_O0.so"
REFERENCE_DIR = Path(os.environ.get("LUMA_REFERENCE_DIR", "/root/reference"))

PTF = {"PSI": 0, "PQ": 1, "LOG": 2, "JND_HDRVDP": 3, "LINEAR": 4}
CS = {"LUV": 0, "RGB": 1, "YCBCR": 2, "XYZ": 3}

FNV_OFFSET = 2166136261


def build(force: bool = False) -> None:
    """Build the checkers (gcc only).  _ref is built only where the reference
    checkout exists (this container); on the GPU box the prebuilt files travel."""
    if force or not ORACLE_SO.exists() or (REFERENCE_DIR.exists() and not REF_SO.exists()):
        subprocess.run(["make", "-C", str(HERE), f"REF={REFERENCE_DIR}"], check=True,
                       stdout=subprocess.DEVNULL)


def fnv1a32(data: bytes | np.ndarray, seed: int = FNV_OFFSET) -> int:
    lib = _oracle_lib()
    buf = np.ascontiguousarray(np.frombuffer(data, dtype=np.uint8) if isinstance(data, (bytes, bytearray))
                               else data).view(np.uint8).reshape(-1)
    return int(lib.lo_fnv1a32(buf.ctypes.data_as(C.c_void_p), C.c_size_t(buf.size), C.c_uint32(seed)))


def plane_dims(w: int, h: int, profile: int):
    sub = profile in (0, 2)
    cw, ch = ((w + 1) >> 1, (h + 1) >> 1) if sub else (w, h)
    return [(w, h), (cw, ch), (cw, ch)]


def vpx_strides(w: int, profile: int, align: int = 32):
    """Pitches of vpx_img_alloc(..., align) as the reference encoder uses them."""
    sub = profile in (0, 2)
    nbytes = 2 if profile > 1 else 1
    aw = (w + 1) & ~1 if sub else w
    s = (aw + align - 1) & ~(align - 1)
    y = s * nbytes
    return [y, y >> 1 if sub else y, y >> 1 if sub else y]


def alloc_planes(w: int, h: int, profile: int, strides=None, fill: int = 0):
    """Three pitched uint8 planes (numpy, shape [rows, stride])."""
    strides = strides or vpx_strides(w, profile)
    return [np.full((ph, st), fill, dtype=np.uint8) for (pw, ph), st in zip(plane_dims(w, h, profile), strides)]


def plane_payload(planes, w: int, h: int, profile: int):
    """Strip pitch padding -> list of [rows, row_bytes] uint8 arrays."""
    nbytes = 2 if profile > 1 else 1
    return [np.ascontiguousarray(p[:ph, : pw * nbytes]) for p, (pw, ph) in zip(planes, plane_dims(w, h, profile))]


def plane_codes(planes, w: int, h: int, profile: int):
    """Integer codes per plane as uint16 arrays [rows, cols]."""
    out = []
    for p in plane_payload(planes, w, h, profile):
        out.append(p.view("<u2").copy() if profile > 1 else p.astype(np.uint16))
    return out


def plane_hashes(planes, w: int, h: int, profile: int):
    return [fnv1a32(p) for p in plane_payload(planes, w, h, profile)]


class _QStruct(C.Structure):
    _fields_ = [("ptf", C.c_int), ("color_space", C.c_int), ("bitdepth", C.c_uint), ("bitdepth_color", C.c_uint),
                ("max_val", C.c_uint), ("max_val_color", C.c_uint), ("l_max", C.c_float), ("l_min", C.c_float),
                ("mapping", C.POINTER(C.c_float))]


_ORACLE = None
_REFS: dict = {}


def _oracle_lib():
    global _ORACLE
    if _ORACLE is None:
        if not ORACLE_SO.exists():
            build()
        lib = C.CDLL(str(ORACLE_SO))
        P = C.POINTER(_QStruct)
        lib.lo_init.argtypes = [P]
        lib.lo_free.argtypes = [P]
        lib.lo_set_quantizer.argtypes = [P, C.c_int, C.c_uint, C.c_int, C.c_uint, C.c_float, C.c_float]
        lib.lo_set_quantizer.restype = C.c_int
        lib.lo_quantize.argtypes = [P, C.c_float, C.c_uint]
        lib.lo_quantize.restype = C.c_float
        lib.lo_dequantize.argtypes = [P, C.c_float, C.c_uint]
        lib.lo_dequantize.restype = C.c_float
        lib.lo_transform_pq.argtypes = [P, C.c_float, C.c_int]
        lib.lo_transform_pq.restype = C.c_float
        lib.lo_transform_color_space.argtypes = [P, C.c_void_p, C.c_uint, C.c_uint, C.c_int, C.c_float]
        lib.lo_transform_color_space.restype = C.c_int
        PP = C.POINTER(C.c_void_p)
        PI = C.POINTER(C.c_int)
        lib.lo_pack_planes.argtypes = [P, C.c_void_p, C.c_uint, C.c_uint, C.c_int, PP, PI, C.POINTER(C.c_float)]
        lib.lo_unpack_planes.argtypes = [P, PP, PI, C.c_uint, C.c_uint, C.c_int, C.c_void_p]
        lib.lo_encode.argtypes = [P, C.c_void_p, C.c_uint, C.c_uint, C.c_int, C.c_float, PP, PI, C.POINTER(C.c_float)]
        lib.lo_decode.argtypes = [P, PP, PI, C.c_uint, C.c_uint, C.c_int, C.c_float, C.c_void_p]
        lib.lo_test_frame.argtypes = [C.c_void_p, C.c_uint, C.c_uint]
        lib.lo_fnv1a32.argtypes = [C.c_void_p, C.c_size_t, C.c_uint32]
        lib.lo_fnv1a32.restype = C.c_uint32
        lib.lo_have_ptf_tables.restype = C.c_int
        _ORACLE = lib
    return _ORACLE


def _plane_args(planes):
    ptrs = (C.c_void_p * 3)(*[p.ctypes.data for p in planes])
    strides = (C.c_int * 3)(*[int(p.strides[0]) for p in planes])
    return ptrs, strides


def test_frame(w: int, h: int) -> np.ndarray:
    """ExrInterface::testFrame restated (src/exr_interface.cpp:50-70) -> [3,h,w] f32."""
    out = np.empty((3, h, w), dtype=np.float32)
    _oracle_lib().lo_test_frame(out.ctypes.data_as(C.c_void_p), w, h)
    return out


def noise_frame(w: int, h: int, seed: int = 0x9E3779B97F4A7C15, lo: float = 0.005, span: float = 2.0e6) -> np.ndarray:
    """Seeded log-uniform HDR noise, v = lo * span**u (SURVEY 8d input (2))."""
    rng = np.random.Generator(np.random.PCG64(seed & 0xFFFFFFFFFFFFFFFF))
    u = rng.random((3, h, w), dtype=np.float64)
    return (lo * np.power(span, u)).astype(np.float32)


class Oracle:
    """Plain-C restatement; method names mirror LumaQuantizer / LumaEncoder / LumaDecoder."""

    def __init__(self):
        self.lib = _oracle_lib()
        self.q = _QStruct()
        self.lib.lo_init(C.byref(self.q))

    def __del__(self):
        try:
            self.lib.lo_free(C.byref(self.q))
        except Exception:
            pass

    def setQuantizer(self, ptf, bitdepth, cs, bitdepthC, maxLum=10000.0, minLum=0.005):
        ptf = PTF[ptf] if isinstance(ptf, str) else int(ptf)
        cs = CS[cs] if isinstance(cs, str) else int(cs)
        rc = self.lib.lo_set_quantizer(C.byref(self.q), ptf, bitdepth, cs, bitdepthC, maxLum, minLum)
        if rc:
            raise RuntimeError(f"lo_set_quantizer failed ({rc})")
        return self

    def getSize(self):
        return int(self.q.max_val)

    def getMapping(self) -> np.ndarray:
        return np.ctypeslib.as_array(self.q.mapping, shape=(self.q.max_val + 1,)).copy()

    def setMapping(self, lut: np.ndarray):
        """Mirror of the decoder's memcpy into getMapping() (src/luma_decoder.cpp:122)."""
        lut = np.ascontiguousarray(lut, dtype=np.float32)
        n = min(lut.size, self.q.max_val + 1)
        C.memmove(self.q.mapping, lut.ctypes.data, n * 4)

    def quantize(self, val, ch):
        return float(self.lib.lo_quantize(C.byref(self.q), float(np.float32(val)), ch))

    def dequantize(self, val, ch):
        return float(self.lib.lo_dequantize(C.byref(self.q), float(np.float32(val)), ch))

    def transformPQ(self, val, encode):
        return float(self.lib.lo_transform_pq(C.byref(self.q), float(np.float32(val)), int(encode)))

    def transformColorSpace(self, frame: np.ndarray, toCs: bool, sc: float = 1.0) -> bool:
        assert frame.dtype == np.float32 and frame.flags.c_contiguous and frame.shape[0] == 3
        _, h, w = frame.shape
        return bool(self.lib.lo_transform_color_space(C.byref(self.q), frame.ctypes.data_as(C.c_void_p), w, h,
                                                      int(bool(toCs)), sc))

    def encode(self, frame: np.ndarray, profile: int = 2, preScaling: float = 1.0, strides=None):
        """LumaEncoder::encode minus run(): returns (planes, avg[3]); mutates `frame` like the reference."""
        assert frame.dtype == np.float32 and frame.flags.c_contiguous and frame.shape[0] == 3
        _, h, w = frame.shape
        planes = alloc_planes(w, h, profile, strides)
        ptrs, st = _plane_args(planes)
        avg = (C.c_float * 3)()
        self.lib.lo_encode(C.byref(self.q), frame.ctypes.data_as(C.c_void_p), w, h, profile, preScaling, ptrs, st, avg)
        return planes, [float(a) for a in avg]

    def decode(self, planes, w: int, h: int, profile: int = 2, preScaling: float = 1.0) -> np.ndarray:
        out = np.empty((3, h, w), dtype=np.float32)
        ptrs, st = _plane_args(planes)
        self.lib.lo_decode(C.byref(self.q), ptrs, st, w, h, profile, preScaling, out.ctypes.data_as(C.c_void_p))
        return out

    def unpack(self, planes, w: int, h: int, profile: int = 2) -> np.ndarray:
        out = np.empty((3, h, w), dtype=np.float32)
        ptrs, st = _plane_args(planes)
        self.lib.lo_unpack_planes(C.byref(self.q), ptrs, st, w, h, profile, out.ctypes.data_as(C.c_void_p))
        return out


class _RefParams(C.Structure):
    _fields_ = [("ptf", C.c_int), ("color_space", C.c_int), ("ptf_bits", C.c_uint), ("color_bits", C.c_uint),
                ("profile", C.c_uint), ("bit_depth", C.c_uint), ("pre_scaling", C.c_float), ("max_lum", C.c_float),
                ("min_lum", C.c_float)]


def reference_available(o0: bool = False) -> bool:
    return (REF_O0_SO if o0 else REF_SO).exists()


def _ref_lib(o0: bool = False):
    key = bool(o0)
    if key not in _REFS:
        so = REF_O0_SO if o0 else REF_SO
        if not so.exists():
            build()
        lib = C.CDLL(str(so))
        lib.lref_quant_new.restype = C.c_void_p
        lib.lref_quant_free.argtypes = [C.c_void_p]
        lib.lref_quant_set.argtypes = [C.c_void_p, C.c_int, C.c_uint, C.c_int, C.c_uint, C.c_float, C.c_float]
        lib.lref_quant_size.argtypes = [C.c_void_p]
        lib.lref_quant_size.restype = C.c_uint
        lib.lref_quant_mapping.argtypes = [C.c_void_p]
        lib.lref_quant_mapping.restype = C.POINTER(C.c_float)
        lib.lref_quant_quantize.argtypes = [C.c_void_p, C.c_float, C.c_uint]
        lib.lref_quant_quantize.restype = C.c_float
        lib.lref_quant_dequantize.argtypes = [C.c_void_p, C.c_float, C.c_uint]
        lib.lref_quant_dequantize.restype = C.c_float
        lib.lref_quant_quantize_n.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_size_t, C.c_uint]
        lib.lref_quant_dequantize_n.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_size_t, C.c_uint]
        lib.lref_quant_transform.argtypes = [C.c_void_p, C.c_void_p, C.c_uint, C.c_uint, C.c_int, C.c_float]
        lib.lref_quant_transform.restype = C.c_int
        lib.lref_encoder_new.argtypes = [C.POINTER(_RefParams), C.c_uint, C.c_uint, C.c_char_p, C.c_size_t]
        lib.lref_encoder_new.restype = C.c_void_p
        lib.lref_encoder_free.argtypes = [C.c_void_p]
        lib.lref_encoder_profile.argtypes = [C.c_void_p]
        lib.lref_encoder_encode.argtypes = [C.c_void_p, C.c_void_p, C.POINTER(C.c_void_p), C.POINTER(C.c_int)]
        lib.lref_encoder_encode.restype = C.c_int
        lib.lref_decoder_new.argtypes = [C.POINTER(_RefParams), C.c_uint, C.c_uint, C.c_char_p, C.c_size_t]
        lib.lref_decoder_new.restype = C.c_void_p
        lib.lref_decoder_free.argtypes = [C.c_void_p]
        lib.lref_decoder_decode.argtypes = [C.c_void_p, C.POINTER(C.c_void_p), C.POINTER(C.c_int), C.c_void_p,
                                            C.POINTER(C.c_int)]
        lib.lref_decoder_decode.restype = C.c_int
        _REFS[key] = lib
    return _REFS[key]


class Reference:
    """The unmodified reference classes (LumaQuantizer / LumaEncoder / LumaDecoder),
    compiled in place from /root/reference with loopback codec + container doubles."""

    def __init__(self, ptf="PQ", ptfBitDepth=11, colorSpace="LUV", colorBitDepth=8, profile=2, bitDepth=12,
                 preScaling=1.0, maxLum=10000.0, minLum=0.005, o0: bool = False):
        self.lib = _ref_lib(o0)
        self.p = _RefParams(PTF[ptf] if isinstance(ptf, str) else ptf,
                            CS[colorSpace] if isinstance(colorSpace, str) else colorSpace, ptfBitDepth, colorBitDepth,
                            profile, bitDepth, preScaling, maxLum, minLum)
        self.quant = self.lib.lref_quant_new()
        self.lib.lref_quant_set(self.quant, self.p.ptf, ptfBitDepth, self.p.color_space, colorBitDepth, maxLum, minLum)
        self._enc = None
        self._dec = None
        self._dims = None
        self.profile = profile

    def close(self):
        if self._enc:
            self.lib.lref_encoder_free(self._enc)
            self._enc = None
        if self._dec:
            self.lib.lref_decoder_free(self._dec)
            self._dec = None
        if self.quant:
            self.lib.lref_quant_free(self.quant)
            self.quant = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # --- LumaQuantizer -----------------------------------------------------------------
    def getSize(self):
        return int(self.lib.lref_quant_size(self.quant))

    def getMapping(self) -> np.ndarray:
        return np.ctypeslib.as_array(self.lib.lref_quant_mapping(self.quant), shape=(self.getSize() + 1,)).copy()

    def quantize(self, val, ch):
        return float(self.lib.lref_quant_quantize(self.quant, float(np.float32(val)), ch))

    def dequantize(self, val, ch):
        return float(self.lib.lref_quant_dequantize(self.quant, float(np.float32(val)), ch))

    def quantize_n(self, vals: np.ndarray, ch: int) -> np.ndarray:
        vals = np.ascontiguousarray(vals, dtype=np.float32)
        out = np.empty_like(vals)
        self.lib.lref_quant_quantize_n(self.quant, vals.ctypes.data, out.ctypes.data, vals.size, ch)
        return out

    def dequantize_n(self, vals: np.ndarray, ch: int) -> np.ndarray:
        vals = np.ascontiguousarray(vals, dtype=np.float32)
        out = np.empty_like(vals)
        self.lib.lref_quant_dequantize_n(self.quant, vals.ctypes.data, out.ctypes.data, vals.size, ch)
        return out

    def transformColorSpace(self, frame: np.ndarray, toCs: bool, sc: float = 1.0) -> bool:
        assert frame.dtype == np.float32 and frame.flags.c_contiguous and frame.shape[0] == 3
        _, h, w = frame.shape
        return bool(self.lib.lref_quant_transform(self.quant, frame.ctypes.data, w, h, int(bool(toCs)), sc))

    # --- LumaEncoder::encode / LumaDecoder::decode --------------------------------------
    def _ensure(self, w, h):
        if self._dims != (w, h):
            if self._enc:
                self.lib.lref_encoder_free(self._enc)
            if self._dec:
                self.lib.lref_decoder_free(self._dec)
            self._enc = self._dec = None
            self._dims = (w, h)

    def encode(self, frame: np.ndarray, strides=None):
        """LumaEncoder::encode(&frame): mutates `frame`; returns the planes handed to the codec."""
        assert frame.dtype == np.float32 and frame.flags.c_contiguous and frame.shape[0] == 3
        _, h, w = frame.shape
        self._ensure(w, h)
        if not self._enc:
            err = C.create_string_buffer(256)
            self._enc = self.lib.lref_encoder_new(C.byref(self.p), w, h, err, 256)
            if not self._enc:
                raise RuntimeError("LumaException: " + err.value.decode())
            self.profile = int(self.lib.lref_encoder_profile(self._enc))
        planes = alloc_planes(w, h, self.profile, strides)
        ptrs, st = _plane_args(planes)
        rc = self.lib.lref_encoder_encode(self._enc, frame.ctypes.data, ptrs, st)
        if rc:
            raise RuntimeError(f"reference encode failed ({rc})")
        return planes

    def decode(self, planes, w: int, h: int) -> np.ndarray:
        """LumaDecoder::decode() on a frame carrying exactly `planes`."""
        self._ensure(w, h)
        if not self._dec:
            err = C.create_string_buffer(256)
            self._dec = self.lib.lref_decoder_new(C.byref(self.p), w, h, err, 256)
            if not self._dec:
                raise RuntimeError("LumaException: " + err.value.decode())
        out = np.empty((3, h, w), dtype=np.float32)
        ptrs, st = _plane_args(planes)
        ds = (C.c_int * 3)()
        rc = self.lib.lref_decoder_decode(self._dec, ptrs, st, out.ctypes.data, ds)
        if rc:
            raise RuntimeError(f"reference decode failed ({rc})")
        self.last_decoder_strides = list(ds)
        return out
