"""Frame-level sharding across GPUs (one process per GPU).

Frames are independent given the quantizer, so the only thing ranks ever
exchange is the quantizer metadata in the reference's own wire format (the attachment payloads the encoder
writes into the Matroska file), once per stream; no pixel data crosses NVLink and there is no data-path
collective."""
from __future__ import annotations

import numpy as np


def frame_shard(n_frames: int, rank: int, world: int) -> range:
    """Contiguous block of frame indices owned by `rank` (sizes differ by at most one)."""
    base, extra = divmod(int(n_frames), int(world))
    start = rank * base + min(rank, extra)
    return range(start, start + base + (1 if rank < extra else 0))


def pack_quantizer(lut: np.ndarray, max_val_color: int, color_space: int, max_lum: float, min_lum: float,
                   pre_scaling: float, profile: int, ptf: int = 1) -> np.ndarray:
    """The quantizer as the reference's own wire format: the payloads of Matroska attachments 430..436
    (src/luma_encoder.cpp:78-106) back to back, as bytes (lumacu_metadata_pack).  `profile` is not part of the
    reference's metadata (the decoder learns it from the VP9 stream); it travels in one extra trailing byte."""
    import ctypes as C

    from . import _lib

    lut = np.ascontiguousarray(lut, dtype=np.float32)
    bits = int(lut.size).bit_length() - 1
    if lut.size != 1 << bits:
        raise _lib.LumaException("LUT length must be a power of two", 1)
    cbits = int(max_val_color + 1).bit_length() - 1
    m = _lib.Metadata(bits, cbits, int(ptf), int(color_space), float(pre_scaling), float(max_lum), float(min_lum))
    used = C.c_size_t(0)
    lib = _lib.lib()
    lib.lumacu_metadata_pack(C.byref(m), lut.ctypes.data, lut.size, None, 0, C.byref(used))
    blob = np.zeros(used.value + 1, dtype=np.uint8)
    _lib.check(lib.lumacu_metadata_pack(C.byref(m), lut.ctypes.data, lut.size, blob.ctypes.data, used.value, C.byref(used)),
               None, "lumacu_metadata_pack")
    blob[-1] = profile
    return blob


def packed_quantizer_size(ptf_bit_depth: int) -> int:
    """Bytes pack_quantizer produces for a 2^ptf_bit_depth-entry LUT (what the receiving ranks allocate)."""
    return 7 * 8 + 4 * 4 + ((1 << ptf_bit_depth) - 1) * 4 + 4 + 8 + 1


def unpack_quantizer(blob: np.ndarray) -> dict:
    """What LumaDecoder::initialize does with the attachments (src/luma_decoder.cpp:79-122): scalars, table rebuilt
    from them with the local libm, stored entries overlaid (lumacu_metadata_unpack)."""
    import ctypes as C

    from . import _lib

    blob = np.ascontiguousarray(blob, dtype=np.uint8)
    m = _lib.Metadata()
    lut = np.zeros(1 << 16, dtype=np.float32)
    n = C.c_uint32(0)
    _lib.check(_lib.lib().lumacu_metadata_unpack(blob.ctypes.data, blob.size - 1, C.byref(m), lut.ctypes.data, lut.size,
                                                  C.byref(n)), None, "lumacu_metadata_unpack")
    return {"max_val_color": (1 << m.color_bit_depth) - 1, "color_space": int(m.color_space), "profile": int(blob[-1]),
            "max_lum": float(m.max_lum), "min_lum": float(m.min_lum), "pre_scaling": float(m.pre_scaling), "ptf": int(m.ptf),
            "ptf_bit_depth": int(m.ptf_bit_depth), "color_bit_depth": int(m.color_bit_depth), "lut": lut[:n.value].copy()}


def broadcast_quantizer(blob: np.ndarray | None, length: int, device, src: int = 0) -> np.ndarray:
    """Broadcast the packed quantizer bytes from `src` (NCCL on CUDA tensors, gloo on CPU tensors)."""
    import torch
    import torch.distributed as dist

    t = torch.empty(length, dtype=torch.uint8, device=device)
    if dist.get_rank() == src:
        t.copy_(torch.from_numpy(np.ascontiguousarray(blob, dtype=np.uint8)))
    dist.broadcast(t, src=src)
    return t.cpu().numpy()
