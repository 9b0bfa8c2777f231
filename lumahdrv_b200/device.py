"""Device-resident batch API: frames and planes stay in HBM as torch tensors.

torch is plumbing only (allocation, streams, ``torch.distributed``); every pixel
is produced by the kernels behind ``lumacu_encode_dev`` / ``lumacu_decode_dev``.

Frame batches are ``float32 [n, 3, h, w]`` (n LumaFrames back to back); plane
batches are three ``uint8 [n, rows, stride]`` tensors (n vpx-style pitched planes).
"""
from __future__ import annotations

import ctypes as C

import numpy as np
import torch

from ._lib import LumaException, check
from .luma import Context, LumaQuantizer, plane_dims, vpx_strides

STATS_DTYPE = np.dtype([("sum", "<f8"), ("max", "<f4"), ("min", "<f4")])


CUDA_STREAM_LEGACY = 0x1  # cudaStreamLegacy: the C ABI reads a NULL stream as "the context's own stream"


def _stream_ptr(device) -> int:
    """Handle of torch's current stream; the legacy default stream (handle 0) is passed as cudaStreamLegacy."""
    return int(torch.cuda.current_stream(device).cuda_stream) or CUDA_STREAM_LEGACY


class DeviceTransform:
    """Encode/decode frame batches that already live on one GPU."""

    def __init__(self, device: int | torch.device = 0, ptf="PQ", ptfBitDepth: int = 11, colorSpace="LUV",
                 colorBitDepth: int = 8, maxLum: float = 1e4, minLum: float = 0.005, profile: int = 2,
                 preScaling: float = 1.0, lut: np.ndarray | None = None):
        if not torch.cuda.is_available():
            raise LumaException("DeviceTransform needs a CUDA device (there is no CPU fallback)", 6)
        self.device = torch.device("cuda", device) if isinstance(device, int) else torch.device(device)
        self.index = self.device.index if self.device.index is not None else torch.cuda.current_device()
        self.quant = LumaQuantizer(context=Context(self.index))
        self.quant.setQuantizer(ptf, ptfBitDepth, colorSpace, colorBitDepth, maxLum, minLum)
        if lut is not None:  # e.g. a LUT received from rank 0
            self.quant.setMapping(lut)
            self.quant._upload()
        self.profile = int(profile)
        self.preScaling = float(preScaling)
        self._lib = self.quant._lib

    # ------------------------------------------------------------------ geometry
    def alloc_planes(self, n: int, w: int, h: int, strides=None):
        strides = strides or vpx_strides(w, self.profile)
        return [torch.zeros((n, ph, st), dtype=torch.uint8, device=self.device)
                for (pw, ph), st in zip(plane_dims(w, h, self.profile), strides)]

    def alloc_stats(self, n: int) -> torch.Tensor:
        return torch.zeros((n, STATS_DTYPE.itemsize), dtype=torch.uint8, device=self.device)

    @staticmethod
    def stats_to_numpy(stats: torch.Tensor) -> np.ndarray:
        return stats.cpu().numpy().view(STATS_DTYPE).reshape(-1)

    @staticmethod
    def _plane_args(planes):
        n = planes[0].shape[0]
        for p in planes:
            if p.dtype != torch.uint8 or p.dim() != 3 or not p.is_contiguous() or p.shape[0] != n:
                raise LumaException("planes must be contiguous uint8 [n, rows, stride] tensors", 1)
        ptrs = (C.c_void_p * 3)(*[p.data_ptr() for p in planes])
        strides = (C.c_int32 * 3)(*[p.shape[2] for p in planes])
        fstr = (C.c_size_t * 3)(*[p.shape[1] * p.shape[2] for p in planes])
        return n, ptrs, strides, fstr

    # ------------------------------------------------------------------ transform
    def encode(self, rgb: torch.Tensor, planes=None, stats: torch.Tensor | None = None,
               write_back: torch.Tensor | None = None):
        """n x LumaEncoder::encode (minus run()) in one launch.  Asynchronous on the current stream."""
        if rgb.dtype != torch.float32 or rgb.dim() != 4 or rgb.shape[1] != 3 or not rgb.is_contiguous() or not rgb.is_cuda:
            raise LumaException("rgb must be a contiguous CUDA float32 [n, 3, h, w] tensor", 1)
        n, _, h, w = rgb.shape
        planes = planes if planes is not None else self.alloc_planes(n, w, h)
        pn, ptrs, strides, fstr = self._plane_args(planes)
        if pn != n:
            raise LumaException("frame and plane batch sizes differ", 1)
        hnd = self.quant.ctx.handle
        check(self._lib.lumacu_encode_dev(hnd, rgb.data_ptr(), write_back.data_ptr() if write_back is not None else None,
                                          w, h, self.profile, self.preScaling, ptrs, strides, n, 3 * h * w, fstr,
                                          stats.data_ptr() if stats is not None else None, _stream_ptr(self.device)),
              hnd, "lumacu_encode_dev")
        return planes

    def decode(self, planes, w: int, h: int, out: torch.Tensor | None = None) -> torch.Tensor:
        """n x LumaDecoder::decode (minus run()) in one launch.  Asynchronous on the current stream."""
        n, ptrs, strides, fstr = self._plane_args(planes)
        if out is None:
            out = torch.empty((n, 3, h, w), dtype=torch.float32, device=self.device)
        hnd = self.quant.ctx.handle
        check(self._lib.lumacu_decode_dev(hnd, ptrs, strides, w, h, self.profile, self.preScaling, out.data_ptr(), n,
                                          3 * h * w, fstr, _stream_ptr(self.device)), hnd, "lumacu_decode_dev")
        return out

    # ------------------------------------------------------------------ frame sources
    def test_frame(self, w: int, h: int, out: torch.Tensor | None = None) -> torch.Tensor:
        """ExrInterface::testFrame(frame, w, h) generated in device memory ([3, h, w] f32)."""
        if out is None:
            out = torch.empty((3, h, w), dtype=torch.float32, device=self.device)
        hnd = self.quant.ctx.handle
        check(self._lib.lumacu_test_frame_dev(hnd, out.data_ptr(), w, h, _stream_ptr(self.device)), hnd, "lumacu_test_frame_dev")
        return out

    def half_rgba_to_frame(self, rgba: torch.Tensor, channels: int = 7, out: torch.Tensor | None = None) -> torch.Tensor:
        """[h, w, 4] float16 (Imf::Rgba pixels) -> [3, h, w] f32, like ExrInterface::readFrame's pixel loop."""
        if rgba.dtype != torch.float16 or rgba.dim() != 3 or rgba.shape[2] != 4 or not rgba.is_contiguous() or not rgba.is_cuda:
            raise LumaException("rgba must be a contiguous CUDA float16 [h, w, 4] tensor", 1)
        h, w, _ = rgba.shape
        if out is None:
            out = torch.empty((3, h, w), dtype=torch.float32, device=self.device)
        hnd = self.quant.ctx.handle
        check(self._lib.lumacu_half_rgba_to_frame_dev(hnd, rgba.data_ptr(), w, h, int(channels), out.data_ptr(),
                                                      _stream_ptr(self.device)), hnd, "lumacu_half_rgba_to_frame_dev")
        return out

    def frame_to_half_rgba(self, rgb: torch.Tensor, out: torch.Tensor | None = None) -> torch.Tensor:
        """[3, h, w] f32 -> [h, w, 4] float16 (Imf::Rgba pixels, alpha 0), like ExrInterface::writeFrame's pixel loop."""
        if rgb.dtype != torch.float32 or rgb.dim() != 3 or rgb.shape[0] != 3 or not rgb.is_contiguous() or not rgb.is_cuda:
            raise LumaException("rgb must be a contiguous CUDA float32 [3, h, w] tensor", 1)
        _, h, w = rgb.shape
        if out is None:
            out = torch.empty((h, w, 4), dtype=torch.float16, device=self.device)
        hnd = self.quant.ctx.handle
        check(self._lib.lumacu_frame_to_half_rgba_dev(hnd, rgb.data_ptr(), w, h, out.data_ptr(), _stream_ptr(self.device)),
              hnd, "lumacu_frame_to_half_rgba_dev")
        return out

    def pfs_xyz_to_frame(self, x: torch.Tensor, y: torch.Tensor, z: torch.Tensor, out: torch.Tensor | None = None) -> torch.Tensor:
        """PfsInterface::readFrame's colour step (src/pfs_interface.cpp:80-102): the X, Y, Z channel arrays of a PFS
        frame ([h, w] f32 each) -> planar RGB frame [3, h, w]."""
        for c in (x, y, z):
            if c.dtype != torch.float32 or c.dim() != 2 or not c.is_contiguous() or not c.is_cuda or c.shape != x.shape:
                raise LumaException("x, y, z must be contiguous CUDA float32 [h, w] tensors of one size", 1)
        h, w = x.shape
        if out is None:
            out = torch.empty((3, h, w), dtype=torch.float32, device=self.device)
        hnd = self.quant.ctx.handle
        check(self._lib.lumacu_pfs_xyz_to_frame_dev(hnd, x.data_ptr(), y.data_ptr(), z.data_ptr(), w, h, out.data_ptr(),
                                                    _stream_ptr(self.device)), hnd, "lumacu_pfs_xyz_to_frame_dev")
        return out

    def frame_to_pfs_xyz(self, rgb: torch.Tensor):
        """PfsInterface::writeFrame's colour step (src/pfs_interface.cpp:136-140): planar RGB frame -> X, Y, Z arrays."""
        if rgb.dtype != torch.float32 or rgb.dim() != 3 or rgb.shape[0] != 3 or not rgb.is_contiguous() or not rgb.is_cuda:
            raise LumaException("rgb must be a contiguous CUDA float32 [3, h, w] tensor", 1)
        _, h, w = rgb.shape
        xyz = torch.empty((3, h, w), dtype=torch.float32, device=self.device)
        hnd = self.quant.ctx.handle
        check(self._lib.lumacu_frame_to_pfs_xyz_dev(hnd, rgb.data_ptr(), w, h, xyz[0].data_ptr(), xyz[1].data_ptr(),
                                                    xyz[2].data_ptr(), _stream_ptr(self.device)), hnd, "lumacu_frame_to_pfs_xyz_dev")
        return xyz[0], xyz[1], xyz[2]

    @property
    def launch_count(self) -> int:
        return self.quant.ctx.launch_count


def broadcast_quantizer_lut(lut: np.ndarray | None, n_entries: int, device: torch.device, src: int = 0) -> np.ndarray:
    """Share rank ``src``'s host-built LUT with every rank (NCCL broadcast of <= 256 KB).

    The LUT is built with the host libm on one rank so that all shards quantise with the
    same table even if host libms differ; no pixel data ever crosses GPUs."""
    import torch.distributed as dist

    t = torch.empty(n_entries, dtype=torch.float32, device=device)
    if dist.get_rank() == src:
        t.copy_(torch.from_numpy(np.ascontiguousarray(lut, dtype=np.float32)))
    dist.broadcast(t, src=src)
    return t.cpu().numpy()
