"""Frame-level sharding across GPUs (one process per GPU).

Frames are independent given the quantizer, so the only thing ranks ever
exchange is the host-built LUT plus a few scalars, once per stream; no pixel data crosses
NVLink and there is no data-path collective."""
from __future__ import annotations

import numpy as np


def frame_shard(n_frames: int, rank: int, world: int) -> range:
    """Contiguous block of frame indices owned by `rank` (sizes differ by at most one)."""
    base, extra = divmod(int(n_frames), int(world))
    start = rank * base + min(rank, extra)
    return range(start, start + base + (1 if rank < extra else 0))


def pack_quantizer(lut: np.ndarray, max_val_color: int, color_space: int, max_lum: float, min_lum: float,
                   pre_scaling: float, profile: int) -> np.ndarray:
    """LUT + parameters as one float32 vector (what rank 0 broadcasts)."""
    head = np.array([lut.size, max_val_color, color_space, profile], dtype=np.float32)
    tail = np.array([max_lum, min_lum, pre_scaling], dtype=np.float32)
    return np.concatenate([head, tail, np.ascontiguousarray(lut, dtype=np.float32)])


def unpack_quantizer(vec: np.ndarray) -> dict:
    n = int(vec[0])
    return {"max_val_color": int(vec[1]), "color_space": int(vec[2]), "profile": int(vec[3]), "max_lum": float(vec[4]),
            "min_lum": float(vec[5]), "pre_scaling": float(vec[6]), "lut": np.array(vec[7:7 + n], dtype=np.float32)}


def broadcast_quantizer(vec: np.ndarray | None, length: int, device, src: int = 0) -> np.ndarray:
    """Broadcast the packed quantizer from `src` (NCCL on CUDA tensors, gloo on CPU tensors)."""
    import torch
    import torch.distributed as dist

    t = torch.empty(length, dtype=torch.float32, device=device)
    if dist.get_rank() == src:
        t.copy_(torch.from_numpy(np.ascontiguousarray(vec, dtype=np.float32)))
    dist.broadcast(t, src=src)
    return t.cpu().numpy()
