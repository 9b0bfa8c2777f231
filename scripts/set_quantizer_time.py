#!/usr/bin/env python
"""How long do lumacu_set_quantizer and a new context take?  CS_YCBCR builds 45 MB of PQ tables, the v-keyed search table and
runs the exhaustive val / Lmax division check (1.35e9 operands) per new Lmax; Lu'v' only derives thresholds on the host.
Measured on B200: 1.6 ms for a CS_YCBCR quantizer with a new Lmax, 0.4 ms with a known one, 0.23 ms for Lu'v', 8-14 ms for a
new context + quantizer."""
import sys
import time
from pathlib import Path

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import torch
from lumahdrv_b200.device import DeviceTransform
t = DeviceTransform(0, ptf="PQ", ptfBitDepth=10, colorSpace="YCBCR", colorBitDepth=10)
for lm in (1e4, 1000.0, 4000.0, 1e4, 1234.5, 1e4):
    torch.cuda.synchronize(); t0 = time.perf_counter()
    t.quant.setQuantizer("PQ", 10, "YCBCR", 10, lm, 0.005); t.quant._upload()
    torch.cuda.synchronize(); print(f"set_quantizer YCbCr Lmax={lm}: {(time.perf_counter()-t0)*1e3:.2f} ms", flush=True)
t2 = DeviceTransform(0)
torch.cuda.synchronize(); t0 = time.perf_counter()
t2.quant.setQuantizer("PQ", 11, "LUV", 8, 1e4, 0.005); t2.quant._upload()
torch.cuda.synchronize(); print(f"set_quantizer LUV: {(time.perf_counter()-t0)*1e3:.2f} ms")
for i in range(4):
    torch.cuda.synchronize(); t0 = time.perf_counter()
    t3 = DeviceTransform(0, ptf="PQ", ptfBitDepth=10, colorSpace="YCBCR", colorBitDepth=10, maxLum=1e4 if i % 2 else 1000.0)
    torch.cuda.synchronize(); print(f"new context + YCbCr quantizer: {(time.perf_counter()-t0)*1e3:.2f} ms", flush=True)
    del t3
import numpy as np
g = torch.Generator(device="cuda").manual_seed(1)
rgb = (100.0 * torch.rand((4, 3, 1080, 1920), generator=g, device="cuda")).contiguous()
idx = torch.randint(0, rgb.numel(), (20000,), generator=g, device="cuda")
vals = torch.tensor([0.0, -1.0, float("inf"), float("nan"), 1e-45, 1e-38, 65504.0, 6e-8, 3e38], device="cuda")
rgb.view(-1)[idx] = vals[torch.randint(0, vals.numel(), (20000,), generator=g, device="cuda")]
for tables in (True, False):
    t.quant.ctx.set_pq_tables(tables)
    for _ in range(2):
        torch.cuda.synchronize(); t0 = time.perf_counter()
        t.encode(rgb)
        torch.cuda.synchronize(); print(f"encode 4 x 1080p with specials, tables={tables}: {(time.perf_counter()-t0)*1e3:.2f} ms", flush=True)
