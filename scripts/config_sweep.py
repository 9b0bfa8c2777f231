#!/usr/bin/env python
"""Device-resident throughput of the BASELINE.json configurations other than the headline one (parity for all of
them is covered by tests/; these numbers are context for DESIGN.md, not bench lines)."""
import json
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))

import torch  # noqa: E402

from lumahdrv_b200.device import DeviceTransform  # noqa: E402

CONFIGS = [
    ("cfg2 1080p PQ Lu'v' 11/8 p2", 1920, 1080, 64, dict(ptf="PQ", ptfBitDepth=11, colorSpace="LUV", colorBitDepth=8, profile=2)),
    ("headline 4K PQ Lu'v' 11/8 p2", 3840, 2160, 32, dict(ptf="PQ", ptfBitDepth=11, colorSpace="LUV", colorBitDepth=8, profile=2)),
    ("cfg3 4K PQ-10 YCbCr 10 p2", 3840, 2160, 8, dict(ptf="PQ", ptfBitDepth=10, colorSpace="YCBCR", colorBitDepth=10, profile=2)),
    ("cfg3' 4K PQ-10 YCbCr HDR10 recipe", 3840, 2160, 8, dict(ptf="PQ", ptfBitDepth=10, colorSpace="YCBCR", colorBitDepth=10, profile=2,
                                                              maxLum=1000.0, minLum=0.01, preScaling=20.0)),
    ("cfg4 4K LOG-12 Lu'v' 8 p2", 3840, 2160, 32, dict(ptf="LOG", ptfBitDepth=12, colorSpace="LUV", colorBitDepth=8, profile=2)),
    ("cfg4' 4K LOG-12 Lu'v' 12 p2", 3840, 2160, 32, dict(ptf="LOG", ptfBitDepth=12, colorSpace="LUV", colorBitDepth=12, profile=2)),
    ("cfg5 8K PQ Lu'v' 11/8 p2 (+stats)", 7680, 4320, 8, dict(ptf="PQ", ptfBitDepth=11, colorSpace="LUV", colorBitDepth=8, profile=2)),
    ("4K PQ-12 Lu'v' 12 p3 (4:4:4)", 3840, 2160, 16, dict(ptf="PQ", ptfBitDepth=12, colorSpace="LUV", colorBitDepth=12, profile=3)),
    ("4K PQ-12 Lu'v' 8 p2 (wide LUT, 4:2:0)", 3840, 2160, 16, dict(ptf="PQ", ptfBitDepth=12, colorSpace="LUV", colorBitDepth=8, profile=2)),
    ("4K PQ-16 Lu'v' 8 p2 (-pb 16: global direct table)", 3840, 2160, 8, dict(ptf="PQ", ptfBitDepth=16, colorSpace="LUV", colorBitDepth=8, profile=2)),
    ("4K PQ-11 Lu'v' 8 p3 (4:4:4, 8-bit chroma)", 3840, 2160, 16, dict(ptf="PQ", ptfBitDepth=11, colorSpace="LUV", colorBitDepth=8, profile=3)),
    ("4K PQ-11 Lu'v' 10 p3 (4:4:4, 10-bit chroma)", 3840, 2160, 16, dict(ptf="PQ", ptfBitDepth=11, colorSpace="LUV", colorBitDepth=10, profile=3)),
    ("4K PQ-11 Lu'v' 10 p2 (4:2:0, 10-bit chroma)", 3840, 2160, 16, dict(ptf="PQ", ptfBitDepth=11, colorSpace="LUV", colorBitDepth=10, profile=2)),
    ("4K PQ-11 XYZ p2", 3840, 2160, 16, dict(ptf="PQ", ptfBitDepth=11, colorSpace="XYZ", colorBitDepth=8, profile=2)),
    ("4K PQ-11 RGB p2", 3840, 2160, 16, dict(ptf="PQ", ptfBitDepth=11, colorSpace="RGB", colorBitDepth=8, profile=2)),
    ("1080p PQ-8 Lu'v' 8 p0 (8-bit 4:2:0)", 1920, 1080, 64, dict(ptf="PQ", ptfBitDepth=8, colorSpace="LUV", colorBitDepth=8, profile=0)),
]


def main():
    dev = torch.device("cuda", 0)
    out = []
    only = sys.argv[1] if len(sys.argv) > 1 else ""
    for name, w, h, F, kw in CONFIGS:
        if only and only not in name:
            continue
        t = DeviceTransform(0, **kw)
        if len(sys.argv) > 2:  # encoder tuning variant (lumacu_set_tuning), e.g. 1004 = bucket/threshold search instead of the direct tables
            t.quant.ctx.set_tuning(int(sys.argv[2]), 0, 0)
        g = torch.Generator(device=dev).manual_seed(7)
        rgb = 0.005 * torch.pow(torch.tensor(2.0e6, device=dev), torch.rand((F, 3, h, w), generator=g, device=dev))
        planes = t.alloc_planes(F, w, h)
        res = torch.empty_like(rgb)
        stats = t.alloc_stats(F)
        for _ in range(3):
            t.encode(rgb, planes=planes, stats=stats)
            t.decode(planes, w, h, out=res)
        torch.cuda.synchronize()
        n = 20
        ev = [torch.cuda.Event(enable_timing=True) for _ in range(2 * n + 1)]
        ev[0].record()
        for i in range(n):
            t.encode(rgb, planes=planes, stats=stats)
            ev[2 * i + 1].record()
            t.decode(planes, w, h, out=res)
            ev[2 * i + 2].record()
        torch.cuda.synchronize()
        enc = sorted(ev[2 * i].elapsed_time(ev[2 * i + 1]) for i in range(n))[n // 2]
        dec = sorted(ev[2 * i + 1].elapsed_time(ev[2 * i + 2]) for i in range(n))[n // 2]
        px = F * w * h
        prof = kw["profile"]
        bpp = {0: 12 + 1.5, 1: 12 + 3, 2: 12 + 3, 3: 12 + 6}[prof]
        info = t.quant.search_info() if hasattr(t.quant, "search_info") else {}
        row = {"config": name, "frames": F, "enc_us": enc * 1e3, "dec_us": dec * 1e3, "roundtrip_mpx_s": px / (enc + dec) / 1e3,
               "enc_gbs": bpp * px / enc / 1e6, "dec_gbs": bpp * px / dec / 1e6, "bytes_per_px": bpp, "search": info}
        out.append(row)
        print(f"{name:40s} F={F:3d} enc {enc*1e3:8.1f} us ({row['enc_gbs']:6.0f} GB/s)  dec {dec*1e3:8.1f} us ({row['dec_gbs']:6.0f} GB/s)  "
              f"round trip {row['roundtrip_mpx_s']:9.0f} Mpx/s  {info}", flush=True)
        del rgb, planes, res, t
        torch.cuda.empty_cache()
    print(json.dumps(out))


if __name__ == "__main__":
    main()
