#!/usr/bin/env python
"""cfg3 decode (4K PQ-10 YCbCr, 8 frames): green of every other pixel evaluated (default) vs of every pixel (decoder tuning
2000), on float noise and on the reference's test pattern; identical bits either way."""
import sys
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import torch  # noqa: E402
from lumahdrv_b200.device import DeviceTransform  # noqa: E402

dev = torch.device("cuda", 0)
w, h, F = 3840, 2160, 8
t = DeviceTransform(0, ptf="PQ", ptfBitDepth=10, colorSpace="YCBCR", colorBitDepth=10)
g = torch.Generator(device=dev).manual_seed(7)
noise = 0.005 * torch.pow(torch.tensor(2.0e6, device=dev), torch.rand((F, 3, h, w), generator=g, device=dev))
pattern = t.test_frame(w, h)[None].expand(F, -1, -1, -1).contiguous()
for name, rgb in (("noise", noise), ("test pattern", pattern)):
    planes = t.encode(rgb)
    outs = []
    for tune in (0, 2000, 0, 2000):
        t.quant.ctx.set_tuning(0, tune, 0)
        out = torch.empty_like(rgb)
        for _ in range(3):
            t.decode(planes, w, h, out=out)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize()
        e0.record()
        for _ in range(20):
            t.decode(planes, w, h, out=out)
        e1.record()
        torch.cuda.synchronize()
        print(f"{name}: decoder tuning {tune}: {e0.elapsed_time(e1) / 20 * 1e3:.1f} us per {F} frames", flush=True)
        outs.append(out.clone())
    assert torch.equal(outs[0].view(torch.int32), outs[1].view(torch.int32))
