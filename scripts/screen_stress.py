#!/usr/bin/env python
"""Long randomized comparison of the screened-chroma encode kernel (lumacu_set_tuning 67) with the exact-chain kernel
(tuning 4) on the GPU: many seeds, dynamic ranges, colour casts, chroma depths and preScaling values.  Prints the number
of pixels compared and of mismatching plane bytes (must be 0).

    python scripts/screen_stress.py [--seconds 120]
"""
import argparse
import sys
import time
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))

import torch  # noqa: E402

from lumahdrv_b200.device import DeviceTransform  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--seconds", type=float, default=120.0)
    a = ap.parse_args()
    dev = torch.device("cuda", 0)
    w, h, n = 3840, 2160, 8
    total_px = 0
    bad = 0
    rounds = 0
    t0 = time.perf_counter()
    g = torch.Generator(device=dev).manual_seed(2024)
    while time.perf_counter() - t0 < a.seconds:
        for cbits in (8, 10, 12):
            for sc in (1.0, 0.37):
                t = DeviceTransform(0, colorBitDepth=cbits, preScaling=sc)
                u = torch.rand((n, 3, h, w), generator=g, device=dev)
                kind = rounds % 6
                if kind == 0:    # log-uniform over 7 decades
                    rgb = 0.001 * torch.pow(torch.tensor(1.0e7, device=dev), u)
                elif kind == 1:  # narrow range around a random level: chroma nearly constant -> many samples near one t
                    base = torch.pow(torch.tensor(10.0, device=dev), torch.rand((n, 3, 1, 1), generator=g, device=dev) * 5 - 2)
                    rgb = base * (1 + 1e-3 * u)
                elif kind == 2:  # smooth ramps (every t value along a line) with a little noise
                    x = torch.linspace(0, 1, w, device=dev).view(1, 1, 1, w)
                    y = torch.linspace(0, 1, h, device=dev).view(1, 1, h, 1)
                    rgb = (0.01 + 5000 * x * y).expand(n, 3, h, w) * (0.5 + u)
                elif kind == 3:  # strong colour casts (one or two channels tiny)
                    scale = torch.pow(torch.tensor(10.0, device=dev), torch.randint(-6, 1, (n, 3, 1, 1), generator=g, device=dev).float())
                    rgb = 100.0 * u * scale
                elif kind == 4:  # very dark: around and below the 1e-4 clamp
                    rgb = 1e-3 * u * u * u
                else:            # very bright: around the 9e7 precondition and the 1e8 clamp
                    rgb = 3e8 * u * u
                rgb = (rgb / sc).contiguous()
                t.quant.ctx.set_tuning(4)
                ref = [p.clone() for p in t.encode(rgb)]
                t.quant.ctx.set_tuning(67)
                got = t.encode(rgb)
                for x1, x2 in zip(got, ref):
                    bad += int((x1 != x2).sum().item())
                total_px += n * w * h
                del rgb, u, ref, got, t
        rounds += 1
    print(f"screen_stress: {total_px / 1e9:.2f} Gpixel compared in {rounds} rounds (6 content kinds x 3 chroma depths x 2 preScaling), "
          f"{bad} mismatching plane bytes")
    return 1 if bad else 0


if __name__ == "__main__":
    sys.exit(main())
