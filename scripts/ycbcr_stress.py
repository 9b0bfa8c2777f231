#!/usr/bin/env python
"""Randomized comparison, on the GPU, of the CS_YCBCR tuned kernels WITH their device-built tables (PQ decode step table,
outer encode power, v-keyed luma search, half-float input table) against the same kernels evaluating every powf per
pixel (lumacu_set_pq_tables(0)) and against the generic kernels: many quantizers (transfer function, bit depths, Lmax,
preScaling), content kinds (float noise over 18 decades, half-float values, mixed, specials) and random code planes.
Prints pixels compared and mismatches (must be 0).

    python scripts/ycbcr_stress.py [--seconds 120]
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
    w, h, n = 1920, 1080, 4
    g = torch.Generator(device=dev).manual_seed(99)
    quantizers = [("PQ", 10, 10, 1e4, 0.005, 1.0), ("PQ", 10, 10, 1000.0, 0.01, 20.0), ("PQ", 8, 8, 1e4, 0.005, 1.0),
                  ("PQ", 11, 8, 4000.0, 0.005, 0.5), ("PQ", 12, 12, 1e4, 0.005, 1.0), ("LOG", 10, 10, 1e4, 0.005, 1.0),
                  ("LOG", 12, 10, 1e4, 0.01, 3.0), ("LINEAR", 10, 10, 1e4, 0.005, 1.0), ("PSI", 11, 8, 1e4, 0.005, 1.0)]
    px = bad_enc = bad_dec = rounds = 0
    t0 = time.perf_counter()
    info = {}
    while time.perf_counter() - t0 < a.seconds:
        for ptf, bits, cbits, lmax, lmin, sc in quantizers:
            profile = 2 if bits > 8 else 0
            t = DeviceTransform(0, ptf=ptf, ptfBitDepth=bits, colorSpace="YCBCR", colorBitDepth=cbits, maxLum=lmax, minLum=lmin,
                                preScaling=sc, profile=profile)
            ctx = t.quant.ctx
            u = torch.rand((n, 3, h, w), generator=g, device=dev)
            kind = rounds % 4
            if kind == 0:
                rgb = torch.pow(torch.tensor(10.0, device=dev), u * 18 - 12)          # 1e-12 .. 1e6
            elif kind == 1:
                rgb = (0.005 * torch.pow(torch.tensor(2.0e6, device=dev), u)).half().float()  # EXR-like
            elif kind == 2:
                rgb = 0.005 * torch.pow(torch.tensor(2.0e6, device=dev), u)
                m = torch.rand((n, 3, h, w), generator=g, device=dev) < 0.5
                rgb = torch.where(m, rgb.half().float(), rgb)
            else:
                rgb = 100.0 * u
                idx = torch.randint(0, rgb.numel(), (20000,), generator=g, device=dev)
                vals = torch.tensor([0.0, -1.0, float("inf"), float("nan"), 1e-45, 1e-38, 65504.0, 6e-8, 3e38], device=dev)
                rgb.view(-1)[idx] = vals[torch.randint(0, vals.numel(), (20000,), generator=g, device=dev)]
            rgb = (rgb / sc).contiguous()
            ctx.set_pq_tables(False)
            ref = [p.clone() for p in t.encode(rgb)]
            ctx.set_pq_tables(True)
            for stats in (None, t.alloc_stats(n)):                                    # v-keyed search / luminance search
                got = t.encode(rgb, stats=stats)
                for x1, x2 in zip(got, ref):
                    bad_enc += int((x1 != x2).sum().item())
            ctx.set_kernel_path(1)
            gen = t.encode(rgb)
            ctx.set_kernel_path(0)
            for x1, x2 in zip(gen, ref):
                bad_enc += int((x1 != x2).sum().item())
            # decode random code planes (mostly in range)
            planes = t.alloc_planes(n, w, h)
            for pl in planes:
                if profile > 1:
                    codes = torch.randint(0, 1 << max(bits, cbits), pl.shape[:2] + (pl.shape[2] // 2,), generator=g, device=dev, dtype=torch.int32)
                    pl.copy_(torch.stack([(codes & 0xFF), (codes >> 8)], dim=-1).to(torch.uint8).reshape(pl.shape))
                else:
                    pl.copy_(torch.randint(0, 256, pl.shape, generator=g, device=dev, dtype=torch.int32).to(torch.uint8))
            ctx.set_pq_tables(False)
            dref = t.decode(planes, w, h).clone()
            ctx.set_pq_tables(True)
            dgot = t.decode(planes, w, h)
            same = (dgot.view(torch.int32) == dref.view(torch.int32)) | (torch.isnan(dgot) & torch.isnan(dref))
            bad_dec += int((~same).sum().item())
            px += 2 * n * w * h
            info[(ptf, bits)] = t.quant.search_info()
            del t, rgb, u, ref, got, gen, planes, dref, dgot
        rounds += 1
    print(f"ycbcr_stress: {px / 1e9:.2f} Gpixel in {rounds} rounds x {len(quantizers)} quantizers x 4 content kinds: "
          f"{bad_enc} mismatching plane bytes, {bad_dec} mismatching decoded floats")
    return 1 if (bad_enc or bad_dec) else 0


if __name__ == "__main__":
    sys.exit(main())
