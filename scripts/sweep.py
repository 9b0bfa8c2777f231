#!/usr/bin/env python
"""Tuning sweep over the extra instantiations of the headline kernels (lumacu_set_tuning): prints encode /
decode time per variant on the bench workload (4K PQ Lu'v' 11/8, profile 2, F frames per launch)."""
import argparse
import json
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))

import torch  # noqa: E402

from lumahdrv_b200._lib import check  # noqa: E402
from lumahdrv_b200.device import DeviceTransform  # noqa: E402

W, H = 3840, 2160


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--frames", type=int, default=8)
    ap.add_argument("--iters", type=int, default=30)
    ap.add_argument("--enc", default="0,1004,3,5,84,34,44,54")
    ap.add_argument("--dec", default="0,3,5,14")
    ap.add_argument("--caps", default="0")
    ap.add_argument("--preroll", type=float, default=0.0, help="seconds of the same work before timing (sustained, power-capped regime)")
    ap.add_argument("--width", type=int, default=W)
    ap.add_argument("--height", type=int, default=H)
    a = ap.parse_args()
    w, h = a.width, a.height
    dev = torch.device("cuda", 0)
    t = DeviceTransform(0)
    F = a.frames
    g = torch.Generator(device=dev).manual_seed(1234)
    rgb = 0.005 * torch.pow(torch.tensor(2.0e6, device=dev), torch.rand((F, 3, h, w), generator=g, device=dev))
    planes = t.alloc_planes(F, w, h)
    out = torch.empty_like(rgb)
    stats = t.alloc_stats(F)
    ref_planes = None
    px = F * w * h
    res = []

    def timeit(fn):
        import time
        t0 = time.perf_counter()
        n = 0
        while n < 5 or time.perf_counter() - t0 < a.preroll:
            fn()
            n += 1
            if n % 8 == 0:
                torch.cuda.synchronize()
        torch.cuda.synchronize()
        ev = [torch.cuda.Event(enable_timing=True) for _ in range(a.iters + 1)]
        ev[0].record()
        for i in range(a.iters):
            fn()
            ev[i + 1].record()
        torch.cuda.synchronize()
        ts = sorted(ev[i].elapsed_time(ev[i + 1]) for i in range(a.iters))
        return ts[len(ts) // 2], ts[0]

    hnd = t.quant.ctx.handle
    for cap in [int(c) for c in a.caps.split(",")]:
        for v in [int(x) for x in a.enc.split(",")]:
            check(t._lib.lumacu_set_tuning(hnd, v, 0, cap), hnd, "set_tuning")
            for p in planes:
                p.zero_()
            med, best = timeit(lambda: t.encode(rgb, planes=planes, stats=stats))
            cur = [p.clone() for p in planes]
            if ref_planes is None:
                ref_planes = cur
            same = all(torch.equal(x, y) for x, y in zip(cur, ref_planes))
            res.append(("enc", v, cap, med, best, 15 * px / med / 1e6, same))
            print(f"enc variant {v:2d} cap {cap}: median {med*1e3:8.1f} us  best {best*1e3:8.1f} us  "
                  f"{15*px/med/1e6:7.1f} GB/s  same={same}", flush=True)
        ref_out = None
        for v in [int(x) for x in a.dec.split(",")]:
            check(t._lib.lumacu_set_tuning(hnd, 0, v, cap), hnd, "set_tuning")
            out.zero_()
            med, best = timeit(lambda: t.decode(ref_planes, w, h, out=out))
            if ref_out is None:
                ref_out = out.clone()
            same = torch.equal(out.view(torch.int32), ref_out.view(torch.int32))
            res.append(("dec", v, cap, med, best, 15 * px / med / 1e6, same))
            print(f"dec variant {v:2d} cap {cap}: median {med*1e3:8.1f} us  best {best*1e3:8.1f} us  "
                  f"{15*px/med/1e6:7.1f} GB/s  same={same}", flush=True)
    print(json.dumps(res))


if __name__ == "__main__":
    main()
