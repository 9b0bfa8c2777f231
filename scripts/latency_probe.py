#!/usr/bin/env python
"""Single-frame launch time of the tuned kernels (device-resident frame, CUDA events around every launch, an L2 flush
between launches so that no launch finds its input in L2): median / best in microseconds for 1080p, 4K and 8K.

    python scripts/latency_probe.py [--iters 60] [--enc VARIANT]
"""
import argparse
import json
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))

import torch  # noqa: E402

from lumahdrv_b200.device import DeviceTransform  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--iters", type=int, default=60)
    ap.add_argument("--enc", type=int, default=0)
    ap.add_argument("--cap", type=int, default=0, help="lumacu_set_tuning blocks_per_sm_cap (+ 100 * tiles per thread)")
    a = ap.parse_args()
    dev = torch.device("cuda", 0)
    t = DeviceTransform(0)
    t.quant.ctx.set_tuning(a.enc, 0, a.cap)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    out = {}
    for w, h in ((1920, 1080), (3840, 2160), (7680, 4320)):
        g = torch.Generator(device=dev).manual_seed(7)
        rgb = 0.005 * torch.pow(torch.tensor(2.0e6, device=dev), torch.rand((1, 3, h, w), generator=g, device=dev))
        planes = t.alloc_planes(1, w, h)
        dec = torch.empty_like(rgb)
        stats = t.alloc_stats(1)
        res = {}
        for name, fn in (("encode", lambda: t.encode(rgb, planes=planes, stats=stats)), ("decode", lambda: t.decode(planes, w, h, out=dec))):
            for _ in range(5):
                fn()
            ts = []
            for _ in range(a.iters):
                flush.fill_(1)
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record()
                fn()
                e1.record()
                e1.synchronize()
                ts.append(e0.elapsed_time(e1) * 1e3)
            ts.sort()
            floor = 15.0 * w * h / 6458.4e9 * 1e6
            res[name] = {"median_us": ts[len(ts) // 2], "best_us": ts[0], "hbm_floor_us": floor}
            print(f"{w}x{h} {name}: median {ts[len(ts)//2]:7.1f} us  best {ts[0]:7.1f} us  (15 B/px at the measured HBM peak: {floor:5.1f} us)", flush=True)
        out[f"{w}x{h}"] = res
    print(json.dumps(out))


if __name__ == "__main__":
    main()
