#!/usr/bin/env python
"""PCIe ceiling of the box: pinned H2D, D2H and both at once (two streams), 100 MB blocks."""
import time

import torch

n = 100 * 1024 * 1024
h_a = torch.empty(n, dtype=torch.uint8).pin_memory()
h_b = torch.empty(n, dtype=torch.uint8).pin_memory()
d_a = torch.empty(n, dtype=torch.uint8, device="cuda")
d_b = torch.empty(n, dtype=torch.uint8, device="cuda")
s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()


def run(fn, reps=20):
    fn()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(reps):
        fn()
    torch.cuda.synchronize()
    return (time.perf_counter() - t0) / reps


def h2d():
    with torch.cuda.stream(s1):
        d_a.copy_(h_a, non_blocking=True)


def d2h():
    with torch.cuda.stream(s2):
        h_b.copy_(d_b, non_blocking=True)


def both():
    h2d()
    d2h()


t = run(h2d)
print(f"H2D  {n / t / 1e9:6.1f} GB/s")
t = run(d2h)
print(f"D2H  {n / t / 1e9:6.1f} GB/s")
t = run(both)
print(f"both {2 * n / t / 1e9:6.1f} GB/s total ({n / t / 1e9:.1f} per direction)")
# many small copies (8 bands x 3 planes): per-copy overhead
m = n // 24
def small():
    with torch.cuda.stream(s1):
        for i in range(24):
            d_a[i * m:(i + 1) * m].copy_(h_a[i * m:(i + 1) * m], non_blocking=True)
t = run(small)
print(f"H2D in 24 pieces {n / t / 1e9:6.1f} GB/s")
