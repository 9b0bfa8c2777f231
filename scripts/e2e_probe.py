#!/usr/bin/env python
"""Where does the host-pointer path spend its time?  encode alone, decode alone, both on two threads; by band count."""
import sys
import threading
import time
from pathlib import Path

import numpy as np
import torch

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
import lumahdrv_b200 as L  # noqa: E402
from lumahdrv_b200._lib import check  # noqa: E402

W, H, N = 3840, 2160, 12
enc = L.LumaEncoder(0)
enc.initialize(None, W, H)
dec = L.LumaDecoder(0)
dec.setParams(L.LumaDecoderParams(ptf=L.PTF_PQ, colorSpace=L.CS_LUV))
dec.initialize()
rng = np.random.default_rng(1)
h_in = torch.empty((3, H, W), dtype=torch.float32).pin_memory()
h_in.copy_(torch.from_numpy((0.005 * np.power(2.0e6, rng.random((3, H, W), dtype=np.float32))).astype(np.float32)))
frame = h_in.numpy()
strides = L.vpx_strides(W, 2)
planes = [[torch.empty((ph, s), dtype=torch.uint8).pin_memory().numpy() for (pw, ph), s in zip(L.plane_dims(W, H, 2), strides)]
          for _ in range(2)]
outs = [torch.empty((3, H, W), dtype=torch.float32).pin_memory().numpy() for _ in range(2)]


def bands(n):
    for o in (enc, dec):
        hnd = o.m_quant.ctx.handle
        check(o.m_quant._lib.lumacu_set_host_bands(hnd, n), hnd, "bands")


def t_enc():
    t0 = time.perf_counter()
    for i in range(N):
        enc.encode(frame, planes[i & 1])
    return (time.perf_counter() - t0) / N * 1e3


def t_dec():
    t0 = time.perf_counter()
    for i in range(N):
        dec.m_frame = outs[i & 1]
        dec.decode(planes[i & 1], W, H)
    return (time.perf_counter() - t0) / N * 1e3


def t_both():
    res = {}

    def a():
        res["e"] = t_enc()

    def b():
        res["d"] = t_dec()
    t0 = time.perf_counter()
    ta, tb = threading.Thread(target=a), threading.Thread(target=b)
    ta.start(); tb.start(); ta.join(); tb.join()
    return (time.perf_counter() - t0) / N * 1e3, res


for nb in (1, 2, 4, 8, 16):
    bands(nb)
    t_enc(); t_dec()
    e, d = t_enc(), t_dec()
    both, res = t_both()
    print(f"bands {nb:2d}: encode {e:6.3f} ms  decode {d:6.3f} ms  concurrent (independent loops) {both:6.3f} ms per frame pair "
          f"(enc {res['e']:.3f}, dec {res['d']:.3f})", flush=True)
