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

W, H, N = (int(sys.argv[1]), int(sys.argv[2]), 12) if len(sys.argv) > 2 else (3840, 2160, 12)
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
          for _ in range(6)]
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


for nb in (1, 2, 3, 4, 8):
    bands(nb)
    t_enc(); t_dec()
    e, d = t_enc(), t_dec()
    both, res = t_both()
    print(f"bands {nb:2d}: encode {e:6.3f} ms  decode {d:6.3f} ms  concurrent (independent loops) {both:6.3f} ms per frame pair "
          f"(enc {res['e']:.3f}, dec {res['d']:.3f})", flush=True)

if len(sys.argv) > 2:
    sys.exit(0)

# ---- the bench's pipeline: encoder thread -> queue -> decoder thread, two plane slots
import queue


def pipelined(nframes, nslots=2):
    free_q, full_q = queue.Queue(), queue.Queue()
    for i in range(nslots):
        free_q.put(i)
    te, td, we, wd = [], [], [], []

    def e():
        for i in range(nframes):
            t0 = time.perf_counter()
            s = free_q.get()
            t1 = time.perf_counter()
            enc.encode(frame, planes[s])
            t2 = time.perf_counter()
            full_q.put(s)
            we.append(t1 - t0); te.append(t2 - t1)
        full_q.put(None)

    def d():
        k = 0
        while True:
            t0 = time.perf_counter()
            s = full_q.get()
            if s is None:
                return
            t1 = time.perf_counter()
            dec.m_frame = outs[k & 1]
            dec.decode(planes[s], W, H)
            t2 = time.perf_counter()
            free_q.put(s)
            wd.append(t1 - t0); td.append(t2 - t1)
            k += 1
    t0 = time.perf_counter()
    a, b = threading.Thread(target=e), threading.Thread(target=d)
    a.start(); b.start(); a.join(); b.join()
    tot = (time.perf_counter() - t0) / nframes * 1e3
    ms = lambda v: 1e3 * sum(v[2:]) / max(1, len(v) - 2)
    return tot, ms(te), ms(td), ms(we), ms(wd)


for nb, ns in ((1, 2), (1, 3), (1, 4), (1, 6), (2, 2), (2, 4), (2, 6), (8, 4)):
    bands(nb)
    pipelined(6, ns)
    tot, te, td, we, wd = pipelined(24, ns)
    print(f"pipelined bands {nb} slots {ns}: {tot:6.3f} ms per frame | encode call {te:.3f} (waited {we:.3f}) | decode call {td:.3f} (waited {wd:.3f})", flush=True)
