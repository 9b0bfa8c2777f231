#!/usr/bin/env python
"""e2e pipeline with several encoder / decoder objects (one host thread each): does keeping >1 call in flight per
direction close the gap between 2.9 ms per frame and the 2.6 ms PCIe duplex floor?"""
import queue
import sys
import threading
import time
from pathlib import Path

import numpy as np
import torch

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
import lumahdrv_b200 as L  # noqa: E402

W, H = 3840, 2160


def mk(n_enc, n_dec):
    encs, decs = [], []
    for _ in range(n_enc):
        e = L.LumaEncoder(0)
        e.initialize(None, W, H)
        e.m_quant.ctx.set_host_bands(1)
        encs.append(e)
    for _ in range(n_dec):
        d = L.LumaDecoder(0)
        d.setParams(L.LumaDecoderParams(ptf=L.PTF_PQ, colorSpace=L.CS_LUV))
        d.initialize()
        d.m_quant.ctx.set_host_bands(1)
        decs.append(d)
    return encs, decs


rng = np.random.default_rng(1)
h_in = torch.empty((3, H, W), dtype=torch.float32).pin_memory()
h_in.copy_(torch.from_numpy((0.005 * np.power(2.0e6, rng.random((3, H, W), dtype=np.float32))).astype(np.float32)))
frame = h_in.numpy()
strides = L.vpx_strides(W, 2)
NSLOT = 8
planes = [[torch.empty((ph, s), dtype=torch.uint8).pin_memory().numpy() for (pw, ph), s in zip(L.plane_dims(W, H, 2), strides)]
          for _ in range(NSLOT)]


def run(n_enc, n_dec, nframes, nslots):
    encs, decs = mk(n_enc, n_dec)
    outs = [torch.empty((3, H, W), dtype=torch.float32).pin_memory().numpy() for _ in range(n_dec)]
    for e in encs:
        e.encode(frame, planes[0])
    for i, d in enumerate(decs):
        d.m_frame = outs[i]
        d.decode(planes[0], W, H)
    free_q, full_q = queue.Queue(), queue.Queue()
    for i in range(nslots):
        free_q.put(i)
    todo = queue.Queue()
    for i in range(nframes):
        todo.put(i)

    def enc_thread(e):
        while True:
            try:
                todo.get_nowait()
            except queue.Empty:
                return
            s = free_q.get()
            e.encode(frame, planes[s])
            full_q.put(s)

    def dec_thread(d):
        while True:
            s = full_q.get()
            if s is None:
                return
            d.decode(planes[s], W, H)
            free_q.put(s)

    te = [threading.Thread(target=enc_thread, args=(e,)) for e in encs]
    td = [threading.Thread(target=dec_thread, args=(d,)) for d in decs]
    t0 = time.perf_counter()
    [t.start() for t in te + td]
    [t.join() for t in te]
    for _ in decs:
        full_q.put(None)
    [t.join() for t in td]
    return (time.perf_counter() - t0) / nframes * 1e3


for n_enc, n_dec, ns in ((1, 1, 4), (2, 2, 6), (2, 2, 8), (2, 1, 6), (1, 2, 6), (3, 3, 8)):
    run(n_enc, n_dec, 8, ns)
    ms = run(n_enc, n_dec, 40, ns)
    print(f"{n_enc} encoder(s) + {n_dec} decoder(s), {ns} plane slots: {ms:6.3f} ms per frame  ({W*H/ms/1e3:7.1f} Mpx/s)", flush=True)
