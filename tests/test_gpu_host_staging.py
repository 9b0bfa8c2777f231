"""GPU (-m gpu): the host-pointer entry points (lumacu_encode / lumacu_decode) cut frames into row bands that
overlap H2D copy, kernel and D2H copy on three streams.  Whatever the band count, the
planes, the decoded floats and the frame statistics must equal the single-band result and the oracle."""
import threading

import numpy as np
import pytest

from conftest import bits_equal

pytestmark = pytest.mark.gpu


def _set_bands(L, obj, n):
    from lumahdrv_b200._lib import check
    hnd = obj.m_quant.ctx.handle
    check(obj.m_quant._lib.lumacu_set_host_bands(hnd, n), hnd, "lumacu_set_host_bands")


@pytest.mark.parametrize("w,h,profile,cs", [(1920, 1080, 2, "LUV"), (1280, 722, 3, "LUV"), (1026, 1030, 0, "YCBCR"),
                                            (2048, 1024, 2, "XYZ")])
def test_banded_equals_single_band_and_oracle(lumalib, po, w, h, profile, cs):
    L = lumalib
    cbits = 10 if cs == "YCBCR" else 8
    bits = 8 if profile < 2 else 11
    enc = L.LumaEncoder()
    enc.setParams(L.LumaEncoderParams(ptf="PQ", ptfBitDepth=bits, colorSpace=cs, colorBitDepth=min(cbits, 8 if profile < 2 else 12),
                                      profile=profile, bitDepth=8 if profile < 2 else 12))
    enc.initialize(None, w, h)
    cb = enc.getParams().colorBitDepth
    dec = L.LumaDecoder()
    dec.setParams(L.LumaDecoderParams(ptf=L.PTF_PQ, colorSpace=enc.getParams().colorSpace, ptfBitDepth=bits, colorBitDepth=cb,
                                      profile=profile))
    dec.initialize()
    o = po.Oracle().setQuantizer("PQ", bits, cs, cb)
    frame = po.noise_frame(w, h, seed=99)
    frame[:, 5, 7] = np.nan
    frame[0, h - 1, w - 1] = -3.0
    ref_planes, ref_avg = o.encode(frame.copy(), profile, 1.0)
    ref_out = o.decode(ref_planes, w, h, profile, 1.0)
    nbytes = 2 if profile > 1 else 1
    results = []
    for bands in (1, 3, 8, 0):
        _set_bands(L, enc, bands)
        _set_bands(L, dec, bands)
        planes = enc.encode(frame.copy(), L.alloc_planes(w, h, profile))
        for p, (a, b, (pw, ph)) in enumerate(zip(planes, ref_planes, po.plane_dims(w, h, profile))):
            assert np.array_equal(a[:ph, :pw * nbytes], b[:ph, :pw * nbytes]), f"bands={bands}: plane {p} differs from the oracle"
        out = dec.decode(planes, w, h).copy()
        assert bits_equal(out, ref_out), f"bands={bands}: decoded floats differ from the oracle"
        results.append(enc.last_stats)
    # statistics: NaN pixel is ignored by max/min, poisons nothing else; band sums add up
    for st in results[1:]:
        assert st["max"] == results[0]["max"] and st["min"] == results[0]["min"]
        if np.isfinite(results[0]["sum"]):
            assert abs(st["sum"] - results[0]["sum"]) <= 1e-9 * abs(results[0]["sum"])


def test_encoder_and_decoder_objects_run_concurrently_on_two_threads(lumalib, po):
    """bench.py's end-to-end arm: encode(i+1) on one thread while decode(i) runs on another."""
    L = lumalib
    w, h, n = 1920, 1080, 6
    enc = L.LumaEncoder()
    enc.initialize(None, w, h)
    dec = L.LumaDecoder()
    dec.setParams(L.LumaDecoderParams(ptf=L.PTF_PQ, colorSpace=L.CS_LUV))
    dec.initialize()
    o = po.Oracle().setQuantizer("PQ", 11, "LUV", 8)
    frames = [po.noise_frame(w, h, seed=1000 + i) for i in range(n)]
    planes = [L.alloc_planes(w, h, 2) for _ in range(n)]
    outs = [None] * n
    done = [threading.Event() for _ in range(n)]
    errors = []

    def encoder():
        try:
            for i in range(n):
                enc.encode(frames[i], planes[i])
                done[i].set()
        except Exception as e:  # noqa: BLE001
            errors.append(e)
            for d in done:
                d.set()

    def decoder():
        try:
            for i in range(n):
                done[i].wait()
                outs[i] = dec.decode(planes[i], w, h).copy()
        except Exception as e:  # noqa: BLE001
            errors.append(e)

    ts = [threading.Thread(target=encoder), threading.Thread(target=decoder)]
    [t.start() for t in ts]
    [t.join() for t in ts]
    assert not errors, errors
    for i in range(n):
        ref_planes, _ = o.encode(frames[i].copy(), 2, 1.0)
        for a, b, (pw, ph) in zip(planes[i], ref_planes, po.plane_dims(w, h, 2)):
            assert np.array_equal(a[:ph, :pw * 2], b[:ph, :pw * 2])
        assert bits_equal(outs[i], o.decode(ref_planes, w, h, 2, 1.0))


def test_host_register_roundtrip(lumalib):
    import ctypes as C

    lib = lumalib.lib()
    buf = np.zeros(1 << 20, dtype=np.uint8)
    assert lib.lumacu_host_register(C.c_void_p(buf.ctypes.data), buf.nbytes) == 0
    assert lib.lumacu_host_register(C.c_void_p(buf.ctypes.data), buf.nbytes) == 0  # twice is fine
    assert lib.lumacu_host_unregister(C.c_void_p(buf.ctypes.data)) == 0


def test_broadcast_quantizer_and_async_pair_through_the_c_abi(lumalib, po):
    """lumacu_broadcast_quantizer: contexts that never saw lumacu_set_quantizer encode/decode with the root's tables
    (an arbitrary LUT, so nothing could have been rebuilt locally).  lumacu_encode_async / lumacu_decode_async +
    lumacu_wait_input / lumacu_wait give the same bytes as the blocking calls, and a second call on a busy context
    first completes the one in flight."""
    import ctypes as C

    import torch
    L = lumalib
    lib = L.lib()
    n_dev = torch.cuda.device_count()
    ctxs = [L.Context(i % n_dev) for i in range(3)]
    rng = np.random.default_rng(5)
    lut = np.sort(np.unique(np.power(10.0, rng.uniform(-3, 4, 4096)).astype(np.float32)))[:1024].copy()
    h0 = ctxs[0].handle
    L._lib.check(lib.lumacu_set_quantizer(h0, lut.ctypes.data, lut.size, 255, L.CS_LUV, 1e4), h0, "set")
    arr = (C.c_void_p * 3)(*[c.handle for c in ctxs])
    assert lib.lumacu_broadcast_quantizer(arr, 3, 0) == 0
    assert lib.lumacu_broadcast_quantizer(arr, 3, 5) == 1 and lib.lumacu_broadcast_quantizer(None, 3, 0) == 1
    # the receivers hold the root's host-side copy too
    got = np.zeros(1024, np.float32)
    n, mvc, cs, lm = C.c_uint32(), C.c_uint32(), C.c_int(), C.c_float()
    assert lib.lumacu_get_quantizer(ctxs[2].handle, got.ctypes.data, got.size, C.byref(n), C.byref(mvc), C.byref(cs), C.byref(lm)) == 0
    assert n.value == 1024 and mvc.value == 255 and cs.value == L.CS_LUV and lm.value == 1e4 and np.array_equal(got, lut)

    o = po.Oracle().setQuantizer("LINEAR", 10, "LUV", 8)
    o.setMapping(lut)
    w, h = 512, 128
    frames = [po.noise_frame(w, h, seed=40 + i) for i in range(3)]
    refs = [o.encode(f.copy(), 2, 1.0)[0] for f in frames]
    strides = L.vpx_strides(w, 2)
    pinned = []

    def pin(shape, dtype):
        t = torch.empty(shape, dtype=dtype).pin_memory()
        pinned.append(t)
        return t.numpy()

    ins = [pin((3, h, w), torch.float32) for _ in range(3)]
    outs = [pin((3, h, w), torch.float32) for _ in range(3)]
    planes = [[pin((ph, st), torch.uint8) for (pw, ph), st in zip(L.plane_dims(w, h, 2), strides)] for _ in range(3)]
    stats = [L._lib.FrameStats() for _ in range(3)]
    for i, c in enumerate(ctxs):  # queue on every context first, collect afterwards
        ins[i][...] = frames[i]
        ptrs, st = L.luma._plane_args(planes[i])
        assert lib.lumacu_encode_async(c.handle, ins[i].ctypes.data, w, h, 2, 1.0, ptrs, st, 0, C.byref(stats[i])) == 0
        assert lib.lumacu_pending(c.handle) == 1
    for i, c in enumerate(ctxs):
        assert lib.lumacu_wait_input(c.handle) == 0
        ins[i][...] = -1.0  # the input may be reused now
        assert lib.lumacu_wait(c.handle) == 0 and lib.lumacu_pending(c.handle) == 0
        for a, b, (pw, ph) in zip(planes[i], refs[i], L.plane_dims(w, h, 2)):
            assert np.array_equal(a[:ph, :pw * 2], b[:ph, :pw * 2]), f"context {i}"
        fc = frames[i].copy()
        o.transformColorSpace(fc, True, 1.0)
        assert stats[i].max == float(fc[0].max()) and stats[i].sum == pytest.approx(float(fc[0].astype(np.float64).sum()), rel=1e-6)
    for i, c in enumerate(ctxs):
        ptrs, st = L.luma._plane_args(planes[i])
        assert lib.lumacu_decode_async(c.handle, ptrs, st, w, h, 2, 1.0, outs[i].ctypes.data) == 0
    # a blocking call on a busy context completes the pending one first (context 0: decode again, blocking)
    ptrs, st = L.luma._plane_args(planes[0])
    assert lib.lumacu_decode(ctxs[0].handle, ptrs, st, w, h, 2, 1.0, outs[0].ctypes.data) == 0
    for i, c in enumerate(ctxs):
        assert lib.lumacu_wait(c.handle) == 0
        assert bits_equal(outs[i], o.decode(refs[i], w, h, 2, 1.0)), f"context {i}"
    for c in ctxs:
        c.close()


def test_pageable_host_buffers_are_reported_once(lumalib):
    """Pageable caller memory is legal but slow (the driver stages the copies): the first call that sees it says so on
    stderr, once per context; page-locked buffers (lumacu_host_alloc / torch pin_memory) stay silent."""
    import os
    import subprocess
    import sys
    from pathlib import Path

    root = Path(__file__).resolve().parent.parent
    prog = r'''
import sys, numpy as np, torch
sys.path.insert(0, %r)
import lumahdrv_b200 as L
w, h = 256, 64
enc = L.LumaEncoder(); enc.initialize(None, w, h)
frame = np.full((3, h, w), 50.0, np.float32)
if sys.argv[1] == "pinned":
    keep = [torch.empty((3, h, w), dtype=torch.float32).pin_memory()]
    frame = keep[0].numpy(); frame[...] = 50.0
    planes = []
    for (pw, ph), st in zip(L.plane_dims(w, h, 2), L.vpx_strides(w, 2)):
        keep.append(torch.zeros((ph, st), dtype=torch.uint8).pin_memory()); planes.append(keep[-1].numpy())
else:
    planes = L.alloc_planes(w, h, 2)
for _ in range(3):
    enc.encode(frame, planes)
print("done")
''' % str(root)
    env = {k: v for k, v in os.environ.items() if k != "LUMACU_QUIET"}
    for mode, expect in (("pageable", 2), ("pinned", 0)):  # frame + plane buffer, each reported once... per context: once in total
        r = subprocess.run([sys.executable, "-c", prog, mode], capture_output=True, text=True, timeout=300, env=env)
        assert r.returncode == 0 and "done" in r.stdout, r.stderr[-2000:]
        n = r.stderr.count("pageable host memory")
        assert (n == 1) if expect else (n == 0), (mode, r.stderr[-500:])
    quiet = subprocess.run([sys.executable, "-c", prog, "pageable"], capture_output=True, text=True, timeout=300,
                           env=dict(env, LUMACU_QUIET="1"))
    assert quiet.returncode == 0 and "pageable host memory" not in quiet.stderr


@pytest.mark.parametrize("w,h,profile,cs,channels", [(1920, 1080, 2, "LUV", 7), (1280, 722, 3, "YCBCR", 15), (514, 258, 0, "LUV", 2)])
def test_half_rgba_host_entry_points(lumalib, po, w, h, profile, cs, channels):
    """lumacu_encode_half_rgba / lumacu_decode_half_rgba: the EXR pixel loops (src/exr_interface.cpp:73-143, :157-187) run on
    the device inside the banded pipeline, 8 B/px over the bus.  Planes equal those of the reference's sequence (expand
    the half pixels on the host, then encode -- checked against our float entry point AND the CPU oracle); decoded pixels
    equal decode followed by the float -> half rounding of Imf's half (= numpy's float16 cast), alpha 0."""
    L = lumalib
    bits = 8 if profile < 2 else 11
    cb = 8 if profile < 2 else 10
    enc = L.LumaEncoder()
    enc.setParams(L.LumaEncoderParams(ptf="PQ", ptfBitDepth=bits, colorSpace=cs, colorBitDepth=cb, profile=profile,
                                      bitDepth=8 if profile < 2 else 12))
    enc.initialize(None, w, h)
    dec = L.LumaDecoder()
    dec.setParams(L.LumaDecoderParams(ptf=L.PTF_PQ, colorSpace=enc.getParams().colorSpace, ptfBitDepth=bits, colorBitDepth=cb,
                                      profile=profile))
    dec.initialize()
    rng = np.random.default_rng(w + profile)
    rgba = np.empty((h, w, 4), dtype=np.float16)
    rgba[..., :3] = np.moveaxis(po.noise_frame(w, h, seed=31), 0, -1).astype(np.float16)
    rgba[..., 3] = 1.0
    k = rng.integers(0, h * w, 3000)
    rgba.reshape(-1, 4)[k, rng.integers(0, 3, k.size)] = rng.choice(
        np.array([0.0, -0.0, 6e-8, 6.1e-5, 65504.0, np.inf, -np.inf, np.nan, -1.0], np.float16), k.size)
    f = rgba.astype(np.float32)
    if channels in (7, 15):
        frame = np.ascontiguousarray(np.stack([f[..., 0], f[..., 1], f[..., 2]]))
    else:
        frame = np.ascontiguousarray(np.stack([f[..., {1: 0, 2: 1, 4: 2}[channels]]] * 3))
    nbytes = 2 if profile > 1 else 1
    o = po.Oracle().setQuantizer("PQ", bits, cs, cb)
    ref_planes, _ = o.encode(frame.copy(), profile, 1.0)
    for bands in (1, 5, 0):
        _set_bands(L, enc, bands)
        _set_bands(L, dec, bands)
        want = [p.copy() for p in enc.encode(frame.copy(), L.alloc_planes(w, h, profile))]
        want_stats = dict(enc.last_stats)
        got = enc.encode_half_rgba(rgba, channels, L.alloc_planes(w, h, profile))
        for p, (a, b, c, (pw, ph)) in enumerate(zip(got, want, ref_planes, po.plane_dims(w, h, profile))):
            assert np.array_equal(a, b), f"bands={bands}: plane {p} differs from the float entry point"
            assert np.array_equal(a[:ph, :pw * nbytes], c[:ph, :pw * nbytes]), f"bands={bands}: plane {p} differs from the oracle"
        assert enc.last_stats["max"] == want_stats["max"] and enc.last_stats["min"] == want_stats["min"]
        out = dec.decode(got, w, h).copy()
        half = dec.decode_half_rgba(got, w, h)
        with np.errstate(over="ignore", invalid="ignore"):
            ref_half = np.stack([out[0], out[1], out[2], np.zeros_like(out[0])], axis=-1).astype(np.float16)
        assert np.array_equal(half.view(np.uint16), ref_half.view(np.uint16)), f"bands={bands}: half pixels differ"


def test_half_rgba_rejects_luminance_only_files(lumalib):
    L = lumalib
    enc = L.LumaEncoder()
    enc.setParams(L.LumaEncoderParams(profile=2, bitDepth=12))
    enc.initialize(None, 64, 32)
    with pytest.raises(L.LumaException, match="luminance only"):
        enc.encode_half_rgba(np.zeros((32, 64, 4), np.float16), channels=16)


@pytest.mark.parametrize("bits,cs,cbits,lmax", [(10, "YCBCR", 10, 1000.0), (16, "LUV", 8, 1e4), (12, "LUV", 12, 1e4)])
def test_broadcast_carries_every_table_flavour(lumalib, po, bits, cs, cbits, lmax):
    """lumacu_broadcast_quantizer with the quantizers whose device state is more than LUT + thresholds: CS_YCBCR (PQ tables,
    v-keyed search table and the verified Lmax reciprocal are rebuilt on the receiving device), a 16-bit LUT (direct table
    in global memory, decode LUT read in place) and a 12-bit one (64-bit two-threshold table).  The receiver -- on the
    next device when there is one -- must produce the oracle's planes and floats and run the tuned kernels."""
    import ctypes as C

    import torch
    L = lumalib
    lib = L.lib()
    n_dev = torch.cuda.device_count()
    root, recv = L.Context(0), L.Context(1 % n_dev)
    lut = L.build_lut("PQ", bits, lmax, 0.005)
    cs_id = {"LUV": L.CS_LUV, "YCBCR": L.CS_YCBCR}[cs]
    L._lib.check(lib.lumacu_set_quantizer(root.handle, lut.ctypes.data, lut.size, (1 << cbits) - 1, cs_id, lmax), root.handle, "set")
    arr = (C.c_void_p * 2)(root.handle, recv.handle)
    assert lib.lumacu_broadcast_quantizer(arr, 2, 0) == 0
    o = po.Oracle().setQuantizer("PQ", bits, cs, cbits, lmax, 0.005)
    w, h = 1024, 256
    frame = po.noise_frame(w, h, seed=bits)
    ref_planes, _ = o.encode(frame.copy(), 2, 1.0)
    ref_out = o.decode(ref_planes, w, h, 2, 1.0)
    planes = L.alloc_planes(w, h, 2)
    ptrs, st = L.luma._plane_args(planes)
    stats = L._lib.FrameStats()
    f = frame.copy()
    L._lib.check(lib.lumacu_encode(recv.handle, f.ctypes.data, w, h, 2, 1.0, ptrs, st, 0, C.byref(stats)), recv.handle, "encode")
    assert lib.lumacu_last_kernel_path(recv.handle) == 1
    for a, b, (pw, ph) in zip(planes, ref_planes, po.plane_dims(w, h, 2)):
        assert np.array_equal(a[:ph, :pw * 2], b[:ph, :pw * 2])
    out = np.empty((3, h, w), np.float32)
    L._lib.check(lib.lumacu_decode(recv.handle, ptrs, st, w, h, 2, 1.0, out.ctypes.data), recv.handle, "decode")
    assert lib.lumacu_last_kernel_path(recv.handle) == 1
    assert bits_equal(out, ref_out)
