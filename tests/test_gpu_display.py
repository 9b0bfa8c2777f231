"""GPU (-m gpu): display decode = the bit-exact decode path + the tail of the reference's player
shader (src/lumaplay_dequantizer.frag:141-156; scaling = preScaling / userScaling, lumaplay.cpp:406).  The tail is
floating-point pow in both the GLSL original and here, so the check is a numpy restatement applied to the ORACLE's
decoded frame with a tolerance of 1 LSB of the 8-bit output (stated by this test; observed: a handful of samples
at rounding boundaries)."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def tail(rgb, exposure, gamma, scaling, do_tmo, ldr_sim):
    v = rgb.astype(np.float32)
    with np.errstate(all="ignore"):
        if ldr_sim:
            v = np.float32(exposure) * np.maximum(np.float32(1), np.minimum(np.float32(256), np.floor(np.float32(256) * v / np.float32(scaling)))) / np.float32(256)
        else:
            v = v * np.float32(exposure) / np.float32(scaling)
        if do_tmo:
            t = np.power(v, np.float32(0.8))
            v = t / (t + np.float32(0.8) ** np.float32(0.8))
        v = np.power(v, np.float32(1.0 / gamma))
        v = np.where(np.isnan(v), np.float32(0), np.clip(v, 0, 1))
    return np.rint(v * np.float32(255)).astype(np.int32)


@pytest.mark.parametrize("cs,profile,w,h", [("LUV", 2, 640, 360), ("LUV", 3, 322, 202), ("YCBCR", 2, 320, 200), ("XYZ", 0, 320, 200),
                                            ("RGB", 1, 130, 66)])
@pytest.mark.parametrize("mode", [dict(), dict(do_tmo=True, exposure=2.0), dict(ldr_sim=True, gamma=1.8, user_scaling=4.0)])
def test_display_matches_restated_shader_tail(lumalib, po, cs, profile, w, h, mode):
    L = lumalib
    bits = 8 if profile < 2 else 11
    sc = 2.0
    o = po.Oracle().setQuantizer("PQ", bits, cs, 8)
    frame = po.noise_frame(w, h, seed=21) * np.float32(0.01)   # 5e-5 .. 100 cd/m2 so that the 8-bit output is not all white
    planes, _ = o.encode(frame.copy(), profile, sc)
    dec = L.LumaDecoder()
    dec.setParams(L.LumaDecoderParams(ptf=L.PTF_PQ, colorSpace=cs, ptfBitDepth=bits, colorBitDepth=8, profile=profile, preScaling=sc))
    dec.initialize()
    p = dict(exposure=1.0, gamma=2.2, user_scaling=1.0, do_tmo=False, ldr_sim=False)
    p.update(mode)
    got = dec.display(planes, w, h, **p)
    assert got.shape == (h, w, 4) and np.all(got[..., 3] == 255)
    # the player does not divide the frame by preScaling; it folds it into `scaling`
    lin = o.decode(planes, w, h, profile, 1.0)
    ref = tail(lin, p["exposure"], p["gamma"], sc / p["user_scaling"], p["do_tmo"], p["ldr_sim"])
    diff = np.abs(got[..., :3].astype(np.int32) - np.moveaxis(ref, 0, 2))
    assert diff.max() <= 1, f"max difference {diff.max()} LSB"
    assert (diff != 0).mean() < 0.02
    assert got[..., :3].std() > 5  # a real picture, not a constant


def test_display_device_batch(lumalib, po):
    import ctypes as C

    import torch
    from lumahdrv_b200._lib import DisplayParams, check
    from lumahdrv_b200.device import DeviceTransform

    w, h, n = 640, 360, 3
    t = DeviceTransform(0)
    rgb = torch.stack([torch.from_numpy(po.noise_frame(w, h, seed=50 + i) * np.float32(0.01)) for i in range(n)]).cuda()
    planes = t.encode(rgb)
    out = torch.zeros((n, h, w, 4), dtype=torch.uint8, device="cuda")
    pn, ptrs, strides, fstr = t._plane_args(planes)
    p = DisplayParams(1.0, 2.2, 1.0, 1, 0, 0)
    hnd = t.quant.ctx.handle
    check(t._lib.lumacu_display_dev(hnd, ptrs, strides, w, h, 2, 1.0, C.byref(p), out.data_ptr(), w * 4, n, fstr, 0, None), hnd,
          "lumacu_display_dev")
    t.quant.ctx.synchronize() if hasattr(t.quant.ctx, "synchronize") else torch.cuda.synchronize()
    torch.cuda.synchronize()
    lin = t.decode(planes, w, h).cpu().numpy()
    got = out.cpu().numpy()
    for i in range(n):
        ref = tail(lin[i], 1.0, 2.2, 1.0, True, False)
        assert np.abs(got[i, ..., :3].astype(np.int32) - np.moveaxis(ref, 0, 2)).max() <= 1


def gl_linear_restatement(planes, w, h, profile, lut, max_val_color, cs):
    """numpy restatement of how the PLAYER samples (lumaplay.cpp:258-259 GL_LINEAR + CLAMP_TO_EDGE on every texture,
    lumaplay.cpp:371 LUT texture of maxVal texels, src/lumaplay_dequantizer.frag:78-139) at 1:1 scale."""
    nb = 2 if profile > 1 else 1
    sub = profile in (0, 2)

    def codes(p, pw, ph):
        pl = planes[p][:ph, : pw * nb]
        return (pl.view("<u2") if nb == 2 else pl).astype(np.float32)

    cw, ch = ((w + 1) // 2, (h + 1) // 2) if sub else (w, h)
    c0 = codes(0, w, h)

    def bilinear(c):
        if not sub:
            return c
        u = np.float32(0.5) * np.arange(w, dtype=np.float32) - np.float32(0.25)      # texel coordinate of pixel centre x
        v = np.float32(0.5) * np.arange(h, dtype=np.float32) - np.float32(0.25)
        x0 = np.floor(u).astype(np.int32); fx = (u - np.floor(u)).astype(np.float32)
        y0 = np.floor(v).astype(np.int32); fy = (v - np.floor(v)).astype(np.float32)
        assert set(np.unique(fx)) <= {np.float32(0.25), np.float32(0.75)}          # the weights this test pins
        xa, xb = np.clip(x0, 0, cw - 1), np.clip(x0 + 1, 0, cw - 1)
        ya, yb = np.clip(y0, 0, ch - 1), np.clip(y0 + 1, 0, ch - 1)
        top = c[ya][:, xa] + fx[None, :] * (c[ya][:, xb] - c[ya][:, xa])
        bot = c[yb][:, xa] + fx[None, :] * (c[yb][:, xb] - c[yb][:, xa])
        return (top + fy[:, None] * (bot - top)).astype(np.float32)

    def lut_linear(code):
        max_val = lut.size - 1
        t = np.clip(code, 0, max_val).astype(np.float32) - np.float32(0.5)
        fl = np.floor(t); f = (t - fl).astype(np.float32)
        i0 = np.clip(fl.astype(np.int32), 0, max_val - 1); i1 = np.clip(fl.astype(np.int32) + 1, 0, max_val - 1)
        return (lut[i0] + f * (lut[i1] - lut[i0])).astype(np.float32)

    c1, c2 = bilinear(codes(1, cw, ch)), bilinear(codes(2, cw, ch))
    L = lut_linear(c0)
    mi = np.array([[3.240708, -1.537259, -0.498570], [-0.969257, 1.875995, 0.041555], [0.055636, -0.203996, 1.057069]], np.float32)
    with np.errstate(all="ignore"):
        if cs == "LUV":
            u = c1 / np.float32(max_val_color) * np.float32(255) / np.float32(410)
            v = c2 / np.float32(max_val_color) * np.float32(255) / np.float32(410)
            den = np.float32(6) * u - np.float32(16) * v + np.float32(12)
            x, y = np.float32(9) * u / den, np.float32(4) * v / den
            Y = np.clip(L, 1e-4, 1e8).astype(np.float32)
            X = np.clip(x / y * L, 1e-4, 1e8).astype(np.float32)
            Z = np.clip((np.float32(1) - x - y) / y * L, 1e-4, 1e8).astype(np.float32)
            return np.stack([mi[r, 0] * X + mi[r, 1] * Y + mi[r, 2] * Z for r in range(3)]).astype(np.float32)
        if cs == "RGB":
            return np.stack([L, lut_linear(c1), lut_linear(c2)])
        a, b = lut_linear(c1), lut_linear(c2)
        return np.stack([mi[r, 0] * L + mi[r, 1] * a + mi[r, 2] * b for r in range(3)]).astype(np.float32)


@pytest.mark.parametrize("cs,profile,w,h", [("LUV", 2, 640, 360), ("LUV", 0, 322, 202), ("RGB", 2, 320, 200), ("XYZ", 3, 130, 66)])
def test_display_linear_filter_matches_the_players_sampling(lumalib, po, cs, profile, w, h):
    L = lumalib
    bits = 8 if profile < 2 else 11
    o = po.Oracle().setQuantizer("PQ", bits, cs, 8)
    frame = po.noise_frame(w, h, seed=77) * np.float32(0.01)
    planes, _ = o.encode(frame.copy(), profile, 1.0)
    dec = L.LumaDecoder()
    dec.setParams(L.LumaDecoderParams(ptf=L.PTF_PQ, colorSpace=cs, ptfBitDepth=bits, colorBitDepth=8, profile=profile))
    dec.initialize()
    got = dec.display(planes, w, h, exposure=1.5, gamma=2.2, do_tmo=True, linear=True)
    lin = gl_linear_restatement(planes, w, h, profile, np.asarray(o.getMapping(), dtype=np.float32), 255, cs)
    ref = tail(lin, 1.5, 2.2, 1.0, True, False)
    diff = np.abs(got[..., :3].astype(np.int32) - np.moveaxis(ref, 0, 2))
    assert diff.max() <= 1 and (diff != 0).mean() < 0.02
    if profile in (0, 2):  # it is NOT the nearest-neighbour result
        near = dec.display(planes, w, h, exposure=1.5, gamma=2.2, do_tmo=True, linear=False)
        assert (near != got).mean() > 0.2


def test_display_linear_filter_pins_the_interpolation_weights(lumalib):
    """Crafted planes, RGB colour space (every plane goes through the LUT), 8-bit 4:2:0, LUT lut[i] = 2 i, gamma 1,
    exposure 1/255: the 8-bit output IS the sampled value (always an integer here, no rounding ties), so the 1/4 : 3/4
    chroma weights, the edge clamp and the half-code LUT fetch (mean of lut[c-1], lut[c] = 2 c - 1) can be read off."""
    L = lumalib
    w, h = 16, 8
    dec = L.LumaDecoder()
    dec.setParams(L.LumaDecoderParams(ptf=L.PTF_LINEAR, colorSpace=L.CS_RGB, ptfBitDepth=8, colorBitDepth=8, profile=0, maxLum=510.0))
    dec.initialize()
    dec.m_quant.setMapping(2.0 * np.arange(256, dtype=np.float32))
    planes = L.alloc_planes(w, h, 0)
    planes[0][:, :w] = 51                                    # luma code 51 everywhere
    planes[1][:, : w // 2] = np.array([0, 102] * (w // 4))   # chroma columns alternate 0, 102
    planes[2][: h // 2, : w // 2] = (np.arange(h // 2)[:, None] % 2) * 102   # chroma rows alternate 0, 102
    got = dec.display(planes, w, h, exposure=1.0 / 255.0, gamma=1.0, linear=True)
    assert np.all(got[..., 0] == 101)                        # mean of lut[50], lut[51] = 2 * 51 - 1

    def sampled(c, n):
        """texel coordinate code - 1/2 on lut[i] = 2 i: 2 (code - 1/2), clamped at the first texel"""
        out = []
        for x in range(n):
            k = x // 2
            v = 0.25 * c[max(k - 1, 0)] + 0.75 * c[k] if x % 2 == 0 else 0.75 * c[k] + 0.25 * c[min(k + 1, len(c) - 1)]
            out.append(max(2.0 * (v - 0.5), 0.0))
        return np.array(out)
    # plane 1 along x: pixel 2k blends c[k-1], c[k] as 1/4 : 3/4, pixel 2k+1 blends c[k], c[k+1] as 3/4 : 1/4, edges clamped
    want = sampled(np.array([0, 102] * (w // 4), dtype=np.float64), w)
    assert np.all(want == np.rint(want)) and np.array_equal(got[0, :, 1], want.astype(np.uint8)), (got[0, :, 1], want)
    want = sampled(np.array([0, 102] * (h // 4), dtype=np.float64), h)
    assert np.array_equal(got[:, 3, 2], want.astype(np.uint8)), (got[:, 3, 2], want)
