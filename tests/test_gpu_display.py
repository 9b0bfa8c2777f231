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
    p = DisplayParams(1.0, 2.2, 1.0, 1, 0)
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
