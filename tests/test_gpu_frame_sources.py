"""GPU (-m gpu): frame sources on the device against the oracle's restatement of
ExrInterface::testFrame (src/exr_interface.cpp:50-70; its 256x256 output is pinned by the survey's hash b2082ac1)
and against numpy for the exact half -> float conversion of ExrInterface::readFrame's pixel loop."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def dt(lumalib):
    import torch
    from lumahdrv_b200.device import DeviceTransform
    assert torch.cuda.is_available()
    return DeviceTransform(0)


@pytest.mark.parametrize("w,h", [(256, 256), (1280, 720), (1920, 1080), (250, 103), (7, 5), (3840, 2160)])
def test_test_frame_bit_exact(dt, po, golden, w, h):
    got = dt.test_frame(w, h).cpu().numpy()
    ref = po.test_frame(w, h)
    assert got.shape == ref.shape == (3, h, w)
    assert np.array_equal(got.view(np.uint32), ref.view(np.uint32))
    if (w, h) == (256, 256):
        assert "%08x" % po.fnv1a32(got) == golden["frames"]["cfg1_256_pq_luv"]["input"] == "b2082ac1"


def test_test_frame_feeds_encode_without_a_host_copy(dt, po, golden):
    """testFrame on the device -> encode on the device: the survey's plane hashes of the 1080p config."""
    import torch
    w, h = 1920, 1080
    frame = dt.test_frame(w, h)
    planes = dt.encode(frame[None])
    got = [p[0].cpu().numpy() for p in planes]
    assert ["%08x" % v for v in po.plane_hashes(got, w, h, 2)] == golden["frames"]["cfg2_1080p_pq_luv"]["planes"]
    torch.cuda.synchronize()


@pytest.mark.parametrize("channels", [7, 15, 1, 2, 4])
def test_half_rgba_to_frame(dt, channels):
    import torch
    h, w = 123, 250
    rng = np.random.default_rng(5)
    bits = rng.integers(0, 1 << 16, size=(h, w, 4), dtype=np.uint16)  # every half pattern incl. inf / nan / subnormals
    px = bits.view(np.float16)
    got = dt.half_rgba_to_frame(torch.from_numpy(px).cuda(), channels).cpu().numpy()
    f = px.astype(np.float32)
    if channels in (7, 15):
        ref = np.stack([f[..., 0], f[..., 1], f[..., 2]])
    else:
        c = {1: 0, 2: 1, 4: 2}[channels]
        ref = np.stack([f[..., c]] * 3)
    nan = np.isnan(ref)
    assert np.array_equal(np.isnan(got), nan)
    assert np.array_equal(got.view(np.uint32)[~nan], ref.view(np.uint32)[~nan])


def test_frame_to_half_rgba(dt):
    """ExrInterface::writeFrame's pixel loop (src/exr_interface.cpp:167-175): p.r/g/b = float -> half, p.a = 0.  Imf's
    half(float) rounds to nearest even, overflows to infinity and keeps a NaN's sign and top payload bits; numpy's
    float32 -> float16 cast is the same function (OpenEXR is not in the reference tree: parity unpinned against it)."""
    import torch
    h, w = 257, 510
    rng = np.random.default_rng(11)
    bits = rng.integers(0, 1 << 32, size=(3, h, w), dtype=np.uint64).astype(np.uint32)  # any float: NaNs, infs, subnormals
    f = bits.view(np.float32).copy()
    flat = f.reshape(-1)
    # the half range densely: ties (exactly between two halfs), the overflow threshold 65520, half subnormals, signed zeros
    allh = np.arange(65536, dtype=np.uint16).view(np.float16).astype(np.float32)
    fin = allh[np.isfinite(allh)]
    with np.errstate(over="ignore"):  # the neighbour above the largest finite half is infinity
        nxt = np.nextafter(fin.astype(np.float16), np.float16(np.inf)).astype(np.float32)
    ties = ((fin.astype(np.float64) + nxt.astype(np.float64)) / 2).astype(np.float32)
    special = np.concatenate([allh, ties, np.nextafter(ties, np.float32(0)), np.nextafter(ties, np.float32(np.inf)),
                              np.array([65504.0, 65519.99, 65520.0, 65536.0, 1e30, -65520.0, 2.0 ** -25, 2.0 ** -24, 1.5 * 2.0 ** -25,
                                        0.0, -0.0, np.inf, -np.inf], np.float32),
                              np.array([0x7f800001, 0xff800001, 0x7fc00000, 0xffc12345, 0x7f801fff, 0x7f802000], np.uint32).view(np.float32)])
    flat[: special.size] = special
    got = dt.frame_to_half_rgba(torch.from_numpy(f).cuda()).cpu().numpy().view(np.uint16)
    with np.errstate(over="ignore", invalid="ignore"):
        ref = np.stack([f[0], f[1], f[2], np.zeros_like(f[0])], axis=-1).astype(np.float16).view(np.uint16)
    assert got.shape == (h, w, 4)
    assert np.array_equal(got, ref), f"{(got != ref).sum()} half words differ"
    # read back through the source kernel: half -> float is exact, so the round trip is the half-rounded frame
    back = dt.half_rgba_to_frame(torch.from_numpy(got.view(np.float16)).cuda(), 7).cpu().numpy()
    want = ref[..., :3].view(np.float16).astype(np.float32).transpose(2, 0, 1)
    nan = np.isnan(want)
    assert np.array_equal(np.isnan(back), nan) and np.array_equal(back.view(np.uint32)[~nan], want.view(np.uint32)[~nan])


def test_half_rgba_rejects_luminance_only(dt, lumalib):
    import torch
    with pytest.raises(lumalib.LumaException, match="luminance only"):
        dt.half_rgba_to_frame(torch.zeros((4, 4, 4), dtype=torch.float16, device="cuda"), channels=16)


def test_pfs_xyz_channels_to_frame_and_back(dt, po):
    """PfsInterface::readFrame / writeFrame colour steps (src/pfs_interface.cpp:84, :140).  pfstools is not in the
    reference tree (parity unpinned against libpfs); its D65 matrices are the constants of the reference's own
    xyz2rgbMat / rgb2xyzMat, so XYZ -> RGB is pinned against the reference's own XYZ inverse transform with sc = 1
    (src/luma_quantizer.cpp:378-395: same matrix, same association, no clamp) and RGB -> XYZ against a float32
    restatement of m0*a + m1*b + m2*c."""
    import torch
    h, w = 270, 482
    rng = np.random.default_rng(8)
    xyz = np.power(np.float32(10.0), rng.uniform(-4, 5, size=(3, h, w)).astype(np.float32)).astype(np.float32)
    xyz[:, 0, :6] = np.array([0.0, -1.5, np.inf, np.nan, 1e-42, 3e38], dtype=np.float32)
    d = [torch.from_numpy(xyz[c].copy()).cuda() for c in range(3)]
    got = dt.pfs_xyz_to_frame(*d).cpu().numpy()
    ref = xyz.copy()
    assert po.Oracle().setQuantizer("PQ", 11, "XYZ", 8).transformColorSpace(ref, False, 1.0)
    nan = np.isnan(ref)
    assert np.array_equal(np.isnan(got), nan) and np.array_equal(got.view(np.uint32)[~nan], ref.view(np.uint32)[~nan])
    # and back: RGB -> XYZ, unclamped
    m = np.array([[0.412424, 0.357579, 0.180464], [0.212656, 0.715158, 0.072186], [0.019332, 0.119193, 0.950444]], dtype=np.float32)
    rgb = np.abs(xyz)
    with np.errstate(all="ignore"):
        want = np.stack([(m[r, 0] * rgb[0] + m[r, 1] * rgb[1]) + m[r, 2] * rgb[2] for r in range(3)]).astype(np.float32)
    back = torch.stack(dt.frame_to_pfs_xyz(torch.from_numpy(rgb).cuda())).cpu().numpy()
    nan = np.isnan(want)
    assert np.array_equal(np.isnan(back), nan) and np.array_equal(back.view(np.uint32)[~nan], want.view(np.uint32)[~nan])
