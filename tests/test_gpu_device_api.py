"""GPU (-m gpu): device-resident batch API, full-size properties, multi-frame launches."""
import numpy as np
import pytest

from conftest import bits_equal

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def torch_cuda():
    import torch
    assert torch.cuda.is_available()
    return torch


def test_batched_frames_equal_single_frames_and_oracle(lumalib, po, torch_cuda):
    torch = torch_cuda
    from lumahdrv_b200.device import DeviceTransform
    w, h, n = 192, 64, 5
    frames = np.stack([po.noise_frame(w, h, seed=100 + i) for i in range(n)])
    t = DeviceTransform(0)
    rgb = torch.from_numpy(frames).cuda()
    stats = t.alloc_stats(n)
    planes = t.encode(rgb, stats=stats)
    out = t.decode(planes, w, h)
    torch.cuda.synchronize()
    st = t.stats_to_numpy(stats)
    o = po.Oracle().setQuantizer("PQ", 11, "LUV", 8)
    for i in range(n):
        f = frames[i].copy()
        ref_planes, _ = o.encode(f, 2, 1.0)
        for p in range(3):
            assert np.array_equal(planes[p][i].cpu().numpy(), ref_planes[p]), (i, p)
        assert bits_equal(out[i].cpu().numpy(), o.decode(ref_planes, w, h, 2, 1.0))
        assert st["sum"][i] == pytest.approx(float(f[0].astype(np.float64).sum()), rel=1e-6)
        assert st["max"][i] == f[0].max() and st["min"][i] == f[0].min()
    # second launch reuses the self-cleaning stats workspace
    stats2 = t.alloc_stats(n)
    t.encode(rgb, planes=planes, stats=stats2)
    torch.cuda.synchronize()
    assert np.array_equal(t.stats_to_numpy(stats2)["sum"], st["sum"])
    assert t.launch_count == 3


def test_write_back_matches_reference_side_effect(lumalib, po, torch_cuda):
    torch = torch_cuda
    from lumahdrv_b200.device import DeviceTransform
    frame = po.noise_frame(128, 32, seed=9)
    t = DeviceTransform(0)
    rgb = torch.from_numpy(frame[None]).cuda()
    wb = torch.empty_like(rgb)
    t.encode(rgb, write_back=wb)
    f = frame.copy()
    po.Oracle().setQuantizer("PQ", 11, "LUV", 8).transformColorSpace(f, True, 1.0)
    assert bits_equal(wb[0].cpu().numpy(), f)
    assert bits_equal(rgb[0].cpu().numpy(), frame)  # input untouched unless aliased


@pytest.mark.parametrize("w,h", [(3840, 2160), (7680, 4320)])
def test_full_size_properties(lumalib, po, torch_cuda, w, h):
    """Size-independent properties at BASELINE sizes: (1) encode and decode are deterministic, (2) a frame
    processed alone equals the same frame processed inside a batch, (3) bands of rows (whole 4:2:0 blocks)
    are bit-identical to the oracle, (4) the fp64 plane-0 sum equals the sum of the oracle's band sums."""
    torch = torch_cuda
    from lumahdrv_b200.device import DeviceTransform
    t = DeviceTransform(0)
    g = torch.Generator(device="cuda").manual_seed(1234)
    u = torch.rand((2, 3, h, w), generator=g, device="cuda", dtype=torch.float32)
    rgb = (0.005 * torch.pow(torch.tensor(2.0e6, device="cuda"), u)).contiguous()
    del u
    stats = t.alloc_stats(2)
    planes = t.encode(rgb, stats=stats)
    planes2 = t.encode(rgb)
    single = t.encode(rgb[1:2].contiguous())
    for a, b, c in zip(planes, planes2, single):
        assert torch.equal(a, b)
        assert torch.equal(a[1:2], c)
    dec = t.decode(planes, w, h)
    assert torch.equal(dec, t.decode(planes2, w, h))
    assert torch.equal(dec[1:2], t.decode(single, w, h))
    o = po.Oracle().setQuantizer("PQ", 11, "LUV", 8)
    for f, start in [(0, 0), (0, h // 2 - 8), (1, h - 16), (1, 2 * (h // 6))]:
        rows = slice(start, start + 16)
        band = rgb[f, :, rows, :].cpu().numpy().copy()
        ref_planes, _ = o.encode(band, 2, 1.0)
        assert np.array_equal(planes[0][f, rows, : 2 * w].cpu().numpy(), ref_planes[0][:, : 2 * w])
        crow = slice(rows.start // 2, rows.stop // 2)
        assert np.array_equal(planes[1][f, crow, :w].cpu().numpy(), ref_planes[1][:, :w])
        assert np.array_equal(planes[2][f, crow, :w].cpu().numpy(), ref_planes[2][:, :w])
        assert bits_equal(dec[f, :, rows, :].cpu().numpy(), o.decode(ref_planes, w, 16, 2, 1.0))
    st = t.stats_to_numpy(stats)
    # Y of the whole frame from the elementwise colour transform of the oracle is too slow at 8K; use torch fp64
    m = torch.tensor([0.212656, 0.715158, 0.072186], device="cuda", dtype=torch.float64)
    for f in range(2):
        y = (rgb[f].double() * m[:, None, None]).sum(0)
        assert st["sum"][f] == pytest.approx(float(y.sum()), rel=1e-5)
        assert st["max"][f] == pytest.approx(float(y.max()), rel=1e-5)


def test_luma_codes_are_fixed_points(lumalib, torch_cuda):
    """quantize(dequantize(code)) == code for every code of every shipped transfer function (exact property)."""
    import lumahdrv_b200 as L
    for ptf, bits in [("PQ", 8), ("PQ", 10), ("PQ", 11), ("PQ", 12), ("LOG", 11), ("LOG", 12), ("PSI", 11),
                      ("JND_HDRVDP", 12), ("LINEAR", 11), ("PQ", 16)]:
        q = L.LumaQuantizer().setQuantizer(ptf, bits, "LUV", 8)
        codes = np.arange(1 << bits, dtype=np.float32)
        assert np.array_equal(q.quantize(q.dequantize(codes, 0), 0), codes), (ptf, bits)


@pytest.mark.parametrize("w,h", [(1280, 720), (1920, 1080)])
def test_every_tuning_variant_produces_identical_bits(lumalib, po, torch_cuda, w, h):
    """lumacu_set_tuning: pipelining / occupancy / search variants of the tuned kernels (register prefetch, tensor-map
    TMA staging, bucket search, block sizing) are performance knobs only."""
    import torch
    from lumahdrv_b200.device import DeviceTransform

    t = DeviceTransform(0)
    n = 3
    rgb = torch.stack([torch.from_numpy(po.noise_frame(w, h, seed=300 + i)) for i in range(n)]).cuda()
    rgb[1, :, 3, 5] = float("nan")
    ctx = t.quant.ctx
    ctx.set_tuning(0, 0, 0)
    ref_planes = [p.clone() for p in t.encode(rgb)]
    ref_out = t.decode(ref_planes, w, h).clone()
    o = po.Oracle().setQuantizer("PQ", 11, "LUV", 8)
    cpu_planes, _ = o.encode(rgb[0].cpu().numpy().copy(), 2, 1.0)
    for a, b, (pw, ph) in zip(ref_planes, cpu_planes, po.plane_dims(w, h, 2)):
        assert np.array_equal(a[0].cpu().numpy()[:ph, :pw * 2], b[:ph, :pw * 2])
    for enc_v, dec_v, cap in [(3, 3, 0), (5, 5, 0), (84, 13, 0), (1004, 14, 0), (1003, 15, 0), (1012, 0, 0), (1013, 0, 0),
                              (0, 0, 2), (0, 0, 3200), (0, 0, 101)]:
        ctx.set_tuning(enc_v, dec_v, cap)
        planes = t.encode(rgb)
        for a, b in zip(planes, ref_planes):
            assert torch.equal(a, b), f"encode variant {enc_v} cap {cap}"
        out = t.decode(ref_planes, w, h)
        assert torch.equal(out.view(torch.int32), ref_out.view(torch.int32)), f"decode variant {dec_v} cap {cap}"
    ctx.set_tuning(0, 0, 0)


def test_plain_c_example_round_trips(torch_cuda):
    """examples/roundtrip.c (strict C99 against include/lumacu.h) encodes and decodes a frame on the GPU."""
    import shutil
    import subprocess
    from pathlib import Path

    root = Path(__file__).resolve().parent.parent
    exe = root / "examples" / "roundtrip"
    if not exe.exists():
        gcc = shutil.which("gcc")
        if not gcc:
            pytest.skip("examples/roundtrip not built and no gcc")
        subprocess.run([gcc, "-std=c99", "-O2", f"-I{root / 'include'}", str(root / "examples" / "roundtrip.c"),
                        f"-L{root / 'lumahdrv_b200'}", "-llumacu", "-Wl,-rpath,$ORIGIN/../lumahdrv_b200", "-lm", "-o", str(exe)], check=True)
    r = subprocess.run([str(exe)], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stdout + r.stderr
    assert "worst relative round-trip error" in r.stdout
