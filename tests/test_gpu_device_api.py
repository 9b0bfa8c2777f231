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


@pytest.mark.parametrize("seed", range(int(__import__("os").environ.get("LUMA_FUZZ_SEEDS", "4"))))
def test_random_batches_on_the_device_api(lumalib, po, torch_cuda, seed):
    """Seeded sweep through the device-pointer entry points (multi-frame launches): random colour space, profile, PTF, bit
    depths, range, preScaling, frame count, frame size, plane pitches, with and without statistics -- tuned kernels ==
    generic kernels on every byte (encode) / bit (decode, incl. random out-of-range codes), and one frame of each draw
    against the CPU oracle.  LUMA_FUZZ_SEEDS=N widens the sweep."""
    torch = torch_cuda
    from lumahdrv_b200.device import DeviceTransform
    from lumahdrv_b200 import vpx_strides

    rng = np.random.default_rng(5000 + seed)
    for _ in range(12):
        cs = ("LUV", "RGB", "YCBCR", "XYZ")[rng.integers(0, 4)]
        profile = int(rng.integers(0, 4))
        ptf = ("PQ", "LOG", "PQ", "LINEAR")[rng.integers(0, 4)]
        bits, cbits = (8, int(rng.integers(2, 9))) if profile < 2 else (int(rng.integers(6, 17)), int(rng.integers(4, 15)))
        lmax = float(rng.choice([1000.0, 4000.0, 1e4]))
        sc = float(rng.choice([1.0, 1.0, 20.0]))
        sub = profile in (0, 2)
        n = int(rng.integers(1, 5))
        w = int(rng.integers(1, 300)) * 2 if rng.random() < 0.6 else int(rng.integers(1, 9)) * 128
        h = int(rng.integers(1, 120)) * 2
        what = f"cs={cs} profile={profile} ptf={ptf} bits={bits}/{cbits} lmax={lmax} sc={sc} n={n} {w}x{h}"
        t = DeviceTransform(0, ptf=ptf, ptfBitDepth=bits, colorSpace=cs, colorBitDepth=cbits, maxLum=lmax, minLum=0.005,
                            profile=profile, preScaling=sc)
        ctx = t.quant.ctx
        frames = [np.ascontiguousarray(po.noise_frame(w, h, seed=int(rng.integers(1 << 30))) / np.float32(sc)) for _ in range(n)]
        k = rng.integers(0, frames[0].size, 64)
        frames[0].reshape(-1)[k] = rng.choice(np.array([0.0, -1.0, np.nan, np.inf, 1e-45, 3e38, 1e-4, 1e8], np.float32), k.size)
        rgb = torch.from_numpy(np.stack(frames)).cuda()
        strides = None
        if rng.random() < 0.5:
            strides = [int(s) + int(rng.integers(0, 5)) * 16 for s in vpx_strides(w, profile)]
        use_stats = rng.random() < 0.5
        nbytes = 2 if profile > 1 else 1
        results = []
        for path in (1, 0):
            ctx.set_kernel_path(path)
            planes = t.alloc_planes(n, w, h, strides)
            stats = t.alloc_stats(n) if use_stats else None
            t.encode(rgb, planes=planes, stats=stats)
            results.append((planes, t.stats_to_numpy(stats) if use_stats else None))
        for p, (a, b) in enumerate(zip(results[0][0], results[1][0])):
            assert torch.equal(a, b), f"{what}: encode plane {p}: tuned and generic kernels differ in {(a != b).sum().item()} bytes"
        if use_stats:
            for key in ("max", "min"):
                assert np.array_equal(results[0][1][key], results[1][1][key], equal_nan=True), f"{what}: stats {key}"
        o = po.Oracle().setQuantizer(ptf, bits, cs, cbits, lmax, 0.005)
        f = int(rng.integers(0, n))
        ref_planes, _ = o.encode(frames[f].copy(), profile, sc)
        for p, (a, b, (pw, ph)) in enumerate(zip(results[1][0], ref_planes, po.plane_dims(w, h, profile))):
            assert np.array_equal(a[f].cpu().numpy()[:ph, :pw * nbytes], b[:ph, :pw * nbytes]), f"{what}: frame {f} plane {p} vs the oracle"
        # decode: the encoder's planes, then planes of random code words (any 16-bit / 8-bit pattern)
        for trial in range(2):
            planes = results[1][0]
            if trial == 1:
                planes = [torch.from_numpy(rng.integers(0, 256, size=tuple(p.shape), dtype=np.uint8)).cuda() for p in planes]
            outs = []
            for path in (1, 0):
                ctx.set_kernel_path(path)
                outs.append(t.decode(planes, w, h).clone())
            assert torch.equal(outs[0].view(torch.int32), outs[1].view(torch.int32)), f"{what}: decode trial {trial}: tuned != generic"
            host = [p[f].cpu().numpy() for p in planes]
            ref = o.decode(host, w, h, profile, sc)
            got = outs[1][f].cpu().numpy()
            if cs == "YCBCR":
                from conftest import max_ulp
                assert max_ulp(got, ref) <= 1, f"{what}: decode trial {trial} vs the oracle"
            else:
                assert bits_equal(got, ref), f"{what}: decode trial {trial} vs the oracle"
        ctx.set_kernel_path(0)


def test_frame_larger_than_32_bit_offsets(lumalib, po, torch_cuda):
    """Maximum sizes: a 32768 x 40000 frame (1.31 Gpx, 15.7 GB of f32; byte offsets inside a plane pass 2^32, so the tuned
    kernels' 32-bit offsets do not apply and the generic kernels index with 64 bits).  Bands of rows at the start, either
    side of the 2^32-byte mark and at the very end against the oracle; a frame of more than 2^31 - 1 pixels is refused."""
    torch = torch_cuda
    from lumahdrv_b200.device import DeviceTransform
    free, _ = torch.cuda.mem_get_info()
    if free < 70 << 30:
        pytest.skip("needs 70 GB of free device memory")
    w, h = 32768, 40000
    t = DeviceTransform(0)
    g = torch.Generator(device="cuda").manual_seed(77)
    rgb = torch.empty((1, 3, h, w), dtype=torch.float32, device="cuda")
    for c in range(3):  # plane by plane: no 15 GB temporaries
        for y0 in range(0, h, 4000):
            u = torch.rand((4000, w), generator=g, device="cuda", dtype=torch.float32)
            rgb[0, c, y0:y0 + 4000] = 0.005 * torch.pow(torch.tensor(2.0e6, device="cuda"), u)
    del u
    stats = t.alloc_stats(1)
    planes = t.encode(rgb, stats=stats)
    assert t.quant.ctx.last_kernel_path == 0
    dec = t.decode(planes, w, h)
    o = po.Oracle().setQuantizer("PQ", 11, "LUV", 8)
    row_4g = (1 << 32) // (4 * w)  # the row whose first float sits 2^32 bytes into its plane
    for start in (0, row_4g - 8, row_4g, h - 16):
        rows = slice(start, start + 16)
        band = rgb[0, :, rows, :].cpu().numpy().copy()
        ref_planes, _ = o.encode(band, 2, 1.0)
        assert np.array_equal(planes[0][0, rows, : 2 * w].cpu().numpy(), ref_planes[0][:, : 2 * w]), f"rows {start}.."
        crow = slice(rows.start // 2, rows.stop // 2)
        assert np.array_equal(planes[1][0, crow, :w].cpu().numpy(), ref_planes[1][:, :w]), f"rows {start}.."
        assert np.array_equal(planes[2][0, crow, :w].cpu().numpy(), ref_planes[2][:, :w]), f"rows {start}.."
        assert bits_equal(dec[0, :, rows, :].cpu().numpy(), o.decode(ref_planes, w, 16, 2, 1.0)), f"rows {start}.."
    st = t.stats_to_numpy(stats)
    assert 0.005 * 0.2 < st["min"][0] < 0.02 and 5e3 < st["max"][0] <= 1.0001e4 and np.isfinite(st["sum"][0])
    del rgb, dec, planes
    torch.cuda.empty_cache()
    # 46342 x 46342 = 2 147 580 964 pixels > 2^31 - 1: refused before anything is touched
    from lumahdrv_b200._lib import lib
    import ctypes as C
    dummy = torch.zeros(64, dtype=torch.uint8, device="cuda")
    ptrs = (C.c_void_p * 3)(dummy.data_ptr(), dummy.data_ptr(), dummy.data_ptr())
    strides = (C.c_int32 * 3)(46342 * 2, 46342, 46342)
    fstr = (C.c_size_t * 3)(0, 0, 0)
    rc = lib().lumacu_encode_dev(t.quant.ctx.handle, dummy.data_ptr(), None, 46342, 46342, 2, 1.0, ptrs, strides, 1, 0, fstr, None, None)
    assert rc != 0 and b"too large" in lib().lumacu_last_error(t.quant.ctx.handle)


@pytest.mark.parametrize("w,h,n", [(3840, 2160, 3), (7680, 4320, 2)])
def test_whole_frames_of_a_batch_equal_the_oracle(lumalib, po, torch_cuda, w, h, n):
    """Whole 4K / 8K noise frames through the multi-frame launch (the grid geometry bench.py times), every byte of every
    plane and every decoded float of every frame against the CPU oracle, plus the per-frame statistics."""
    torch = torch_cuda
    from lumahdrv_b200.device import DeviceTransform
    t = DeviceTransform(0)
    frames = [po.noise_frame(w, h, seed=900 + i) for i in range(n)]
    rgb = torch.from_numpy(np.stack(frames)).cuda()
    stats = t.alloc_stats(n)
    planes = t.encode(rgb, stats=stats)
    assert t.quant.ctx.last_kernel_path == 1
    dec = t.decode(planes, w, h)
    st = t.stats_to_numpy(stats)
    o = po.Oracle().setQuantizer("PQ", 11, "LUV", 8)
    for f in range(n):
        fc = frames[f].copy()
        ref_planes, _ = o.encode(fc, 2, 1.0)
        for p, (a, b, (pw, ph)) in enumerate(zip(planes, ref_planes, po.plane_dims(w, h, 2))):
            assert np.array_equal(a[f].cpu().numpy()[:ph, :pw * 2], b[:ph, :pw * 2]), f"frame {f} plane {p}"
        assert bits_equal(dec[f].cpu().numpy(), o.decode(ref_planes, w, h, 2, 1.0)), f"frame {f} decoded floats"
        y = fc[0].astype(np.float64)
        assert st["max"][f] == np.float32(y.max()) and st["min"][f] == np.float32(y.min())
        assert st["sum"][f] == pytest.approx(float(y.sum()), rel=1e-9)


def test_bench_line_carries_parity_and_configs(torch_cuda):
    """bench.py on this GPU (2 frames per step, a few steps): ONE JSON line with the contract's keys, `parity` with zero
    mismatches from the CPU checker on the benchmark's own data (device-resident and end-to-end arms), and the cfg2..cfg5
    block, each configuration with its own zero-mismatch parity."""
    import json
    import subprocess
    import sys
    from pathlib import Path

    root = Path(__file__).resolve().parent.parent
    r = subprocess.run([sys.executable, str(root / "bench.py"), "--steps", "3", "--warmup", "3", "--frames", "2", "--e2e-frames", "2",
                        "--config-steps", "2", "--no-cpu-baseline", "--no-sustained"], capture_output=True, text=True, timeout=550)
    assert r.returncode == 0, r.stderr[-3000:]
    lines = [ln for ln in r.stdout.strip().splitlines() if ln.startswith("{")]
    assert len(lines) == 1
    d = json.loads(lines[0])
    for k in ("metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling", "vs_baseline",
              "dtype", "data", "config", "roofline", "e2e", "gpu_launches", "clocks", "parity", "configs"):
        assert k in d, k
    assert d["gpu_launches"] == 6 and d["value"] > 0 and d["roofline"]["bound"] == "hbm" and 0 < d["roofline"]["frac"] < 1.2
    for par in (d["parity"], d["e2e"]["parity"]):
        assert par["plane_mismatch_bytes"] == 0 and par["max_ulp"] == 0 and par["frames_checked"] == 1 and par["ranks_failed"] == 0
        assert par["pixels_checked"] == 3840 * 2160 and par["stats_max_mismatches"] == 0 and par["stats_sum_max_rel_err"] < 1e-6
    assert d["e2e"]["h2d_bytes_per_step"] > 0 and d["e2e"]["d2h_bytes_per_step"] > 0
    assert set(d["configs"]) == {"cfg2", "cfg3", "cfg4", "cfg5"}
    for name, c in d["configs"].items():
        assert "error" not in c, (name, c)
        assert c["value"] > 0 and c["parity"]["plane_mismatch_bytes"] == 0 and c["parity"]["max_ulp"] == 0, name
    assert d["configs"]["cfg5"]["parity"]["stats_max_mismatches"] == 0
    assert d["configs"]["cfg4"]["frames_total_per_step"] >= 64


def test_one_context_on_two_streams(lumalib, po, torch_cuda):
    """The device entry points of ONE context launched on two caller streams at once: planes, decoded floats and the
    per-frame statistics (whose reduction workspace is kept per stream) equal the single-stream results."""
    torch = torch_cuda
    from lumahdrv_b200.device import DeviceTransform
    t = DeviceTransform(0)
    w, h, n = 1920, 1080, 6
    rgb = [torch.from_numpy(np.stack([po.noise_frame(w, h, seed=700 + 10 * k + i) for i in range(n)])).cuda() for k in range(2)]
    ref_planes, ref_stats = [], []
    for k in range(2):
        st = t.alloc_stats(n)
        ref_planes.append([p.clone() for p in t.encode(rgb[k], stats=st)])
        torch.cuda.synchronize()
        ref_stats.append(t.stats_to_numpy(st).copy())
    streams = [torch.cuda.Stream(), torch.cuda.Stream()]
    planes = [t.alloc_planes(n, w, h) for _ in range(2)]
    stats = [t.alloc_stats(n) for _ in range(2)]
    outs = [torch.empty_like(rgb[0]) for _ in range(2)]
    torch.cuda.synchronize()
    for _ in range(20):  # interleave launches of the two streams
        for k in range(2):
            with torch.cuda.stream(streams[k]):
                t.encode(rgb[k], planes=planes[k], stats=stats[k])
                t.decode(planes[k], w, h, out=outs[k])
    torch.cuda.synchronize()
    for k in range(2):
        for a, b in zip(planes[k], ref_planes[k]):
            assert torch.equal(a, b), f"stream {k}: planes"
        got = t.stats_to_numpy(stats[k])
        assert np.array_equal(got["sum"], ref_stats[k]["sum"]) and np.array_equal(got["max"], ref_stats[k]["max"])
        assert np.array_equal(got["min"], ref_stats[k]["min"]), f"stream {k}: statistics"
        assert torch.equal(outs[k].view(torch.int32), t.decode(ref_planes[k], w, h).view(torch.int32))


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
    for enc_v, dec_v, cap in [(4, 4, 0), (0, 24, 0), (0, 64, 0), (27, 0, 0), (6, 0, 0), (7, 0, 0), (26, 0, 0), (67, 0, 0), (64, 0, 0), (86, 0, 0),
                              (87, 0, 0), (1067, 0, 0), (1004, 0, 0), (3, 3, 0), (5, 5, 0), (84, 13, 0), (1004, 14, 0), (1003, 15, 0), (1012, 0, 0), (1013, 0, 0), (2067, 0, 0), (2004, 0, 0), (2067, 0, 400),
                              (0, 0, 2), (0, 0, 3200), (0, 0, 101)]:
        ctx.set_tuning(enc_v, dec_v, cap)
        planes = t.encode(rgb)
        for a, b in zip(planes, ref_planes):
            assert torch.equal(a, b), f"encode variant {enc_v} cap {cap}"
        out = t.decode(ref_planes, w, h)
        assert torch.equal(out.view(torch.int32), ref_out.view(torch.int32)), f"decode variant {dec_v} cap {cap}"
    ctx.set_tuning(0, 0, 0)


def _boundary_frame(w, h, seed, max_c=255.0, rel_span=2e-5):
    """2x2-constant blocks whose chroma code sits within +-rel_span (relative) of a rounding boundary: for every block
    pick R, B at random and solve (float64 bisection on G) 410/255 * maxC * u'(R,G,B) + 0.5 = n (1 + d), d uniform in
    +-rel_span, u' = 4X / (X + 15Y + 3Z); half of the blocks do the same for v' = 9Y / (X + 15Y + 3Z)."""
    rng = np.random.default_rng(seed)
    bh, bw = h // 2, w // 2
    m = np.array([[0.412424, 0.357579, 0.180464], [0.212656, 0.715158, 0.072186], [0.019332, 0.119193, 0.950444]])
    R = np.power(10.0, rng.uniform(-2, 4, (bh, bw)))
    B = np.power(10.0, rng.uniform(-2, 4, (bh, bw)))
    use_v = rng.random((bh, bw)) < 0.5

    def t_of(G):
        X = m[0, 0] * R + m[0, 1] * G + m[0, 2] * B
        Y = m[1, 0] * R + m[1, 1] * G + m[1, 2] * B
        Z = m[2, 0] * R + m[2, 1] * G + m[2, 2] * B
        D = X + 15 * Y + 3 * Z
        return max_c * (410.0 / 255.0) * np.where(use_v, 9 * Y, 4 * X) / D + 0.5

    lo, hi = np.full((bh, bw), 1e-3), np.full((bh, bw), 1e5)
    t_lo, t_hi = t_of(lo), t_of(hi)
    a, b = np.minimum(t_lo, t_hi), np.maximum(t_lo, t_hi)
    n = np.floor(a + rng.random((bh, bw)) * (b - a))
    n = np.clip(n, np.ceil(a) + 0, np.floor(b) - 0)
    target = n * (1.0 + rng.uniform(-rel_span, rel_span, (bh, bw)))
    inc = t_hi > t_lo
    for _ in range(80):
        mid = np.sqrt(lo * hi)
        below = (t_of(mid) < target) == inc
        lo = np.where(below, mid, lo)
        hi = np.where(below, hi, mid)
    G = np.sqrt(lo * hi)
    blocks = np.stack([R, G, B]).astype(np.float32)
    return np.ascontiguousarray(np.repeat(np.repeat(blocks, 2, axis=1), 2, axis=2))


@pytest.mark.parametrize("cbits,sc", [(8, 1.0), (10, 1.0), (12, 1.0), (8, 3.5)])
def test_screened_chroma_equals_exact_chain(lumalib, po, torch_cuda, cbits, sc):
    """The screened-chroma encode kernel (Lu'v' 4:2:0 default; luma_fast.cuh FASTC) against the exact-chain tuned kernel
    (variant 4, itself pinned against the oracle everywhere else) on 100+ Mpixel of content chosen to stress the screen:
    chroma values straddling the code boundaries by relative distances around the acceptance threshold (2^-18),
    near-grey, saturated primaries, very dark, huge / negative / NaN / infinite samples, plus plain noise -- and a
    whole-frame check of one boundary frame against the CPU oracle."""
    import torch
    from lumahdrv_b200.device import DeviceTransform

    w, h = 1920, 1080
    t = DeviceTransform(0, colorBitDepth=cbits, preScaling=sc)
    ctx = t.quant.ctx
    max_c = float((1 << cbits) - 1)
    rng = np.random.default_rng(99)
    frames = [_boundary_frame(w, h, 1, max_c, 2e-5), _boundary_frame(w, h, 2, max_c, 6e-6), _boundary_frame(w, h, 3, max_c, 2e-7)]
    grey = po.noise_frame(w, h, seed=5)
    grey[1] = grey[0] * (1 + rng.uniform(-1e-3, 1e-3, (h, w)).astype(np.float32))
    grey[2] = grey[0] * (1 + rng.uniform(-1e-3, 1e-3, (h, w)).astype(np.float32))
    frames.append(grey)
    prim = po.noise_frame(w, h, seed=6)
    prim[rng.integers(0, 3, (h, w))[None].repeat(3, 0) == np.arange(3)[:, None, None]] *= np.float32(1e-6)  # one channel ~ 0
    frames.append(prim)
    frames.append(po.noise_frame(w, h, seed=7) * np.float32(1e-5))                     # everything below the 1e-4 clamp
    frames.append((po.noise_frame(w, h, seed=8) * np.float32(3e4)).astype(np.float32))  # up to 3e8: above 9e7 and the 1e8 clamp
    wild = po.noise_frame(w, h, seed=9)
    k = rng.integers(0, wild.size, 20000)
    wild.reshape(-1)[k] = rng.choice(np.array([-1.0, -1e-3, 0.0, np.nan, np.inf, -np.inf, 9.0e7, 9.1e7, 1e-45, 3e38], np.float32), k.size)
    frames.append(wild)
    frames += [po.noise_frame(w, h, seed=20 + i) for i in range(4)]
    frames = [np.ascontiguousarray(f / np.float32(sc)) if sc != 1.0 else f for f in frames]
    rgb = torch.from_numpy(np.stack(frames)).cuda()
    ctx.set_tuning(4)
    exact = [p.clone() for p in t.encode(rgb)]
    assert ctx.last_kernel_path == 1
    for tune in (0, 67, 27, 6, 7, 26, 86, 87, 1067):  # default (= screened), the screened variants (queued redo, L2 prefetch, TMA staging), screened + bucket/threshold luma search
        ctx.set_tuning(tune)
        got = t.encode(rgb)
        assert ctx.last_kernel_path == 1
        for p, (a, b) in enumerate(zip(got, exact)):
            if not torch.equal(a, b):
                bad = (a != b).nonzero()
                raise AssertionError(f"tuning {tune}: plane {p} differs in {bad.shape[0]} bytes, first at {bad[0].tolist()}")
    ctx.set_tuning(0)
    o = po.Oracle().setQuantizer("PQ", 11, "LUV", cbits)
    for f in (1, 7):
        cpu_planes, _ = o.encode(frames[f].copy(), 2, sc)
        for a, b, (pw, ph) in zip(exact, cpu_planes, po.plane_dims(w, h, 2)):
            assert np.array_equal(a[f].cpu().numpy()[:ph, :pw * 2], b[:ph, :pw * 2])


@pytest.mark.parametrize("bits,cbits,profile", [(12, 8, 2), (12, 12, 3), (13, 8, 2), (14, 10, 3), (16, 8, 2), (16, 16, 3)])
def test_wide_luts_run_the_tuned_table_kernels(lumalib, po, torch_cuda, bits, cbits, profile):
    """12-bit LUTs search a 64-bit two-threshold table in shared memory (screened kernel) or the bucket walk; 13-16-bit
    LUTs (`lumaenc -pb 16`) a one-threshold direct table that stays in global memory (luma_fast.cuh WALK -3 / -4).
    Every LUT entry, every midpoint between neighbours and the floats either side of both are in the frames; planes
    equal the generic kernel's (binary search over the thresholds) and the CPU oracle's, and the tuned kernel ran."""
    import torch
    from lumahdrv_b200.device import DeviceTransform

    w, h = 1024, 512
    t = DeviceTransform(0, ptf="PQ", ptfBitDepth=bits, colorSpace="LUV", colorBitDepth=cbits, profile=profile)
    ctx = t.quant.ctx
    lut = t.quant.getMapping().astype(np.float32)
    rng = np.random.default_rng(bits * 100 + cbits)
    mid = (0.5 * (lut[:-1].astype(np.float64) + lut[1:])).astype(np.float32)
    edge = np.concatenate([lut, mid, np.nextafter(mid, np.float32(0)), np.nextafter(mid, np.float32(np.inf)),
                           np.nextafter(lut, np.float32(0)), np.nextafter(lut, np.float32(np.inf))]).astype(np.float32)
    frames = []
    for i in range(3):
        f = po.noise_frame(w, h, seed=900 + bits + i)
        flat = f.reshape(3, -1)
        # grey pixels (R = G = B = L gives Y within an ulp or two of L) plus pixels whose Y is near a decision point
        pick = rng.permutation(edge)[: flat.shape[1] // 2]
        pos = rng.choice(flat.shape[1], size=pick.size, replace=False)
        flat[:, pos] = pick[None, :]
        frames.append(f)
    rgb = torch.from_numpy(np.stack(frames)).cuda()
    ctx.set_kernel_path(1)
    generic = [p.clone() for p in t.encode(rgb)]
    assert ctx.last_kernel_path == 0
    ctx.set_kernel_path(0)
    for tune in (0, 4, 67, 1004):
        ctx.set_tuning(tune)
        got = t.encode(rgb)
        # 1000 + v asks for the bucket + threshold walk, which the tuned kernels do not have for LUTs this dense
        tuned = 0 if (tune >= 1000 and bits >= 13) else 1
        assert ctx.last_kernel_path == tuned, f"tuning {tune}: kernel path {ctx.last_kernel_path}"
        for p, (a, b) in enumerate(zip(got, generic)):
            assert torch.equal(a, b), f"tuning {tune}: plane {p} differs in {(a != b).sum().item()} bytes"
    ctx.set_tuning(0)
    o = po.Oracle().setQuantizer("PQ", bits, "LUV", cbits)
    cpu_planes, _ = o.encode(frames[0].copy(), profile, 1.0)
    for a, b, (pw, ph) in zip(generic, cpu_planes, po.plane_dims(w, h, profile)):
        assert np.array_equal(a[0].cpu().numpy()[:ph, :pw * 2], b[:ph, :pw * 2])
    out_tuned = t.decode(generic, w, h).clone()
    # 14-16-bit LUTs do not fit in shared memory: the tuned decode kernel reads them in place (chroma table permitting)
    assert ctx.last_kernel_path == (1 if cbits <= 14 else 0)
    ctx.set_kernel_path(1)
    out_generic = t.decode(generic, w, h)
    ctx.set_kernel_path(0)
    assert torch.equal(out_tuned.view(torch.int32), out_generic.view(torch.int32))


@pytest.mark.parametrize("lmax,sc", [(1e4, 1.0), (1000.0, 20.0)])
def test_ycbcr_pq_tables_equal_per_pixel_powf(lumalib, po, torch_cuda, lmax, sc):
    """CS_YCBCR tuned kernels read PQ decode and the outer power of PQ encode from exhaustive device-built tables
    (luma_pq_tables.cuh).  Same bits as the same kernels evaluating every powf per pixel (lumacu_set_pq_tables(0)), on
    dense content: 18 decades of input incl. specials for encode; every combination class of 10-bit codes (uniform
    random planes, incl. out-of-range codes) for decode; and one frame against the CPU oracle (host libm)."""
    import torch
    from lumahdrv_b200.device import DeviceTransform

    w, h, n = 1920, 1080, 4
    t = DeviceTransform(0, ptf="PQ", ptfBitDepth=10, colorSpace="YCBCR", colorBitDepth=10, maxLum=lmax, minLum=0.01, preScaling=sc)
    ctx = t.quant.ctx
    rng = np.random.default_rng(int(lmax))
    frames = [po.noise_frame(w, h, seed=60 + i) / np.float32(sc) for i in range(n - 1)]
    wide = np.power(np.float32(10.0), rng.uniform(-12.0, 6.0, size=(3, h, w)).astype(np.float32)).astype(np.float32)
    wide.reshape(-1)[rng.integers(0, wide.size, 5000)] = rng.choice(
        np.array([0.0, -1.0, np.inf, np.nan, 1e-45, 1e-38, 1e4, 1e-10, 3e38], np.float32), 5000)
    frames.append(wide)
    # EXR-sourced content: every sample is a half-float value (looked up in the 65 536-entry input table), incl. every
    # finite half pattern, +-0, +-inf, NaN and subnormals; and a frame that mixes such samples with arbitrary floats
    allh = np.arange(65536, dtype=np.uint16).view(np.float16).astype(np.float32)
    exr = (po.noise_frame(w, h, seed=77) / np.float32(sc)).astype(np.float16).astype(np.float32)
    exr.reshape(-1)[: 3 * 65536] = np.tile(allh, 3)
    frames.append(exr)
    mixed = frames[0].copy()
    sel = rng.random(mixed.shape) < 0.5
    mixed[sel] = mixed[sel].astype(np.float16).astype(np.float32)
    frames.append(mixed)
    n = len(frames)
    rgb = torch.from_numpy(np.stack(frames)).cuda()
    ctx.set_pq_tables(False)
    ref_planes = [p.clone() for p in t.encode(rgb)]
    assert ctx.last_kernel_path == 1
    ctx.set_pq_tables(True)
    got = t.encode(rgb)
    for p, (a, b) in enumerate(zip(got, ref_planes)):
        assert torch.equal(a, b), f"encode: plane {p} differs in {(a != b).sum().item()} bytes"
    # without statistics plane 0 is searched by v = (219 y' + 16)/255 in the v-keyed table; with them, by luminance
    with_stats = t.encode(rgb, stats=t.alloc_stats(n))
    for p, (a, b) in enumerate(zip(with_stats, ref_planes)):
        assert torch.equal(a, b), f"encode with statistics: plane {p} differs in {(a != b).sum().item()} bytes"
    # decode: uniform random 16-bit words, mostly within the 10-bit range
    planes = t.alloc_planes(n, w, h)
    for pl, (pw, ph) in zip(planes, po.plane_dims(w, h, 2)):
        codes = rng.integers(0, 1024, size=(n, ph, pw), dtype=np.uint16)
        wild = rng.random((n, ph, pw)) < 0.001
        codes[wild] = rng.integers(0, 65536, size=int(wild.sum()), dtype=np.uint16)
        pl[:, :, : pw * 2] = torch.from_numpy(codes.astype("<u2").view(np.uint8).reshape(n, ph, pw * 2)).cuda()
    ctx.set_pq_tables(False)
    ref_out = t.decode(planes, w, h).clone()
    ctx.set_pq_tables(True)
    out = t.decode(planes, w, h)
    a, b = out.view(torch.int32), ref_out.view(torch.int32)
    same = (a == b) | (torch.isnan(out) & torch.isnan(ref_out))
    assert bool(same.all()), f"decode: {(~same).sum().item()} floats differ"
    ctx.set_tuning(0, 2000, 0)  # green of BOTH pixels of a pair evaluated instead of looked up (sweep knob): same bits
    both = t.decode(planes, w, h)
    ctx.set_tuning(0, 0, 0)
    assert bool(((both.view(torch.int32) == b) | (torch.isnan(both) & torch.isnan(ref_out))).all())
    # the host libm's word on one frame of each
    o = po.Oracle().setQuantizer("PQ", 10, "YCBCR", 10, lmax, 0.01)
    for f in (0, n - 2, n - 1):
        cpu_planes, _ = o.encode(frames[f].copy(), 2, sc)
        for a, b, (pw, ph) in zip(got, cpu_planes, po.plane_dims(w, h, 2)):
            assert np.array_equal(a[f].cpu().numpy()[:ph, :pw * 2], b[:ph, :pw * 2]), f"frame {f} vs the CPU oracle"
    cpu_out = o.decode([p[1].cpu().numpy() for p in planes], w, h, 2, sc)
    assert bits_equal(out[1].cpu().numpy(), cpu_out)


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
