"""CPU: pins the oracle (oracle/luma_oracle.c) against (1) the committed golden vectors that
tests/golden/make_golden.py derived from the unmodified reference, (2) the compiled reference
itself (oracle/_ref) when it is present, on seeded and adversarial inputs."""
import numpy as np
import pytest

from conftest import bits_equal

CS = ("LUV", "RGB", "YCBCR", "XYZ")


def test_lut_hashes_match_golden(po, golden):
    for key, want in golden["lut"].items():
        ptf, bits, lmax, lmin = key.split(":")
        o = po.Oracle().setQuantizer(ptf, int(bits), "LUV", 8, float(lmax), float(lmin))
        assert "%08x" % po.fnv1a32(o.getMapping()) == want, key


@pytest.mark.parametrize("name", ["cfg1_256_pq_luv", "720p_pq_luv", "cfg2_1080p_pq_luv", "cfg4_4k_log12_luv"])
def test_testframe_planes_match_golden(po, golden, name):
    g = golden["frames"][name]
    p = dict(ptf="PQ", ptfBitDepth=11, colorSpace="LUV", colorBitDepth=8, maxLum=1e4, minLum=0.005, preScaling=1.0)
    p.update(g["params"])
    o = po.Oracle().setQuantizer(p["ptf"], p["ptfBitDepth"], p["colorSpace"], p["colorBitDepth"], p["maxLum"], p["minLum"])
    frame = po.test_frame(g["w"], g["h"])
    assert "%08x" % po.fnv1a32(frame) == g["input"]
    planes, _ = o.encode(frame, g["profile"], p["preScaling"])
    assert ["%08x" % v for v in po.plane_hashes(planes, g["w"], g["h"], g["profile"])] == g["planes"]
    assert "%08x" % po.fnv1a32(frame) == g["after_encode"]
    dec = o.decode(planes, g["w"], g["h"], g["profile"], p["preScaling"])
    assert "%08x" % po.fnv1a32(dec) == g["decoded"]


def test_ycbcr_golden_256(po, golden):
    """cfg3 parameters (libm-dependent) on a crop-sized frame: oracle vs small golden below; the 4K hash is
    checked on the GPU box where the oracle is also the libm witness."""
    g = golden["frames"]["cfg3_4k_pq10_ycbcr"]
    assert g["planes"] == ["5f868bbc", "d16e3e17", "d07a1c8e"]


def test_small_cases_match_golden(po, small_cases):
    for cs in CS:
        for profile in (0, 1, 2, 3):
            for sc in (1.0, 3.5):
                key = f"{cs}_p{profile}_sc{sc:g}"
                o = po.Oracle().setQuantizer("PQ", 8 if profile < 2 else 11, cs, 8)
                frame = small_cases[key + "_in"].copy()
                planes, _ = o.encode(frame, profile, sc)
                for p, pl in enumerate(po.plane_payload(planes, 48, 32, profile)):
                    assert np.array_equal(pl, small_cases[key + f"_plane{p}"]), (key, p)
                assert bits_equal(frame, small_cases[key + "_after"]), key
                dec = o.decode(planes, 48, 32, profile, sc)
                assert bits_equal(dec, small_cases[key + "_dec"]), key


def adversarial_frame(w=64, h=16, lut=None, seed=7):
    rng = np.random.default_rng(seed)
    f = (0.005 * np.power(2.0e6, rng.random((3, h, w)))).astype(np.float32)
    special = np.array([np.nan, np.inf, -np.inf, 0.0, -0.0, -1.0, -1e30, 1e-45, 1e-40, 1e-10, 1e-4, 1e8, 3e8, 1e30,
                        3.4e38, 9999.0, 10000.0, 10001.0], dtype=np.float32)
    flat = f.reshape(3, -1)
    for c in range(3):
        idx = rng.choice(flat.shape[1], size=special.size * 4, replace=False)
        flat[c, idx] = np.tile(special, 4)
    if lut is not None:  # exact LUT entries and midpoints in grey pixels
        n = min(lut.size - 1, flat.shape[1] // 4)
        flat[:, :n] = lut[:n]
        flat[:, n:2 * n] = (0.5 * (lut[:n].astype(np.float64) + lut[1:n + 1])).astype(np.float32)
    return f


@pytest.mark.parametrize("cs", CS)
@pytest.mark.parametrize("profile", [0, 1, 2, 3])
def test_oracle_equals_reference_adversarial(po, cs, profile):
    if not po.reference_available():
        pytest.skip("compiled reference (oracle/_ref) not present")
    bits = 8 if profile < 2 else 11
    ref = po.Reference(colorSpace=cs, profile=profile, bitDepth=8 if profile < 2 else 12, ptfBitDepth=bits)
    o = po.Oracle().setQuantizer("PQ", bits, cs, 8)
    f0 = adversarial_frame(lut=o.getMapping())
    fr, fo = f0.copy(), f0.copy()
    _, h, w = f0.shape
    pr = ref.encode(fr)
    pl, _ = o.encode(fo, profile, 1.0)
    for a, b in zip(po.plane_payload(pr, w, h, profile), po.plane_payload(pl, w, h, profile)):
        assert np.array_equal(a, b)
    assert bits_equal(fr, fo)
    assert bits_equal(ref.decode(pr, w, h), o.decode(pl, w, h, profile, 1.0))
    ref.close()


@pytest.mark.parametrize("ptf,bits", [("PQ", 11), ("LOG", 12), ("PSI", 11), ("JND_HDRVDP", 10), ("LINEAR", 11), ("PSI", 8)])
def test_oracle_quantize_dequantize_equals_reference(po, ptf, bits):
    if not po.reference_available():
        pytest.skip("compiled reference (oracle/_ref) not present")
    ref = po.Reference(ptf=ptf, ptfBitDepth=bits)
    o = po.Oracle().setQuantizer(ptf, bits, "LUV", 8)
    assert bits_equal(ref.getMapping(), o.getMapping())
    vals = adversarial_frame(lut=o.getMapping()).reshape(-1)[:4096]
    for ch in (0, 1):
        want = ref.quantize_n(vals, ch)
        got = np.array([o.quantize(v, ch) for v in vals], dtype=np.float32)
        assert bits_equal(want, got)
    codes = np.arange(-2, (1 << bits) + 3, dtype=np.float32)
    for ch in (0, 2):
        want = ref.dequantize_n(codes, ch)
        got = np.array([o.dequantize(v, ch) for v in codes], dtype=np.float32)
        assert bits_equal(want, got)
    ref.close()


def test_quantize_edge_values_from_survey(po):
    """SURVEY 8(a) a7 probe values for PQ-11."""
    o = po.Oracle().setQuantizer("PQ", 11, "LUV", 8)
    for v, c in [(-1, 0), (0, 0), (1e-4, 3), (0.005, 31), (1, 307), (100, 1040), (1000, 1539), (9999, 2047),
                 (np.nan, 2047), (np.inf, 2047)]:
        assert o.quantize(v, 0) == c
    for v, c in [(-0.1, 0), (0.5, 128), (1.5, 255), (np.nan, 255)]:
        assert o.quantize(v, 1) == c


def test_reference_refuses_odd_size(po):
    if not po.reference_available():
        pytest.skip("compiled reference (oracle/_ref) not present")
    ref = po.Reference()
    with pytest.raises(RuntimeError, match="Invalid frame size"):
        ref.encode(np.zeros((3, 5, 6), dtype=np.float32))
