"""CPU: host logic of the product library (no compute calls, no GPU)."""
import ctypes as C
import re
from pathlib import Path

import numpy as np
import pytest

from conftest import bits_equal

ROOT = Path(__file__).resolve().parents[1]


def test_abi_exports_every_declared_symbol(lumalib):
    header = (ROOT / "include" / "lumacu.h").read_text()
    declared = set(re.findall(r"\b(lumacu_[a-z0-9_]+)\s*\(", header))
    declared -= {"lumacu_ctx", "lumacu_status"}  # "a lumacu_status (0 = OK)" in the prose
    from lumahdrv_b200 import _lib
    assert declared == set(_lib.SIGNATURES), declared ^ set(_lib.SIGNATURES)
    handle = lumalib.lib()
    for name in declared:
        assert hasattr(handle, name), name
    assert handle.lumacu_version() == 200
    assert handle.lumacu_status_name(6) == b"LUMACU_ERR_NO_DEVICE"


def test_no_cpu_fallback_without_device(lumalib):
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    with pytest.raises(lumalib.LumaException, match="NO_DEVICE"):
        lumalib.Context(0)
    with pytest.raises(lumalib.LumaException):
        lumalib.LumaQuantizer()


@pytest.mark.parametrize("ptf,bits,lmax,lmin", [("PQ", 11, 1e4, 0.005), ("PQ", 10, 1000.0, 0.01), ("PQ", 12, 1e4, 0.005),
                                                ("PQ", 8, 1e4, 0.005), ("LOG", 12, 1e4, 0.005), ("LOG", 11, 1e4, 0.005),
                                                ("PSI", 11, 1e4, 0.005), ("PSI", 10, 1e4, 0.005), ("PSI", 8, 1e4, 0.005),
                                                ("JND_HDRVDP", 12, 1e4, 0.005), ("JND_HDRVDP", 9, 1e4, 0.005),
                                                ("LINEAR", 11, 1e4, 0.005), ("LINEAR", 11, 400.0, 0.005),
                                                ("PQ", 16, 1e4, 0.005)])
def test_build_lut_equals_oracle(lumalib, po, golden, ptf, bits, lmax, lmin):
    lut = lumalib.build_lut(ptf, bits, lmax, lmin)
    o = po.Oracle().setQuantizer(ptf, bits, "LUV", 8, lmax, lmin)
    assert bits_equal(lut, o.getMapping())
    key = f"{ptf}:{bits}:{lmax:g}:{lmin:g}"
    if key in golden["lut"]:
        assert "%08x" % po.fnv1a32(lut) == golden["lut"][key]


def _float_of_key(k):
    k = np.asarray(k, dtype=np.uint64)
    b = np.where(k & 0x80000000, k ^ 0x80000000, (~k) & 0xFFFFFFFF).astype(np.uint32)
    return b.view(np.float32)


@pytest.mark.parametrize("ptf,bits", [("PQ", 11), ("PQ", 10), ("PQ", 12), ("LOG", 12), ("PSI", 11), ("JND_HDRVDP", 12),
                                      ("LINEAR", 11), ("LINEAR", 12), ("PQ", 8)])
def test_thresholds_pin_reference_decision(lumalib, po, ptf, bits):
    """code(T_k) == k and code(pred(T_k)) == k-1 for every k, with the oracle's quantize."""
    lut = lumalib.build_lut(ptf, bits)
    thr = np.empty(lut.size - 1, dtype=np.uint32)
    assert lumalib.lib().lumacu_derive_thresholds(lut.ctypes.data, lut.size, thr.ctypes.data) == 1
    assert np.all(np.diff(thr.astype(np.int64)) > 0)
    o = po.Oracle().setQuantizer(ptf, bits, "LUV", 8)
    at, below = _float_of_key(thr), _float_of_key(thr - 1)
    k = np.arange(1, lut.size)
    step = max(1, k.size // 512)  # the scalar ctypes oracle is slow; sample + both ends
    sel = np.unique(np.concatenate([k[::step] - 1, np.arange(0, 8), np.arange(k.size - 8, k.size)]))
    for i in sel:
        assert o.quantize(at[i], 0) == k[i]
        assert o.quantize(below[i], 0) == k[i] - 1
    if po.reference_available():
        ref = po.Reference(ptf=ptf, ptfBitDepth=bits)
        assert np.array_equal(ref.quantize_n(at, 0), k.astype(np.float32))
        assert np.array_equal(ref.quantize_n(below, 0), (k - 1).astype(np.float32))
        ref.close()
    sh, base, nb, wk = C.c_uint32(), C.c_uint32(), C.c_uint32(), C.c_uint32()
    assert lumalib.lib().lumacu_plan_buckets(thr.ctypes.data, thr.size, C.byref(sh), C.byref(base), C.byref(nb),
                                             C.byref(wk)) == 1
    assert 1 <= nb.value <= 8192 and wk.value >= 1
    buckets = (thr >> sh.value) - base.value
    assert buckets.min() == 0 and buckets.max() == nb.value - 1
    assert np.bincount(buckets).max() == wk.value


def test_thresholds_refuse_non_monotone_lut(lumalib):
    lut = lumalib.build_lut("PQ", 8).copy()
    thr = np.empty(lut.size - 1, dtype=np.uint32)
    lut[10] = lut[9]
    assert lumalib.lib().lumacu_derive_thresholds(lut.ctypes.data, lut.size, thr.ctypes.data) == 0
    lut = lumalib.build_lut("PQ", 8).copy()
    lut[100] = np.nan
    assert lumalib.lib().lumacu_derive_thresholds(lut.ctypes.data, lut.size, thr.ctypes.data) == 0


def test_build_lut_errors(lumalib):
    with pytest.raises(lumalib.LumaException, match="INVALID_ARGUMENT"):
        lumalib.build_lut(9, 8)
    out = np.empty(4, dtype=np.float32)
    assert lumalib.lib().lumacu_build_lut(1, 8, 1e4, 0.005, out.ctypes.data, out.size) == 1


def test_geometry_helpers(lumalib, po):
    for w in (2, 6, 62, 64, 1280, 1920, 3840):
        for profile in range(4):
            assert lumalib.vpx_strides(w, profile) == po.vpx_strides(w, profile)
            assert lumalib.plane_dims(w, 10, profile) == po.plane_dims(w, 10, profile)


def test_frame_shards():
    from lumahdrv_b200.shard import frame_shard
    for n in (0, 1, 7, 64, 65):
        for world in (1, 2, 3, 8):
            got = [frame_shard(n, r, world) for r in range(world)]
            assert sum(len(g) for g in got) == n
            assert sorted(i for g in got for i in g) == list(range(n))
            assert max(len(g) for g in got) - min(len(g) for g in got) <= 1


def test_bench_reference_arm_json_contract():
    """`bench.py --impl reference` (the reference's own CPU code, oracle/_ref) prints ONE JSON line with the keys the
    driver reads; the CUDA arm's config/metric/unit are the same strings."""
    import json
    import subprocess
    import sys
    from pathlib import Path

    root = Path(__file__).resolve().parent.parent
    r = subprocess.run([sys.executable, str(root / "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "0"],
                       capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [ln for ln in r.stdout.strip().splitlines() if ln.startswith("{")]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["unit"] == "Mpixels/s" and d["higher_is_better"] is True
    assert d["metric"].startswith("Mpixels/s encode+decode") and "3840x2160" in d["config"]["workload"]
    assert d["value"] > 0 and d["cpu_baseline"]["kind"] in ("reference", "port") and d["cpu_baseline"]["cores"] >= 1
    assert d["e2e"] == {"value": d["value"], "unit": "Mpixels/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert d["gpu_launches"] == 0


def _ordered_key(v):
    b = np.asarray(v, dtype=np.float32).view(np.uint32).astype(np.uint64)
    k = np.where(b & 0x80000000, (~b) & 0xFFFFFFFF, b | 0x80000000)
    return np.where(np.isnan(v), 0xFFFFFFFF, k).astype(np.uint64)


def test_thresholds_equal_reference_search_on_arbitrary_monotone_luts(lumalib, po):
    """The core claim behind the device search: for ANY finite strictly increasing LUT (not just the shipped PTFs --
    the decoder overlays whatever attachment 434 holds), LumaQuantizer::quantize(val, 0) == number of derived
    thresholds whose ordered key is <= key(val).  Random LUTs with wildly uneven spacing (neighbouring floats,
    denormals, negative entries, huge gaps), probes at the LUT entries, their float neighbours, midpoints and
    random values; checked against the oracle's bisect-then-nearest loop (src/luma_quantizer.cpp:219-235)."""
    rng = np.random.default_rng(2024)
    lib = lumalib.lib()
    ran = 0
    for trial in range(48):
        bits = int(rng.choice([1, 2, 4, 8, 10]))
        n, m = 1 << bits, 4 << bits
        kind = trial % 4
        if kind == 0:    # log-uniform positives over 30 decades
            cand = np.power(10.0, rng.uniform(-20, 10, m)).astype(np.float32)
        elif kind == 1:  # clusters of adjacent floats: differences of 1-3 ulp
            base = rng.integers(0x3A000000, 0x4A000000, size=max(1, m // 8)).astype(np.uint32)
            cand = np.concatenate([base + i * rng.integers(1, 4) for i in range(8)]).astype(np.uint32).view(np.float32)
        elif kind == 2:  # spans negative, zero and positive, with denormals
            cand = np.concatenate([-np.power(10.0, rng.uniform(-40, 3, m // 2)), [0.0],
                                   np.power(10.0, rng.uniform(-44, 3, m - m // 2 - 1))]).astype(np.float32)
        else:            # linear ramp with tiny jitter
            cand = (np.linspace(0.0, 1000.0, m) + rng.uniform(0, 1e-3, m)).astype(np.float32)
        cand = np.unique(cand[np.isfinite(cand)])  # strictly increasing after fp32 rounding (-0.0 and 0.0 collapse)
        if cand.size < n:
            continue
        start = int(rng.integers(0, cand.size - n + 1))
        lut = np.ascontiguousarray(cand[start:start + n] if kind == 1 else cand[np.sort(rng.choice(cand.size, n, replace=False))])
        ran += 1
        thr = np.empty(n - 1, dtype=np.uint32)
        assert lib.lumacu_derive_thresholds(lut.ctypes.data, n, thr.ctypes.data) == 1, f"trial {trial}"
        thr64 = thr.astype(np.uint64)
        assert np.all(np.diff(thr64.astype(np.int64)) >= 0)
        o = po.Oracle().setQuantizer("LINEAR", bits, "LUV", 8)
        assert o.getSize() == n - 1
        o.setMapping(lut)
        ubits = lut.view(np.uint32)
        nb_up = (ubits + np.where(lut >= 0, 1, -1).astype(np.int64)).astype(np.uint32).view(np.float32)
        nb_dn = (ubits - np.where(lut > 0, 1, -1).astype(np.int64)).astype(np.uint32).view(np.float32)
        mids = ((lut[:-1].astype(np.float64) + lut[1:].astype(np.float64)) / 2).astype(np.float32)
        probes = np.concatenate([lut, nb_up, nb_dn, mids, (mids.view(np.uint32) + 1).view(np.float32),
                                 (mids.view(np.uint32) - 1).view(np.float32), rng.choice(lut, 64) * np.float32(1.000001),
                                 np.array([0.0, -0.0, np.inf, -np.inf, np.nan, 1e-45, -1e-45, 3.4e38], dtype=np.float32)])
        probes = probes[: 1500]
        want = np.array([o.quantize(float(v), 0) for v in probes])
        keys = _ordered_key(probes)
        got = np.searchsorted(thr64, keys, side="right")
        bad = np.nonzero(got != want)[0]
        assert bad.size == 0, f"trial {trial} kind {kind}: val {probes[bad[0]]!r} -> {got[bad[0]]} vs reference {want[bad[0]]}"
    assert ran >= 40


def test_c_abi_header_is_plain_c99_and_example_builds(lumalib):
    """include/lumacu.h must be consumable from C (no C++-isms): compile examples/roundtrip.c as strict C99, link it
    against liblumacu.so, and -- without a GPU -- see it fail loudly at lumacu_create instead of computing on the CPU."""
    import shutil
    import subprocess
    from pathlib import Path

    root = Path(__file__).resolve().parent.parent
    gcc = shutil.which("gcc")
    if not gcc:
        pytest.skip("no gcc")
    src, exe = root / "examples" / "roundtrip.c", root / "examples" / "roundtrip"
    subprocess.run([gcc, "-std=c99", "-Wall", "-Wextra", "-Werror", "-pedantic", "-fsyntax-only", f"-I{root / 'include'}", str(src)],
                   check=True)
    subprocess.run([gcc, "-std=c99", "-O2", f"-I{root / 'include'}", str(src), f"-L{root / 'lumahdrv_b200'}", "-llumacu",
                    "-Wl,-rpath,$ORIGIN/../lumahdrv_b200", "-lm", "-o", str(exe)], check=True)
    import torch
    r = subprocess.run([str(exe)], capture_output=True, text=True, timeout=300)
    if torch.cuda.is_available():
        assert r.returncode == 0, r.stdout + r.stderr
        assert "kernel launches" in r.stdout
    else:
        assert r.returncode == 1 and "LUMACU_ERR_NO_DEVICE" in r.stderr


def test_bench_cuda_arm_refuses_to_run_without_a_gpu():
    """No silent CPU fallback: without a CUDA device the CUDA arm of bench.py stops with a clear message."""
    import subprocess
    import sys
    from pathlib import Path

    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    root = Path(__file__).resolve().parent.parent
    r = subprocess.run([sys.executable, str(root / "bench.py"), "--steps", "1", "--warmup", "0"], capture_output=True, text=True,
                       timeout=600)
    assert r.returncode != 0
    assert "no CUDA device" in (r.stderr + r.stdout)
    assert not [ln for ln in r.stdout.splitlines() if ln.startswith("{")]


def test_device_powf_sequence_equals_host_libm(tmp_path):
    """scripts/powchk.c replays the device's powf sequence (every a*b+c fused, IEEE double: powf_glibc.cuh) on the CPU and
    compares it with the host libm's powf for every float of a range.  The committed log covers every positive normal
    float for the four PQ exponents (profiles/r02_powf_exhaustive.log); here: the decades around the one input where the
    unfused sequence of earlier rounds differed from libm (y = 1/0.1593f, x = 0x1.7b1e06p-11), and the PQ range of the
    encode exponent.  Needs a host whose glibc runs its FMA build (any CPU with FMA + AVX2)."""
    import shutil
    import subprocess
    flags = open("/proc/cpuinfo").read()
    if " fma" not in flags or " avx2" not in flags:
        pytest.skip("host CPU without FMA/AVX2: glibc runs its unfused powf build here")
    if not shutil.which("gcc"):
        pytest.skip("no gcc")
    exe = tmp_path / "powchk"
    subprocess.run(["gcc", "-O2", "-ffp-contract=off", "-fopenmp", str(ROOT / "scripts" / "powchk.c"), "-lm", "-o", str(exe)], check=True)
    for args in (["inv:0.1593", "1e-5", "1e-2"], ["0.1593", "1e-9", "2.0"], ["inv:78.8438", "0.01", "1.0"]):
        out = subprocess.run([str(exe)] + args, check=True, capture_output=True, text=True).stdout
        assert "FMA replay != libm: 0" in out.splitlines()[-1], out
