"""Quantizer metadata wire format: lumacu_metadata_pack / lumacu_metadata_unpack against the
attachment payloads the UNMODIFIED reference encoder writes (src/luma_encoder.cpp:78-106), captured from the
in-memory container double by tests/cxx/build/ref_roundtrip, and against LumaDecoder::initialize's read-back
rules (src/luma_decoder.cpp:79-122).  Host-only: runs without a GPU."""
import ctypes as C
import struct
import subprocess
from pathlib import Path

import numpy as np
import pytest

ROOT = Path(__file__).resolve().parent.parent
B = ROOT / "tests" / "cxx" / "build"
PTF = {"PSI": 0, "PQ": 1, "LOG": 2, "JND_HDRVDP": 3, "LINEAR": 4}
CS = {"LUV": 0, "RGB": 1, "YCBCR": 2, "XYZ": 3}


def _records(blob: bytes):
    out, off = [], 0
    while off + 8 <= len(blob):
        rid, n = struct.unpack_from("<II", blob, off)
        out.append((rid, blob[off + 8:off + 8 + n]))
        off += 8 + n
    assert off == len(blob)
    return out


def _pack(lumalib, ptf, bits, cs, cbits, sc, lmax, lmin):
    from lumahdrv_b200.shard import pack_quantizer
    lut = lumalib.build_lut(ptf, bits, lmax, lmin)
    blob = pack_quantizer(lut, (1 << cbits) - 1, CS[cs], lmax, lmin, sc, 2, ptf=PTF[ptf])
    return lut, blob


@pytest.mark.parametrize("ptf,bits,cs,cbits,sc,lmax,lmin", [("PQ", 11, "LUV", 8, 1.0, 1e4, 0.005), ("PQ", 10, "YCBCR", 10, 20.0, 1000.0, 0.01),
                                                            ("LOG", 12, "LUV", 8, 1.0, 1e4, 0.005), ("PSI", 11, "RGB", 8, 0.5, 1e4, 0.005),
                                                            ("LINEAR", 8, "XYZ", 8, 1.0, 1e4, 0.005)])
def test_payloads_equal_what_the_reference_encoder_writes(lumalib, po, ptf, bits, cs, cbits, sc, lmax, lmin):
    if not (B / "ref_roundtrip").exists():
        pytest.skip("tests/cxx/build/ref_roundtrip not built")
    profile, depth = (2, 12) if bits > 8 else (0, 8)
    r = subprocess.run([str(B / "ref_roundtrip"), "64", "32", str(PTF[ptf]), str(CS[cs]), str(bits), str(cbits), str(profile),
                        str(depth), str(sc), "1", str(lmax), str(lmin)], capture_output=True, text=True, timeout=120, cwd=str(B))
    assert r.returncode == 0, r.stdout + r.stderr
    ref = {}
    for ln in r.stdout.splitlines():
        if ln.startswith("att "):
            _, rid, _, size, h = ln.split()
            ref[int(rid)] = (int(size), h)
    assert sorted(ref) == [430, 431, 432, 433, 434, 435, 436]
    _, blob = _pack(lumalib, ptf, bits, cs, cbits, sc, lmax, lmin)
    recs = _records(bytes(blob[:-1]))
    assert [rid for rid, _ in recs] == [430, 431, 432, 433, 434, 435, 436]
    for rid, payload in recs:
        assert (len(payload), "%08x" % po.fnv1a32(np.frombuffer(payload, dtype=np.uint8))) == ref[rid], f"attachment {rid}"


def test_unpack_follows_the_decoder_rules(lumalib):
    from lumahdrv_b200.shard import unpack_quantizer
    lut, blob = _pack(lumalib, "PQ", 11, "LUV", 8, 1.0, 1e4, 0.005)
    got = unpack_quantizer(blob)
    assert got["ptf"] == PTF["PQ"] and got["ptf_bit_depth"] == 11 and got["color_bit_depth"] == 8 and got["profile"] == 2
    assert np.array_equal(got["lut"].view(np.uint32), lut.view(np.uint32))
    # attachment 434 holds getSize() = maxVal floats: the last entry never travels, the decoder rebuilds it
    lut2 = lut.copy()
    lut2[-1] = 123.0
    lut2[5] = 0.125
    from lumahdrv_b200.shard import pack_quantizer
    got2 = unpack_quantizer(pack_quantizer(lut2, 255, 0, 1e4, 0.005, 1.0, 2, ptf=1))
    assert got2["lut"][5] == 0.125 and got2["lut"][-1] == lut[-1]


def test_unpack_errors(lumalib):
    from lumahdrv_b200 import LumaException
    from lumahdrv_b200.shard import unpack_quantizer
    _, blob = _pack(lumalib, "PQ", 11, "LUV", 8, 1.0, 1e4, 0.005)
    raw = bytes(blob[:-1])
    # drop record 433 (colour space): mandatory in the reference (src/luma_decoder.cpp:113-118)
    recs = [r for r in _records(raw) if r[0] != 433]
    short = b"".join(struct.pack("<II", rid, len(p)) + p for rid, p in recs)
    with pytest.raises(LumaException, match="Failed to locate Luma HDRv meta data"):
        unpack_quantizer(np.frombuffer(short + b"\x02", dtype=np.uint8))
    with pytest.raises(LumaException, match="overruns"):
        unpack_quantizer(np.frombuffer(raw[:100] + b"\x02", dtype=np.uint8))
    # unknown records are skipped; optional 435/436 fall back to the decoder's defaults
    recs = [r for r in _records(raw) if r[0] not in (435, 436)] + [(999, b"abc")]
    odd = b"".join(struct.pack("<II", rid, len(p)) + p for rid, p in recs)
    got = unpack_quantizer(np.frombuffer(odd + b"\x02", dtype=np.uint8))
    assert got["pre_scaling"] == 1.0 and got["max_lum"] == 1e4 and abs(got["min_lum"] - 0.005) < 1e-9


def test_pack_reports_size(lumalib):
    from lumahdrv_b200._lib import Metadata
    from lumahdrv_b200.shard import packed_quantizer_size
    lib = lumalib.lib()
    lut = lumalib.build_lut("PQ", 10, 1e4, 0.005)
    used = C.c_size_t(0)
    m = Metadata(10, 8, 1, 0, 1.0, 1e4, 0.005)
    assert lib.lumacu_metadata_pack(C.byref(m), lut.ctypes.data, lut.size, None, 0, C.byref(used)) != 0
    assert used.value + 1 == packed_quantizer_size(10)
