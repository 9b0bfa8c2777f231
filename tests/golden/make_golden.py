"""Generates tests/golden/golden.json and tests/golden/small_cases.npz from the UNMODIFIED
reference (oracle/_ref/libluma_ref.so = /root/reference sources compiled in place, driven
through LumaEncoder::encode / LumaDecoder::decode).  Run in the build container only:

    python tests/golden/make_golden.py

The hashes listed in SURVEY.md section 8(c) (derived independently during the survey, tier-1
and tier-2 builds of the reference) are asserted here as a cross-check.
"""
import json
import sys
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parents[2]
sys.path.insert(0, str(ROOT))
from oracle import pyoracle as po  # noqa: E402

HERE = Path(__file__).resolve().parent

SURVEY_LUT = {("PQ", 11, 1e4, 0.005): "a8a2c0cb", ("PQ", 10, 1e4, 0.005): "f8b902e5", ("PQ", 10, 1000.0, 0.01): "1020dfe9",
              ("PQ", 12, 1e4, 0.005): "160481f6", ("PQ", 8, 1e4, 0.005): "d2af5103", ("LOG", 12, 1e4, 0.005): "fd136d7d",
              ("LOG", 11, 1e4, 0.005): "e47f4b34", ("PSI", 11, 1e4, 0.005): "06f55592",
              ("JND_HDRVDP", 12, 1e4, 0.005): "f0c47ea3", ("LINEAR", 11, 1e4, 0.005): "8ef83bc1"}

# testFrame(w,h) -> encode, profile 2: name -> (w, h, reference params, survey plane hashes or None)
FRAME_CASES = {
    "cfg1_256_pq_luv": (256, 256, dict(), ["b28f2401", "d4688866", "12dca67c"]),
    "720p_pq_luv": (1280, 720, dict(), ["50667568", "672af8ea", "d6ca5e23"]),
    "cfg2_1080p_pq_luv": (1920, 1080, dict(), ["93e1e551", "ceb15c74", "1becd710"]),
    "cfg4_4k_log12_luv": (3840, 2160, dict(ptf="LOG", ptfBitDepth=12), ["de68e488", "2471849e", "d8680f3b"]),
    "cfg5_8k_pq_luv": (7680, 4320, dict(), ["7b7de9d8", "eef189f0", "9aba493e"]),
    "cfg3_4k_pq10_ycbcr": (3840, 2160, dict(ptfBitDepth=10, colorBitDepth=10, colorSpace="YCBCR", bitDepth=10),
                           ["5f868bbc", "d16e3e17", "d07a1c8e"]),
    "cfg3_4k_hdr10_readme": (3840, 2160, dict(ptfBitDepth=10, colorBitDepth=10, colorSpace="YCBCR", bitDepth=10,
                                              maxLum=1000.0, minLum=0.01, preScaling=20.0),
                             ["4d06a56c", "0a10a784", "9e56c140"]),
    "4k_pq_luv": (3840, 2160, dict(), None),
}


def h8(x) -> str:
    return "%08x" % po.fnv1a32(x)


def main():
    po.build()
    out = {"_generator": "tests/golden/make_golden.py", "_source": "oracle/_ref/libluma_ref.so (unmodified reference)",
           "hash": "FNV-1a-32 over bytes", "lut": {}, "frames": {}}
    for (ptf, bits, lmax, lmin), want in SURVEY_LUT.items():
        ref = po.Reference(ptf=ptf, ptfBitDepth=bits, maxLum=lmax, minLum=lmin)
        got = h8(ref.getMapping())
        assert got == want, (ptf, bits, got, want)
        out["lut"][f"{ptf}:{bits}:{lmax:g}:{lmin:g}"] = got
        ref.close()
    for name, (w, h, params, survey) in FRAME_CASES.items():
        ref = po.Reference(**params)
        frame = po.test_frame(w, h)
        in_hash = h8(frame)
        planes = ref.encode(frame)
        ph = ["%08x" % v for v in po.plane_hashes(planes, w, h, ref.profile)]
        if survey is not None:
            assert ph == survey, (name, ph, survey)
        dec = ref.decode(planes, w, h)
        out["frames"][name] = {"w": w, "h": h, "params": params, "profile": ref.profile, "input": in_hash,
                               "planes": ph, "after_encode": h8(frame), "decoded": h8(dec)}
        print(name, ph, out["frames"][name]["decoded"], flush=True)
        ref.close()
    assert out["frames"]["cfg1_256_pq_luv"]["decoded"] == "cacc4b81"
    assert out["frames"]["cfg1_256_pq_luv"]["input"] == "b2082ac1"
    assert out["frames"]["cfg1_256_pq_luv"]["after_encode"] == "4479ddee"
    (HERE / "golden.json").write_text(json.dumps(out, indent=1) + "\n")

    # small full-data cases: every colour space x profile, seeded noise 48x32 (+ prescaling variant)
    small = {}
    i = 0
    for cs in ("LUV", "RGB", "YCBCR", "XYZ"):
        for profile in (0, 1, 2, 3):
            for sc in (1.0, 3.5):
                bit_depth = 8 if profile < 2 else 12
                ref = po.Reference(colorSpace=cs, profile=profile, bitDepth=bit_depth, preScaling=sc,
                                   ptfBitDepth=8 if profile < 2 else 11)
                frame = po.noise_frame(48, 32, seed=1000 + i)
                key = f"{cs}_p{profile}_sc{sc:g}"
                small[key + "_in"] = frame.copy()
                planes = ref.encode(frame)
                assert ref.profile == profile
                for p, pl in enumerate(po.plane_payload(planes, 48, 32, profile)):
                    small[key + f"_plane{p}"] = pl
                small[key + "_after"] = frame
                small[key + "_dec"] = ref.decode(planes, 48, 32)
                ref.close()
                i += 1
    np.savez_compressed(HERE / "small_cases.npz", **small)
    print("wrote", len(small), "arrays")


if __name__ == "__main__":
    main()
