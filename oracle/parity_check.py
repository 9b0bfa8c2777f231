"""TEST/BENCH INFRASTRUCTURE ONLY: whole-frame parity of GPU results against the CPU checker.

    python oracle/parity_check.py DIR

DIR holds what one rank of bench.py (or a test) copied back from its GPU, outside any timed region:
    params.json   {"w","h","profile","ptf","ptfBitDepth","colorSpace","colorBitDepth","preScaling","maxLum","minLum",
                   optional "stats": {"sum","max","min"}, optional "lut": "lut.npy"}
    in.npy        the input frame, float32 [3,h,w]
    p0.npy p1.npy p2.npy   the GPU's integer planes, uint8 [rows, pitch]
    out.npy       the GPU's decoded frame, float32 [3,h,w] (decoded from the GPU's own planes)

The checker is the unmodified reference (oracle/_ref/libluma_ref.so, "reference") when it was built, else the
C restatement ("port").  Prints one JSON object:
    {"plane_mismatch_bytes": N, "max_ulp": U, "pixels": w*h, "checker": ..., "stats_max_equal": bool|null,
     "stats_sum_rel_err": float|null}
Runs in its own process so that the product process never loads anything from oracle/.
"""
import json
import sys
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parents[1]
if str(ROOT) not in sys.path:
    sys.path.insert(0, str(ROOT))


def max_ulp(a: np.ndarray, b: np.ndarray) -> int:
    ok = ~(np.isnan(a) & np.isnan(b))
    ia = a.view(np.int32).astype(np.int64)
    ib = b.view(np.int32).astype(np.int64)
    ia = np.where(ia < 0, -(ia & 0x7FFFFFFF), ia)
    ib = np.where(ib < 0, -(ib & 0x7FFFFFFF), ib)
    d = np.abs(ia - ib)[ok]
    return int(d.max()) if d.size else 0


def check(d: Path) -> dict:
    from oracle import pyoracle as po

    p = json.loads((d / "params.json").read_text())
    w, h, profile = p["w"], p["h"], p["profile"]
    frame = np.load(d / "in.npy")
    gpu_planes = [np.load(d / f"p{i}.npy") for i in range(3)]
    gpu_out = np.load(d / "out.npy")
    lut = np.load(d / p["lut"]) if p.get("lut") else None
    use_ref = po.reference_available() and lut is None
    if use_ref:
        impl = po.Reference(ptf=p["ptf"], ptfBitDepth=p["ptfBitDepth"], colorSpace=p["colorSpace"],
                            colorBitDepth=p["colorBitDepth"], profile=profile, bitDepth=12 if profile > 1 else 8,
                            preScaling=p["preScaling"], maxLum=p["maxLum"], minLum=p["minLum"])
        f = frame.copy()
        ref_planes = impl.encode(f)
        ref_out = impl.decode(ref_planes, w, h)
    else:
        impl = po.Oracle().setQuantizer(p["ptf"], p["ptfBitDepth"], p["colorSpace"], p["colorBitDepth"], p["maxLum"], p["minLum"])
        if lut is not None:
            impl.setMapping(lut)
        f = frame.copy()
        ref_planes, _ = impl.encode(f, profile, p["preScaling"])
        ref_out = impl.decode(ref_planes, w, h, profile, p["preScaling"])
    nb = 2 if profile > 1 else 1
    bad = 0
    for a, b, (pw, ph) in zip(gpu_planes, ref_planes, po.plane_dims(w, h, profile)):
        bad += int(np.count_nonzero(a[:ph, : pw * nb] != b[:ph, : pw * nb]))
    res = {"plane_mismatch_bytes": bad, "max_ulp": max_ulp(gpu_out, ref_out), "pixels": w * h,
           "checker": "reference" if use_ref else "port", "stats_max_equal": None, "stats_sum_rel_err": None}
    if p.get("stats"):
        y = f[0].astype(np.float64)  # encode() left the colour-transformed frame in place: plane 0 is what was summed
        res["stats_max_equal"] = bool(np.float32(p["stats"]["max"]) == np.float32(y.max()))
        res["stats_sum_rel_err"] = float(abs(p["stats"]["sum"] - y.sum()) / abs(y.sum()))
    return res


if __name__ == "__main__":
    print(json.dumps(check(Path(sys.argv[1]))))
