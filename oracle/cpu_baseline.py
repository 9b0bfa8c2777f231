"""TEST/BENCH INFRASTRUCTURE ONLY: times the reference's own CPU implementation of the hot path.

Runs LumaEncoder::encode + LumaDecoder::decode of the UNMODIFIED reference (oracle/_ref/libluma_ref.so,
kind "reference") -- or the C restatement (kind "port") when the compiled reference is absent -- on
seeded synthetic frames, one independent frame stream per worker process (the reference is
single-threaded on this stage; frames are independent, so this is the CPU's best case).

    python oracle/cpu_baseline.py --workers 8 --frames 2 --width 3840 --height 2160

prints one JSON object: {"value": Mpx/s, "cores": workers, "kind": ..., "sample": ..., ...}
"""
import argparse
import json
import multiprocessing as mp
import os
import sys
import time
from pathlib import Path

ROOT = Path(__file__).resolve().parents[1]
if str(ROOT) not in sys.path:
    sys.path.insert(0, str(ROOT))


def _worker(args):
    idx, cfg = args
    from oracle import pyoracle as po
    w, h, frames = cfg["width"], cfg["height"], cfg["frames"]
    q = cfg["quant"]
    use_ref = po.reference_available() and not cfg["force_port"]
    if use_ref:
        impl = po.Reference(ptf=q["ptf"], ptfBitDepth=q["ptfBitDepth"], colorSpace=q["colorSpace"],
                            colorBitDepth=q["colorBitDepth"], profile=q["profile"], bitDepth=q["bitDepth"],
                            preScaling=q["preScaling"], maxLum=q["maxLum"], minLum=q["minLum"], o0=cfg.get("o0", False))
    else:
        impl = po.Oracle().setQuantizer(q["ptf"], q["ptfBitDepth"], q["colorSpace"], q["colorBitDepth"], q["maxLum"],
                                        q["minLum"])
    src = po.noise_frame(w, h, seed=0x9E3779B97F4A7C15 + idx)
    # warm-up on a small frame (page in the library, build the LUT)
    small = po.noise_frame(64, 32, seed=1)
    if use_ref:
        warm = po.Reference(ptf=q["ptf"], ptfBitDepth=q["ptfBitDepth"], colorSpace=q["colorSpace"],
                            colorBitDepth=q["colorBitDepth"], profile=q["profile"], bitDepth=q["bitDepth"])
        warm.decode(warm.encode(small), 64, 32)
        warm.close()
    t_enc = t_dec = 0.0
    for _ in range(frames):
        f = src.copy()
        t0 = time.perf_counter()
        planes = impl.encode(f) if use_ref else impl.encode(f, q["profile"], q["preScaling"])[0]
        t1 = time.perf_counter()
        out = impl.decode(planes, w, h) if use_ref else impl.decode(planes, w, h, q["profile"], q["preScaling"])
        t2 = time.perf_counter()
        t_enc += t1 - t0
        t_dec += t2 - t1
    assert out.shape == (3, h, w)
    return {"enc_s": t_enc, "dec_s": t_dec, "kind": "reference" if use_ref else "port"}


def run(workers: int, frames: int, width: int, height: int, quant: dict, force_port: bool = False, o0: bool = False) -> dict:
    cfg = {"width": width, "height": height, "frames": frames, "quant": quant, "force_port": force_port, "o0": o0}
    t0 = time.perf_counter()
    if workers == 1:
        res = [_worker((0, cfg))]
    else:
        with mp.get_context("fork").Pool(workers) as pool:
            res = pool.map(_worker, [(i, cfg) for i in range(workers)])
    wall = time.perf_counter() - t0
    px = float(width) * height * frames
    slowest = max(r["enc_s"] + r["dec_s"] for r in res)
    return {
        "value": workers * px / slowest / 1e6,  # all workers run concurrently; the slowest one bounds the job
        "unit": "Mpixels/s",
        "cores": workers,
        "kind": res[0]["kind"],
        "sample": f"{workers} worker(s) x {frames} frame(s) {width}x{height} round trip (encode+decode), "
                  + (("libluma_ref_O0.so (-O0, the reference's own CMake default)" if o0 else
                      "libluma_ref.so -O2 (bit-identical to the reference -O0 build)") if res[0]["kind"] == "reference"
                     else "C restatement -O2"),
        "per_core_mpx_s": px / (sum(r["enc_s"] + r["dec_s"] for r in res) / len(res)) / 1e6,
        "encode_mpx_s_per_core": px / (sum(r["enc_s"] for r in res) / len(res)) / 1e6,
        "decode_mpx_s_per_core": px / (sum(r["dec_s"] for r in res) / len(res)) / 1e6,
        "wall_s": wall,
        "host_cpus": os.cpu_count(),
    }


DEFAULT_QUANT = dict(ptf="PQ", ptfBitDepth=11, colorSpace="LUV", colorBitDepth=8, profile=2, bitDepth=12, preScaling=1.0,
                     maxLum=1e4, minLum=0.005)

if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--workers", type=int, default=0)
    ap.add_argument("--frames", type=int, default=2)
    ap.add_argument("--width", type=int, default=3840)
    ap.add_argument("--height", type=int, default=2160)
    ap.add_argument("--port", action="store_true")
    ap.add_argument("--o0", action="store_true", help="time the -O0 build of the reference (its CMake default)")
    a = ap.parse_args()
    n = a.workers or len(os.sched_getaffinity(0))
    print(json.dumps(run(n, a.frames, a.width, a.height, DEFAULT_QUANT, a.port, a.o0)))
