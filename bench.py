#!/usr/bin/env python
"""bench.py -- Mpixels/s of the HDR<->integer transform round trip (encode + decode).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]

Workload (BASELINE.json metric): 3840x2160 float32 RGB frames, PQ transfer function, Lu'v'
11-bit luma / 8-bit chroma, profile 2 (4:2:0, LE-u16 planes).  One step = one pass of the hot
path over one batch of `--frames` distinct synthetic frames per GPU: LumaEncoder::encode (minus
VP9) for every frame, then LumaDecoder::decode for every frame.

* value      device-resident throughput (inputs/outputs in HBM), CUDA events, max over ranks
* e2e        same metric through the host-pointer C ABI (lumacu_encode / lumacu_decode) with pinned
             host buffers; H2D/D2H copies inside the timed region
* roofline   algorithmic bytes (15 B/px per direction) / measured kernel time vs the measured HBM peak
* cpu_baseline  the reference's own CPU code (oracle/_ref) on this box's host cores, bounded sample
* parity     outside the timed region every rank copies one WHOLE frame of the benchmark's own input, its planes and
             its decoded floats to the host and has the CPU checker (oracle/parity_check.py, own process) redo it:
             mismatching plane bytes (summed over ranks) and max ulp of the decoded floats (max over ranks)
* configs    the other BASELINE.json configurations (cfg2 1080p, cfg3 4K PQ-10 YCbCr, cfg4 4K LOG-12, cfg5 8K + per-frame
             sum/max of Y), same timing code, a few steps each, each with its own whole-frame parity check

--impl reference times the reference CPU implementation instead (rank 0 only).
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time
from pathlib import Path

ROOT = Path(__file__).resolve().parent
sys.path.insert(0, str(ROOT))

W, H = 3840, 2160
QUANT = dict(ptf="PQ", ptfBitDepth=11, colorSpace="LUV", colorBitDepth=8, profile=2, bitDepth=12, preScaling=1.0,
             maxLum=1e4, minLum=0.005)
BYTES_PER_PX_PASS = 15.0  # 12 B f32 RGB + 2 B Y + 2 * 2 B / 4 chroma, either direction
WORKLOAD = "4K (3840x2160) f32 RGB, PQ, Lu'v' 11/8-bit, profile 2 (4:2:0 LE16): encode+decode round trip"


# The other BASELINE.json configurations; all profile 2 = 15 B/px per direction.
OTHER_CONFIGS = {
    "cfg2": dict(label="1920x1080 PQ Lu'v' 11/8-bit, profile 2", w=1920, h=1080, ptf="PQ", bits=11, cs="LUV", cbits=8,
                 frames=64, stats=False),
    "cfg3": dict(label="3840x2160 PQ 10-bit + BT.2020 YCbCr 10-bit (HDR10-equivalent), profile 2", w=3840, h=2160, ptf="PQ",
                 bits=10, cs="YCBCR", cbits=10, frames=8, stats=False, also_half_float_input=True, also_test_pattern_input=True),
    "cfg4": dict(label="3840x2160 LOG 12-bit + Lu'v' 8-bit, profile 2, frame stream sharded over the ranks", w=3840, h=2160,
                 ptf="LOG", bits=12, cs="LUV", cbits=8, frames=32, min_total_frames=64, stats=False),
    "cfg5": dict(label="7680x4320 PQ Lu'v' 11/8-bit, 0.005..10000 cd/m2, per-frame sum/max/min of Y", w=7680, h=4320, ptf="PQ",
                 bits=11, cs="LUV", cbits=8, frames=8, stats=True),
}


def config_dict(n_gpus: int, frames: int) -> dict:
    """`config` of the JSON line: the same object in both arms (the driver compares them)."""
    return {"workload": WORKLOAD, "frames_per_gpu_per_step": frames,
            "input": "seeded log-uniform noise 0.005..1e4 cd/m2, distinct per frame",
            "l2": f"inputs larger than L2: {frames * (12 + 3 + 12) * W * H / 1e6:.0f} MB touched per step per GPU vs 126 MB L2",
            "parallelism": f"frame shards x{n_gpus}, LUT broadcast only"}


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=100)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--frames", type=int, default=32, help="frames per GPU per step")
    ap.add_argument("--e2e-frames", type=int, default=4, help="frames per GPU per end-to-end step")
    ap.add_argument("--cpu-frames", type=int, default=1, help="frames per worker in the CPU baseline sample")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-sustained", action="store_true", help="skip the second, power-capped-regime measurement")
    ap.add_argument("--no-parity", action="store_true", help="skip the whole-frame parity checks against the CPU checker")
    ap.add_argument("--no-configs", action="store_true", help="skip the cfg2..cfg5 block")
    ap.add_argument("--config-steps", type=int, default=10, help="timed steps per configuration of the configs block")
    return ap.parse_args()


def host_cores() -> int:
    try:
        return len(os.sched_getaffinity(0))
    except AttributeError:
        return os.cpu_count() or 1


def host_cpu_model() -> str:
    try:
        for ln in open("/proc/cpuinfo"):
            if ln.lower().startswith("model name"):
                return ln.split(":", 1)[1].strip()
    except OSError:
        pass
    return "unknown"


def run_cpu_baseline(frames_per_worker: int, workers: int | None = None) -> dict:
    """The reference's CPU path in a separate process tree (never shares a process with CUDA)."""
    workers = workers or host_cores()
    cmd = [sys.executable, str(ROOT / "oracle" / "cpu_baseline.py"), "--workers", str(workers), "--frames",
           str(frames_per_worker), "--width", str(W), "--height", str(H)]
    out = subprocess.run(cmd, check=True, capture_output=True, text=True).stdout.strip().splitlines()[-1]
    return json.loads(out)


def reference_arm(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    # the workers are forked pool processes that exit without the driver's loader hook seeing them: open the very
    # library they time in this process too, once (dlopen only, nothing is computed here)
    ref_so = None
    try:
        from oracle import pyoracle as po
        if po.reference_available():
            po._ref_lib()
            ref_so = str(po.REF_SO.relative_to(ROOT))
        else:
            po._oracle_lib()
            ref_so = str(po.ORACLE_SO.relative_to(ROOT))
    except Exception as e:  # noqa: BLE001
        ref_so = f"unavailable in parent: {e}"
    res = None
    times = []
    for i in range(args.warmup + args.steps):
        t0 = time.perf_counter()
        res = run_cpu_baseline(args.cpu_frames)
        if i >= args.warmup:
            times.append((time.perf_counter() - t0, res))
    vals = sorted(r["value"] for _, r in times)
    value = vals[len(vals) // 2]
    px_per_step = res["cores"] * args.cpu_frames * W * H
    line = {
        "impl": "reference", "metric": "Mpixels/s encode+decode (PQ Lu'v' 4K float32)", "value": value,
        "unit": "Mpixels/s", "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": px_per_step / value / 1e3, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f32", "data": "synthetic",
        "config": config_dict(args.gpus, args.frames),
        "note": "reference CPU implementation (LumaEncoder::encode + LumaDecoder::decode, VP9 stubbed out), one independent "
                f"frame stream per host core; each step = {res['cores'] * args.cpu_frames} frames of the workload",
        "cpu_baseline": {"value": value, "unit": "Mpixels/s", "cores": res["cores"], "kind": res["kind"],
                         "sample": res["sample"], "per_core_mpx_s": res["per_core_mpx_s"], "library": ref_so,
                         "cpu_model": host_cpu_model()},
        "e2e": {"value": value, "unit": "Mpixels/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled during the timed region."""
    FIELDS = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
              "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
              "clocks_event_reasons.sw_power_cap")

    def __init__(self, index: int):
        self.index = index
        self.proc = None
        self.lines = []

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.FIELDS}", "--format=csv,noheader,nounits",
                                          "-lms", "20", "-i", str(self.index)], stdout=subprocess.PIPE, text=True)
            self.thread = threading.Thread(target=self._pump, daemon=True)
            self.thread.start()
        except OSError:
            self.proc = None

    def _pump(self):
        for line in self.proc.stdout:
            self.lines.append((time.perf_counter(), line.strip()))

    def finish(self):
        if not self.proc:
            return
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except subprocess.TimeoutExpired:
            self.proc.kill()
        self.proc = None
        self.done = True

    def window(self, t_begin: float, t_end: float) -> dict:
        """Clocks / power / throttle reasons of the samples received inside [t_begin, t_end] (nvidia-smi reports with
        about one sampling period (20 ms) of delay, hence the shifted window)."""
        if not getattr(self, "done", False) and not self.lines:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        lines = [ln for t, ln in self.lines if t_begin + 0.02 <= t <= t_end + 0.02]
        if not lines:
            lines = [ln for _, ln in self.lines[-3:]]
        sm, mx, power, reasons = [], [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for ln in lines:
            parts = [p.strip() for p in ln.split(",")]
            if len(parts) < 7:
                continue
            try:
                sm.append(float(parts[0]))
                mx.append(float(parts[1]))
                power.append(float(parts[2]))
            except ValueError:
                continue
            for name, val in zip(names, parts[3:7]):
                if val.lower().startswith("active"):
                    reasons.add(name)
        sm.sort()
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "power_w_max": max(power) if power else None, "samples": len(sm), "reasons": sorted(reasons)}


def measured_peak_gbs():
    p = ROOT / "MEASURED_PEAKS.json"
    if p.exists():
        try:
            return float(json.loads(p.read_text())["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


def measured_traffic(kind: str, px_per_launch: float):
    """DRAM bytes per launch of the dominant kernel from the newest committed `ncu --set full` capture
    (profiles/*_traffic.json, written by scripts/summarize_ncu.py), scaled to this run's pixels per launch."""
    best = None
    for p in sorted((ROOT / "profiles").glob("*_traffic.json")):
        try:
            d = json.loads(p.read_text())
            best = (d["kernels"][kind]["dram_bytes_per_pixel"] * px_per_launch, p.name)
        except Exception:
            continue
    return best if best else (None, None)


def bind_to_gpu_numa(index: int):
    """Restrict this process to the CPUs NVML reports as local to GPU `index` (no-op when unavailable)."""
    try:
        import pynvml
        pynvml.nvmlInit()
        h = pynvml.nvmlDeviceGetHandleByIndex(index)
        words = pynvml.nvmlDeviceGetCpuAffinity(h, (os.cpu_count() + 63) // 64)
        cpus = {64 * i + b for i, wd in enumerate(words) for b in range(64) if (wd >> b) & 1}
        cpus &= os.sched_getaffinity(0)
        if cpus:
            os.sched_setaffinity(0, cpus)
            return len(cpus)
    except Exception:  # noqa: BLE001
        pass
    return None


def run_parity_check(params: dict, frame, planes, out) -> dict:
    """One whole frame through the CPU checker in its own process (oracle/parity_check.py); numpy arrays in, dict out."""
    import shutil
    import tempfile

    import numpy as np
    need = 2 * (frame.nbytes + out.nbytes + sum(p.nbytes for p in planes)) + (64 << 20)
    base = None  # RAM-backed /dev/shm when it has room for this rank's frame (it is small in some containers), else TMPDIR
    try:
        if os.path.isdir("/dev/shm") and os.access("/dev/shm", os.W_OK) and shutil.disk_usage("/dev/shm").free > 8 * need:
            base = "/dev/shm"
    except OSError:
        base = None
    d = Path(tempfile.mkdtemp(prefix="luma_parity_", dir=base))
    try:
        try:
            (d / "params.json").write_text(json.dumps(params))
            np.save(d / "in.npy", frame)
            for i, pl in enumerate(planes):
                np.save(d / f"p{i}.npy", pl)
            np.save(d / "out.npy", out)
        except OSError as e:
            return {"error": f"could not stage the frame for the checker in {d}: {e}"[:300]}
        r = subprocess.run([sys.executable, str(ROOT / "oracle" / "parity_check.py"), str(d)], capture_output=True, text=True)
        if r.returncode != 0:
            return {"error": (r.stderr or r.stdout).strip().splitlines()[-1][:300] if (r.stderr or r.stdout).strip() else
                    f"parity_check.py exited with {r.returncode}"}
        return json.loads(r.stdout.strip().splitlines()[-1])
    finally:
        shutil.rmtree(d, ignore_errors=True)


def parity_over_ranks(local: dict, dev, world: int) -> dict:
    """Sum of mismatching plane bytes / frames / pixels, max of ulp over all ranks (one small all_reduce each)."""
    import torch
    import torch.distributed as dist
    failed = 1 if "error" in local else 0
    sums = torch.tensor([local.get("plane_mismatch_bytes", 0), 0 if failed else 1, local.get("pixels", 0), failed,
                         0 if local.get("stats_max_equal") in (None, True) else 1], dtype=torch.float64, device=dev)
    maxs = torch.tensor([local.get("max_ulp", 0), local.get("stats_sum_rel_err") or 0.0], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(sums, op=dist.ReduceOp.SUM)
        dist.all_reduce(maxs, op=dist.ReduceOp.MAX)
    res = {"plane_mismatch_bytes": int(sums[0].item()), "max_ulp": int(maxs[0].item()), "frames_checked": int(sums[1].item()),
           "pixels_checked": int(sums[2].item()), "ranks": world, "ranks_failed": int(sums[3].item()),
           "checker": local.get("checker"), "whole_frames": True,
           "what": "one whole frame per rank of the timed batch: planes byte for byte, decoded floats in ulp, vs the CPU checker"}
    if local.get("stats_max_equal") is not None:
        res["stats_max_mismatches"] = int(sums[4].item())
        res["stats_sum_max_rel_err"] = float(maxs[1].item())
    if failed:
        res["error_rank_local"] = local["error"]
    return res


def make_transform(cfg: dict, local: int, rank: int, world: int, dev):
    """DeviceTransform for one configuration; at N > 1 rank 0's host-built quantizer is broadcast (the only collective)."""
    import lumahdrv_b200 as L
    from lumahdrv_b200.device import DeviceTransform
    from lumahdrv_b200.shard import broadcast_quantizer, pack_quantizer, packed_quantizer_size, unpack_quantizer
    shared_lut = None
    if world > 1:
        vec = None
        if rank == 0:
            lut = L.build_lut(cfg["ptf"], cfg["bits"], cfg["maxLum"], cfg["minLum"])
            vec = pack_quantizer(lut, (1 << cfg["cbits"]) - 1, getattr(L, "CS_" + cfg["cs"]), cfg["maxLum"], cfg["minLum"],
                                 cfg["preScaling"], cfg["profile"], ptf=getattr(L, "PTF_" + cfg["ptf"]))
        shared_lut = unpack_quantizer(broadcast_quantizer(vec, packed_quantizer_size(cfg["bits"]), dev, src=0))["lut"]
    return DeviceTransform(local, ptf=cfg["ptf"], ptfBitDepth=cfg["bits"], colorSpace=cfg["cs"], colorBitDepth=cfg["cbits"],
                           maxLum=cfg["maxLum"], minLum=cfg["minLum"], profile=cfg["profile"], preScaling=cfg["preScaling"],
                           lut=shared_lut)


def synth_frames(F: int, w: int, h: int, first_index: int, dev):
    """F distinct seeded log-uniform frames, 0.005 .. 1e4 cd/m2 (SURVEY 8d input (2)), generated in HBM."""
    import torch
    rgb = torch.empty((F, 3, h, w), dtype=torch.float32, device=dev)
    for i in range(F):
        g = torch.Generator(device=dev).manual_seed(0x9E3779B9 + first_index + i)
        u = torch.rand((3, h, w), generator=g, device=dev, dtype=torch.float32)
        rgb[i] = 0.005 * torch.pow(torch.tensor(2.0e6, device=dev), u)
    return rgb


def frame_parity(t, cfg: dict, rgb, planes, out, stats, idx: int, dev, world: int) -> dict:
    """Copy frame `idx` of the batch (input, planes, decoded floats, stats) to the host and check it on the CPU."""
    import torch
    torch.cuda.synchronize()
    params = {"w": cfg["w"], "h": cfg["h"], "profile": cfg["profile"], "ptf": cfg["ptf"], "ptfBitDepth": cfg["bits"],
              "colorSpace": cfg["cs"], "colorBitDepth": cfg["cbits"], "preScaling": cfg["preScaling"], "maxLum": cfg["maxLum"],
              "minLum": cfg["minLum"]}
    if stats is not None:
        st = t.stats_to_numpy(stats)[idx]
        params["stats"] = {"sum": float(st["sum"]), "max": float(st["max"]), "min": float(st["min"])}
    local = run_parity_check(params, rgb[idx].cpu().numpy(), [p[idx].cpu().numpy() for p in planes], out[idx].cpu().numpy())
    return parity_over_ranks(local, dev, world)


def measure_config(name: str, cfg: dict, args, local: int, rank: int, world: int, dev, peak: float) -> dict:
    """One of cfg2..cfg5: device-resident round trip with the headline's timing code (barrier + synchronize on both sides,
    CUDA events on the launching stream around every kernel, max over ranks), then a whole-frame parity check."""
    import torch
    import torch.distributed as dist
    from lumahdrv_b200.shard import frame_shard
    cfg = dict(dict(profile=2, preScaling=1.0, maxLum=1e4, minLum=0.005), **cfg)
    w, h = cfg["w"], cfg["h"]
    F = max(cfg["frames"], -(-cfg.get("min_total_frames", 0) // world))
    mine = frame_shard(F * world, rank, world)
    t = make_transform(cfg, local, rank, world, dev)
    rgb = synth_frames(F, w, h, 1000 + mine.start, dev)
    planes = t.alloc_planes(F, w, h)
    out = torch.empty_like(rgb)
    stats = t.alloc_stats(F) if cfg["stats"] else None
    n_steps, n_warm = max(1, args.config_steps), 3
    for _ in range(n_warm):
        t.encode(rgb, planes=planes, stats=stats)
        t.decode(planes, w, h, out=out)
    ev = [[torch.cuda.Event(enable_timing=True) for _ in range(3)] for _ in range(n_steps)]
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    for k in range(n_steps):
        ev[k][0].record()
        t.encode(rgb, planes=planes, stats=stats)
        ev[k][1].record()
        t.decode(planes, w, h, out=out)
        ev[k][2].record()
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    tt = torch.tensor([ev[0][0].elapsed_time(ev[-1][2]), sum(e[0].elapsed_time(e[1]) for e in ev) / n_steps,
                       sum(e[1].elapsed_time(e[2]) for e in ev) / n_steps], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
    tot, enc_ms, dec_ms = (float(v) for v in tt.tolist())
    px = F * w * h
    bytes_pass = BYTES_PER_PX_PASS * px
    res = {"workload": cfg["label"], "frames_per_gpu_per_step": F, "frames_total_per_step": F * world, "steps": n_steps,
           "warmup": n_warm, "value": world * px * n_steps / (tot / 1e3) / 1e6, "unit": "Mpixels/s",
           "encode_ms": enc_ms, "decode_ms": dec_ms, "bytes_per_px_per_direction": BYTES_PER_PX_PASS,
           "encode_gbs": bytes_pass / (enc_ms / 1e3) / 1e9, "decode_gbs": bytes_pass / (dec_ms / 1e3) / 1e9,
           "frac": bytes_pass / (max(enc_ms, dec_ms) / 1e3) / 1e9 / peak,
           "round_trip_frac_of_peak": 2 * bytes_pass / ((enc_ms + dec_ms) / 1e3) / 1e9 / peak,
           "search": t.quant.search_info()}
    if cfg.get("also_half_float_input"):
        # Informational: the same frames rounded to half-float values -- what the reference's own input path delivers
        # (OpenEXR half pixels, src/exr_interface.cpp:73-143).  Such samples are looked up in a 65 536-entry table
        # instead of being pushed through powf; arbitrary floats (the configuration's synthetic noise above) are not.
        rgb_h = rgb.half().float().contiguous()
        for _ in range(n_warm):
            t.encode(rgb_h, planes=planes, stats=stats)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize()
        e0.record()
        for _ in range(n_steps):
            t.encode(rgb_h, planes=planes, stats=stats)
        e1.record()
        torch.cuda.synchronize()
        enc_h = e0.elapsed_time(e1) / n_steps
        th = torch.tensor([enc_h], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(th, op=dist.ReduceOp.MAX)
        enc_h = float(th.item())
        res["half_float_input"] = {"what": "same frames rounded to half-float (EXR-sourced content): encode reads PQ from the input table",
                                   "encode_ms": enc_h, "encode_gbs": bytes_pass / (enc_h / 1e3) / 1e9,
                                   "round_trip_value": world * px / ((enc_h + dec_ms) / 1e3) / 1e6}
        t.encode(rgb, planes=planes, stats=stats)  # the planes of the configuration's own input again (parity below)
        del rgb_h
    if cfg.get("also_test_pattern_input"):
        # Informational: the reference's own test content (ExrInterface::testFrame, src/exr_interface.cpp:50-70: smooth
        # ramps and bands, arbitrary floats) instead of noise.  Same arithmetic per pixel; the table lookups of
        # neighbouring pixels now share L2 sectors, which is what bounds this configuration on noise.
        rgb_t = t.test_frame(w, h)[None].expand(F, -1, -1, -1).contiguous()
        for _ in range(n_warm):
            t.encode(rgb_t, planes=planes, stats=stats)
            t.decode(planes, w, h, out=out)
        evt = [torch.cuda.Event(enable_timing=True) for _ in range(3)]
        torch.cuda.synchronize()
        evt[0].record()
        for _ in range(n_steps):
            t.encode(rgb_t, planes=planes, stats=stats)
        evt[1].record()
        for _ in range(n_steps):
            t.decode(planes, w, h, out=out)
        evt[2].record()
        torch.cuda.synchronize()
        tp = torch.tensor([evt[0].elapsed_time(evt[1]) / n_steps, evt[1].elapsed_time(evt[2]) / n_steps], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(tp, op=dist.ReduceOp.MAX)
        enc_t, dec_t = (float(v) for v in tp.tolist())
        res["test_pattern_input"] = {"what": "the reference's test pattern (ExrInterface::testFrame) in every frame instead of noise",
                                     "encode_ms": enc_t, "decode_ms": dec_t, "encode_gbs": bytes_pass / (enc_t / 1e3) / 1e9,
                                     "decode_gbs": bytes_pass / (dec_t / 1e3) / 1e9,
                                     "round_trip_value": world * px / ((enc_t + dec_t) / 1e3) / 1e6}
        del rgb_t
        t.encode(rgb, planes=planes, stats=stats)  # back to the configuration's own input (parity below)
        t.decode(planes, w, h, out=out)
    if stats is not None:
        st = t.stats_to_numpy(stats)
        res["stats_frame0"] = {"mean_Y": float(st["sum"][0]) / (w * h), "max_Y": float(st["max"][0]), "min_Y": float(st["min"][0])}
    if not args.no_parity:
        res["parity"] = frame_parity(t, cfg, rgb, planes, out, stats, rank % F, dev, world)
    del rgb, planes, out, stats, t
    torch.cuda.empty_cache()
    return res


def ours_arm(args):
    import numpy as np
    import torch
    import torch.distributed as dist

    import lumahdrv_b200 as L
    from lumahdrv_b200.shard import frame_shard

    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device (the transform has no CPU fallback)")
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)

    # ---- CPU baseline first (rank 0, N=1 only), before the GPU is busy
    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        cpu = run_cpu_baseline(args.cpu_frames)

    # ---- NUMA: keep this rank's host threads (and therefore its pinned buffers, first touch) next to its GPU
    affinity = None
    if os.environ.get("LUMA_BENCH_AFFINITY", "1") != "0":
        affinity = bind_to_gpu_numa(local)

    # ---- quantizer: rank 0 builds the LUT with its libm, everyone receives it (the only collective)
    HEAD = dict(w=W, h=H, ptf=QUANT["ptf"], bits=QUANT["ptfBitDepth"], cs=QUANT["colorSpace"], cbits=QUANT["colorBitDepth"],
                profile=QUANT["profile"], preScaling=QUANT["preScaling"], maxLum=QUANT["maxLum"], minLum=QUANT["minLum"])
    t = make_transform(HEAD, local, rank, world, dev)

    # ---- this rank's shard of the synthetic frame stream (weak scaling: F frames per GPU per step)
    F = args.frames
    mine = frame_shard(F * world, rank, world)
    rgb = synth_frames(F, W, H, mine.start, dev)
    planes = t.alloc_planes(F, W, H)
    out = torch.empty_like(rgb)
    stats = t.alloc_stats(F)
    torch.cuda.synchronize()

    def step():
        t.encode(rgb, planes=planes, stats=stats)
        t.decode(planes, W, H, out=out)

    def barrier():
        if world > 1:
            dist.barrier()

    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    n_warm = max(args.warmup, 3)
    for _ in range(n_warm):
        step()
    torch.cuda.synchronize()

    def timed(n_steps):
        """n_steps steps bracketed by barrier + synchronize; CUDA events on the launching stream around every kernel."""
        ev = [[torch.cuda.Event(enable_timing=True) for _ in range(3)] for _ in range(n_steps)]
        barrier()
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        for k in range(n_steps):
            ev[k][0].record()
            t.encode(rgb, planes=planes, stats=stats)
            ev[k][1].record()
            t.decode(planes, W, H, out=out)
            ev[k][2].record()
        torch.cuda.synchronize()
        barrier()
        t1 = time.perf_counter()
        tot = ev[0][0].elapsed_time(ev[-1][2])
        enc = sum(e[0].elapsed_time(e[1]) for e in ev) / n_steps
        dec = sum(e[1].elapsed_time(e[2]) for e in ev) / n_steps
        tt = torch.tensor([tot, enc, dec], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        return [float(v) for v in tt.tolist()] + [t0, t1]

    # ---- timed region: W warm-up steps (above), then exactly K steps
    launches0 = t.launch_count
    total_ms, enc_ms, dec_ms, t_wall0, t_wall1 = timed(args.steps)
    t_wall = t_wall1 - t_wall0
    launches = t.launch_count - launches0

    # ---- the same measurement in the sustained regime: ~0.5 s of back-to-back steps first, so that the board sits at
    # its power cap (SM clock ~1.7-1.9 GHz) when the clock starts.  Reported next to `value`, not instead of it.
    sustained = None
    if not args.no_sustained:
        t_p0 = time.perf_counter()
        n_pre = 0
        while time.perf_counter() - t_p0 < 0.5:
            step()
            n_pre += 1
            if n_pre % 8 == 0:
                torch.cuda.synchronize()
        s_tot, s_enc, s_dec, s_t0, s_t1 = timed(args.steps)
        sustained = {"value": world * F * W * H * args.steps / (s_tot / 1e3) / 1e6, "unit": "Mpixels/s", "preroll_steps": n_pre,
                     "encode_ms": s_enc, "decode_ms": s_dec, "window": (s_t0, s_t1)}
    if rank == 0:
        sampler.finish()
    clocks = sampler.window(t_wall0, t_wall1) if rank == 0 else None
    if sustained is not None:
        w0, w1 = sustained.pop("window")
        sustained["clocks"] = sampler.window(w0, w1) if rank == 0 else None

    px_step = F * W * H  # per GPU
    value = world * px_step * args.steps / (total_ms / 1e3) / 1e6

    # context for the roofline fraction: a plain device-to-device copy on THIS lease, measured the way
    # MEASURED_PEAKS.json was (torch copy_ over 1 GiB, read + write bytes, best of 10, CUDA events)
    copy_gbs = None
    if rank == 0:
        src_c = torch.empty(1 << 30, dtype=torch.uint8, device=dev)
        dst_c = torch.empty_like(src_c)
        best = float("inf")
        for i in range(12):
            c0, c1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            c0.record()
            dst_c.copy_(src_c)
            c1.record()
            c1.synchronize()
            if i >= 2:
                best = min(best, c0.elapsed_time(c1))
        copy_gbs = 2 * (1 << 30) / (best / 1e3) / 1e9
        del src_c, dst_c

    # ---- parity of the benchmark's OWN data, outside the timed region: every rank has the CPU checker redo one whole
    # frame of the batch the timed steps just produced (input -> planes -> decoded floats, plus the Y statistics)
    st = t.stats_to_numpy(stats)
    assert np.all(np.isfinite(st["sum"])) and np.all(st["sum"] > 0)
    parity = None
    if not args.no_parity:
        parity = frame_parity(t, HEAD, rgb, planes, out, stats, (rank * 5 + 3) % F, dev, world)

    # ---- end to end through the host-pointer C ABI, pinned host memory, copies inside the timed region
    e2e = None
    if not args.no_e2e:
        Fe = args.e2e_frames
        h_in = torch.empty((Fe, 3, H, W), dtype=torch.float32).pin_memory()
        h_in.copy_(rgb[:Fe].cpu())
        h_out = torch.empty((3, H, W), dtype=torch.float32).pin_memory()
        strides = L.vpx_strides(W, QUANT["profile"])
        h_planes = [torch.empty((ph, s), dtype=torch.uint8).pin_memory()
                    for (pw, ph), s in zip(L.plane_dims(W, H, QUANT["profile"]), strides)]
        enc = L.LumaEncoder(local)
        enc.setParams(L.LumaEncoderParams(**{k: QUANT[k] for k in ("ptfBitDepth", "colorBitDepth", "preScaling", "minLum",
                                                                     "maxLum", "profile", "bitDepth")},
                                          ptf=L.PTF_PQ, colorSpace=L.CS_LUV))
        enc.initialize(None, W, H)
        dec = L.LumaDecoder(local)
        dec.setParams(L.LumaDecoderParams(ptf=L.PTF_PQ, colorSpace=L.CS_LUV, ptfBitDepth=QUANT["ptfBitDepth"],
                                          colorBitDepth=QUANT["colorBitDepth"], profile=QUANT["profile"]))
        dec.initialize()
        # one band per call: with the encoder and the decoder running concurrently the PCIe link is already busy in
        # both directions, and banding only adds API calls (measured: 2.80 ms per frame pair at 1 band, 3.62 at 8;
        # a lone encoder or decoder is fastest at 8 bands: 1.96 / 2.05 ms vs 2.33 / 2.28 -- scripts/e2e_probe.py)
        for obj in (enc, dec):
            L._lib.check(obj.m_quant._lib.lumacu_set_host_bands(obj.m_quant.ctx.handle, 1), obj.m_quant.ctx.handle, "bands")
        dec.m_frame = h_out.numpy()
        np_in = h_in.numpy()
        np_planes = [p.numpy() for p in h_planes]

        # Two host threads, like the two programs of the reference (lumaenc | lumadec): the encoder thread runs
        # LumaEncoder.encode on frame i+1 while the decoder thread runs LumaDecoder.decode on the planes of
        # frame i (each object has its own context and streams; ctypes releases the GIL inside the C call), so
        # both PCIe directions are busy.  Planes travel GPU -> host -> GPU like they would through VP9.
        import queue
        # four host plane buffers between the two threads: with only two the stages run in lockstep and the decoder's
        # small H2D copy queues behind the encoder's 100 MB one (3.8 ms per frame); with slack the loops settle
        # into a phase where both PCIe directions stay busy (2.9 ms per frame; scripts/e2e_probe.py)
        NSLOTS = 4
        slots = [[p.clone().pin_memory().numpy() for p in h_planes] for _ in range(NSLOTS)]
        h_outs = [torch.empty((3, H, W), dtype=torch.float32).pin_memory().numpy() for _ in range(2)]
        free_q, full_q = queue.Queue(), queue.Queue()
        errors = []
        last_slot = [0]

        half_io = {"in": None, "out": None}  # set for the informational half-RGBA pass below

        def enc_thread(nframes):
            try:
                for i in range(nframes):
                    s = free_q.get()
                    if half_io["in"] is not None:
                        enc.encode_half_rgba(half_io["in"][i % Fe], 7, slots[s])   # H2D 8 B/px, 2 kernels, D2H 3 B/px
                    else:
                        enc.encode(np_in[i % Fe], slots[s])       # H2D 12 B/px, kernel, D2H 3 B/px
                    full_q.put(s)
            except Exception as e:  # noqa: BLE001
                errors.append(e)
            full_q.put(None)

        def dec_thread():
            try:
                k = 0
                while True:
                    s = full_q.get()
                    if s is None:
                        return
                    if half_io["out"] is not None:
                        dec.decode_half_rgba(slots[s], W, H, out=half_io["out"][k & 1])   # H2D 3 B/px, 2 kernels, D2H 8 B/px
                    else:
                        dec.m_frame = h_outs[k & 1]
                        dec.decode(slots[s], W, H)            # H2D 3 B/px, kernel, D2H 12 B/px
                    last_slot[0] = s
                    free_q.put(s)
                    k += 1
            except Exception as e:  # noqa: BLE001
                errors.append(e)

        def e2e_run(nframes):
            while not free_q.empty():
                free_q.get()
            for i in range(NSLOTS):
                free_q.put(i)
            te = threading.Thread(target=enc_thread, args=(nframes,))
            td = threading.Thread(target=dec_thread)
            te.start()
            td.start()
            te.join()
            td.join()
            if errors:
                raise errors[0]

        def e2e_step():
            e2e_run(Fe)

        e2e_steps = max(3, min(args.steps, 10))
        for _ in range(2):
            e2e_step()
        l0 = enc.m_quant.ctx.launch_count + dec.m_quant.ctx.launch_count
        barrier()
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        e2e_run(Fe * e2e_steps)   # K steps back to back, the pipeline stays full across step boundaries
        torch.cuda.synchronize()
        dt = time.perf_counter() - t0
        e2e_launches = enc.m_quant.ctx.launch_count + dec.m_quant.ctx.launch_count - l0
        # the last decoded frame must be the round trip of its input (spot check, outside the timed region)
        chk = t.decode(t.encode(rgb[(Fe * e2e_steps - 1) % Fe][None]), W, H)[0].cpu().numpy()
        assert np.array_equal(chk.view(np.uint32), h_outs[(Fe * e2e_steps - 1) & 1].view(np.uint32)), "e2e result differs"
        # ... and the CPU checker redoes that frame from the host buffers the C ABI filled (planes + decoded floats)
        e2e_parity = None
        if not args.no_parity:
            prm = {"w": W, "h": H, "profile": QUANT["profile"], "ptf": QUANT["ptf"], "ptfBitDepth": QUANT["ptfBitDepth"],
                   "colorSpace": QUANT["colorSpace"], "colorBitDepth": QUANT["colorBitDepth"], "preScaling": QUANT["preScaling"],
                   "maxLum": QUANT["maxLum"], "minLum": QUANT["minLum"],
                   "stats": {k: float(enc.last_stats[k]) for k in ("sum", "max", "min")}}
            e2e_parity = parity_over_ranks(run_parity_check(prm, np_in[(Fe * e2e_steps - 1) % Fe], slots[last_slot[0]],
                                                            h_outs[(Fe * e2e_steps - 1) & 1]), dev, world)
        # Informational: the same pipeline with OpenEXR's pixel format at both ends (what lumaenc reads and lumadec writes:
        # half-float Imf::Rgba, src/exr_interface.cpp:73-143, :157-187) -- lumacu_encode_half_rgba / lumacu_decode_half_rgba
        # move 8 B/px over the bus instead of 12 and run the two pixel loops on the device.  Not the headline metric (its
        # input is not f32 RGB); checked against the device path below.
        h_half = torch.empty((Fe, H, W, 4), dtype=torch.float16).pin_memory()
        h_half[..., :3].copy_(h_in.permute(0, 2, 3, 1))
        h_half[..., 3] = 1.0
        half_io["in"] = h_half.numpy()
        half_io["out"] = [torch.empty((H, W, 4), dtype=torch.float16).pin_memory().numpy() for _ in range(2)]
        e2e_step()
        barrier()
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        e2e_run(Fe * e2e_steps)
        torch.cuda.synchronize()
        dt_half = time.perf_counter() - t0
        last = (Fe * e2e_steps - 1) % Fe
        chk_h = t.frame_to_half_rgba(t.decode(t.encode(t.half_rgba_to_frame(h_half[last].to(dev), 7)[None]), W, H)[0]).cpu().numpy()
        half_ok = bool(np.array_equal(chk_h.view(np.uint16), half_io["out"][(Fe * e2e_steps - 1) & 1].view(np.uint16)))
        assert half_ok, "half-RGBA e2e result differs from the device path"
        half_io["in"] = half_io["out"] = None
        tt = torch.tensor([dt, dt_half], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        dt, dt_half = (float(v) for v in tt.tolist())
        plane_bytes = sum(pw * ph * 2 for pw, ph in L.plane_dims(W, H, QUANT["profile"]))
        e2e = {"value": world * Fe * W * H * e2e_steps / dt / 1e6, "unit": "Mpixels/s",
               "h2d_bytes_per_step": Fe * (12 * W * H + plane_bytes), "d2h_bytes_per_step": Fe * (12 * W * H + plane_bytes),
               "steps": e2e_steps, "frames_per_step": Fe, "gpu_launches": e2e_launches, "parity": e2e_parity,
               "api": "LumaEncoder.encode / LumaDecoder.decode -> lumacu_encode + lumacu_decode (host pointers, pinned), "
                      "encoder and decoder objects on two host threads (frame i+1 encodes while frame i decodes)",
               "half_rgba_io": {"value": world * Fe * W * H * e2e_steps / dt_half / 1e6, "unit": "Mpixels/s",
                                "h2d_bytes_per_step": Fe * (8 * W * H + plane_bytes), "d2h_bytes_per_step": Fe * (8 * W * H + plane_bytes),
                                "equals_device_path": half_ok,
                                "what": "informational, NOT the headline metric: same pipeline with OpenEXR half-float RGBA pixels "
                                        "at both ends (lumacu_encode_half_rgba / lumacu_decode_half_rgba), 8 B/px on the bus"}}

    # ---- the other BASELINE configurations (a few steps each, same timing code, own parity check)
    configs = None
    if not args.no_configs:
        del rgb, planes, out
        torch.cuda.empty_cache()
        peak_c, _ = measured_peak_gbs()
        configs = {}
        for name, cfg in OTHER_CONFIGS.items():
            try:
                configs[name] = measure_config(name, cfg, args, local, rank, world, dev, peak_c)
            except Exception as e:  # noqa: BLE001 -- a failing side configuration must not lose the headline line
                if world > 1:
                    raise
                configs[name] = {"error": f"{type(e).__name__}: {e}"[:300]}

    if rank == 0:
        peak, peak_src = measured_peak_gbs()
        bytes_pass = BYTES_PER_PX_PASS * px_step
        dom = "encode_kernel" if enc_ms >= dec_ms else "decode_kernel"
        dom_ms = max(enc_ms, dec_ms)
        achieved = bytes_pass / (dom_ms / 1e3) / 1e9
        traffic, traffic_src = measured_traffic("encode" if dom == "encode_kernel" else "decode", px_step)
        line = {
            "metric": "Mpixels/s encode+decode (PQ Lu'v' 4K float32)", "value": value, "unit": "Mpixels/s",
            "n_gpus": world, "steps": args.steps, "warmup": n_warm, "ms_per_step": total_ms / args.steps,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": config_dict(world, F),
            "roofline": {"bound": "hbm", "kernel": dom, "achieved": achieved, "peak": peak, "unit": "GB/s",
                         "frac": achieved / peak, "peak_source": peak_src, "traffic": traffic,
                         "traffic_source": (f"ncu --set full capture {traffic_src}: dram__bytes_read.sum + dram__bytes_write.sum per "
                                            f"pixel x pixels per launch") if traffic else None,
                         "algorithmic_bytes_per_launch": bytes_pass, "kernel_ms": dom_ms,
                         "encode_ms": enc_ms, "decode_ms": dec_ms,
                         "encode_gbs": bytes_pass / (enc_ms / 1e3) / 1e9, "decode_gbs": bytes_pass / (dec_ms / 1e3) / 1e9,
                         "round_trip_frac_of_peak": (2 * bytes_pass / ((enc_ms + dec_ms) / 1e3) / 1e9) / peak,
                         "frac_of_nominal_8tbs": (2 * bytes_pass / ((enc_ms + dec_ms) / 1e3) / 1e9) / 8000.0,
                         "copy_gbs_this_lease": copy_gbs},
            "gpu_launches": launches,
            "clocks": clocks,
            "sustained": sustained,
            "host_cpus_bound": affinity,
            "wall_s_timed_region": t_wall,
        }
        if e2e:
            line["e2e"] = e2e
        if parity:
            line["parity"] = parity
        if configs:
            line["configs"] = configs
        if cpu:
            line["cpu_baseline"] = {"value": cpu["value"], "unit": "Mpixels/s", "cores": cpu["cores"], "kind": cpu["kind"],
                                    "sample": cpu["sample"], "per_core_mpx_s": cpu["per_core_mpx_s"],
                                    "cpu_model": host_cpu_model()}
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


def main():
    args = parse_args()
    if args.impl == "reference":
        reference_arm(args)
    else:
        ours_arm(args)


if __name__ == "__main__":
    main()
