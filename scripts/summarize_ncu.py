#!/usr/bin/env python
"""Turn the ncu outputs a gpurun call brought back (gpurun_out/) into the small, tracked summaries under
profiles/.

    python scripts/summarize_ncu.py TAG [--round r01]

reads   gpurun_out/TAG_launches.csv   (ncu --metrics gpu__time_duration.sum,... --csv launch list of bench.py)
        gpurun_out/TAG_prof.ncu-rep   (ncu --set full capture of the two transform kernels)
writes  profiles/<round>_<TAG>_launches.csv      one row per launch: kernel, grid, block, us, DRAM bytes, instructions
        profiles/<round>_<TAG>_launch_summary.json   per-kernel means + each kernel's share of the step
        profiles/<round>_<TAG>_full_metrics.csv  the `--set full` metrics that explain the roofline number
"""
import argparse
import collections
import csv
import json
import subprocess
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
FULL_KEEP = [
    "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "dram__bytes.sum.per_second",
    "dram__throughput.avg.pct_of_peak_sustained_elapsed", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
    "smsp__inst_executed.sum", "smsp__issue_active.avg.pct_of_peak_sustained_active",
    "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread", "launch__grid_size",
    "launch__block_size", "launch__waves_per_multiprocessor", "launch__occupancy_limit_registers",
    "launch__occupancy_limit_shared_mem", "launch__shared_mem_per_block_dynamic", "lts__t_sector_hit_rate.pct",
    "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
    "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio",
    "sm__inst_executed_pipe_fma.sum", "sm__inst_executed_pipe_alu.sum", "sm__inst_executed_pipe_xu.sum",
    "sm__inst_executed_pipe_lsu.sum", "smsp__cycles_active.avg", "sm__cycles_elapsed.max",
]


def short(name: str) -> str:
    name = name.replace("void ", "").replace("lumacu::", "")
    return name if len(name) < 100 else name[:97] + "..."


def launches(tag: str, rnd: str):
    src = ROOT / "gpurun_out" / f"{tag}_launches.csv"
    if not src.exists():
        return
    rows = list(csv.reader(ln for ln in open(src) if ln.startswith('"')))
    hdr = rows[0]
    per = collections.OrderedDict()
    for r in rows[1:]:
        d = dict(zip(hdr, r))
        e = per.setdefault(d["ID"], {"id": int(d["ID"]), "kernel": short(d["Kernel Name"]), "grid": d["Grid Size"],
                                     "block": d["Block Size"]})
        try:
            e[d["Metric Name"]] = float(d["Metric Value"].replace(",", ""))
        except ValueError:
            pass
    out = ROOT / "profiles" / f"{rnd}_{tag}_launches.csv"
    cols = ["id", "kernel", "grid", "block", "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
            "smsp__inst_executed.sum"]
    with open(out, "w", newline="") as f:
        w = csv.writer(f)
        w.writerow(["id", "kernel", "grid", "block", "time_ns", "dram_read_bytes", "dram_write_bytes", "warp_instructions"])
        for e in per.values():
            w.writerow([e.get(c, "") for c in cols])
    agg = collections.OrderedDict()
    for e in per.values():
        a = agg.setdefault(e["kernel"], {"launches": 0, "time_ns": 0.0, "dram_read": 0.0, "dram_write": 0.0, "inst": 0.0})
        a["launches"] += 1
        a["time_ns"] += e.get("gpu__time_duration.sum", 0.0)
        a["dram_read"] += e.get("dram__bytes_read.sum", 0.0)
        a["dram_write"] += e.get("dram__bytes_write.sum", 0.0)
        a["inst"] += e.get("smsp__inst_executed.sum", 0.0)
    ours = {k: v for k, v in agg.items() if "_kernel<" in k and ("encode" in k or "decode" in k)}
    step_ns = sum(v["time_ns"] for v in ours.values())
    summ = {"source": src.name, "note": "ncu launch list (cold-cache, serialised): use the SHARES, not the absolutes",
            "kernels": {}}
    for k, v in agg.items():
        n = v["launches"]
        summ["kernels"][k] = {"launches": n, "mean_us": v["time_ns"] / n / 1e3, "mean_dram_read_MB": v["dram_read"] / n / 1e6,
                              "mean_dram_write_MB": v["dram_write"] / n / 1e6, "mean_warp_inst": v["inst"] / n,
                              "share_of_transform_step": (v["time_ns"] / step_ns) if k in ours and step_ns else None,
                              "in_timed_region": k in ours}
    (ROOT / "profiles" / f"{rnd}_{tag}_launch_summary.json").write_text(json.dumps(summ, indent=1))
    print("wrote", out.name, "and launch summary")


def full(tag: str, rnd: str, frames: int = 32):
    rep = ROOT / "gpurun_out" / f"{tag}_prof.ncu-rep"
    if not rep.exists():
        return
    raw = subprocess.run(["ncu", "-i", str(rep), "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(ln for ln in raw.splitlines() if ln.startswith('"')))
    hdr, units = rows[0], rows[1]
    out = ROOT / "profiles" / f"{rnd}_{tag}_full_metrics.csv"
    with open(out, "w", newline="") as f:
        w = csv.writer(f)
        w.writerow(["metric", "unit"] + [short(r[hdr.index("Kernel Name")]) for r in rows[2:]])
        for m in FULL_KEEP:
            if m in hdr:
                i = hdr.index(m)
                w.writerow([m, units[i]] + [r[i] for r in rows[2:]])
    print("wrote", out.name)
    # DRAM traffic per launch of the two transform kernels (bench.py reports it as roofline.traffic, scaled per pixel)
    names = [r[hdr.index("Kernel Name")] for r in rows[2:]]
    rd, wr = hdr.index("dram__bytes_read.sum"), hdr.index("dram__bytes_write.sum")
    scale = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
    tr = {"source": rep.name, "frames_per_launch": frames, "pixels_per_launch": frames * 3840 * 2160, "kernels": {}}
    for r, nm in zip(rows[2:], names):
        key = "encode" if "encode" in nm else "decode" if "decode" in nm else None
        if key:
            b = float(r[rd]) * scale[units[rd]] + float(r[wr]) * scale[units[wr]]
            tr["kernels"][key] = {"kernel": short(nm), "dram_bytes_per_launch": b, "dram_bytes_per_pixel": b / tr["pixels_per_launch"]}
    (ROOT / "profiles" / f"{rnd}_{tag}_traffic.json").write_text(json.dumps(tr, indent=1))


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("tag")
    ap.add_argument("--round", default="r01")
    ap.add_argument("--frames", type=int, default=32, help="frames per launch of the profiled bench run (scripts/gpu_profile.sh: 32)")
    a = ap.parse_args()
    (ROOT / "profiles").mkdir(exist_ok=True)
    launches(a.tag, a.round)
    full(a.tag, a.round, a.frames)
