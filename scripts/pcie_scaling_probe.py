#!/usr/bin/env python
"""What the HOST can move when N GPUs copy at once: plain pinned cudaMemcpyAsync, no kernels.

    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port 29533 \
        scripts/pcie_scaling_probe.py [--mb 100] [--reps 20] [--out profiles/r02_pcie_scaling_N.json]

One process per GPU (exactly how bench.py's end-to-end arm runs), each with two pinned 100 MB host buffers and two
streams.  Three phases, every rank at the same time (barrier before, max over ranks of the elapsed time after):
H2D only, D2H only, both directions at once.  Rank 0 prints / writes one JSON object with the aggregate GB/s and the
per-GPU figure -- the ceiling the end-to-end arm (12+3 B/px up, 3+12 B/px down per round trip) can be compared with.
Also reports which CPUs each rank was allowed to run on and the NUMA node of its GPU, because on these boxes the
limiter is the host side (root complexes / memory), not the GPUs.
"""
import argparse
import json
import os
import time

import torch
import torch.distributed as dist


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--mb", type=int, default=100)
    ap.add_argument("--reps", type=int, default=20)
    ap.add_argument("--out", default=None)
    a = ap.parse_args()
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    n = a.mb << 20
    h_a = torch.empty(n, dtype=torch.uint8).pin_memory()
    h_b = torch.empty(n, dtype=torch.uint8).pin_memory()
    h_a.fill_(1)
    h_b.fill_(2)
    d_a = torch.empty(n, dtype=torch.uint8, device=dev)
    d_b = torch.empty(n, dtype=torch.uint8, device=dev)
    s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()

    def h2d():
        with torch.cuda.stream(s1):
            d_a.copy_(h_a, non_blocking=True)

    def d2h():
        with torch.cuda.stream(s2):
            h_b.copy_(d_b, non_blocking=True)

    def both():
        h2d()
        d2h()

    def phase(fn, directions):
        fn()
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        for _ in range(a.reps):
            fn()
        torch.cuda.synchronize()
        dt = torch.tensor([time.perf_counter() - t0], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(dt, op=dist.ReduceOp.MAX)
        total = world * directions * n * a.reps / float(dt.item()) / 1e9
        return {"aggregate_gbs": total, "per_gpu_gbs": total / world, "per_gpu_per_direction_gbs": total / world / directions}

    res = {"n_gpus": world, "block_mb": a.mb, "reps": a.reps, "h2d": phase(h2d, 1), "d2h": phase(d2h, 1), "both": phase(both, 2)}
    try:
        cpus = sorted(os.sched_getaffinity(0))
        res_local = {"rank": rank, "cpus_allowed": len(cpus)}
        numa = f"/sys/bus/pci/devices/{torch.cuda.get_device_properties(local).pci_bus_id if hasattr(torch.cuda.get_device_properties(local), 'pci_bus_id') else ''}/numa_node"
        if os.path.exists(numa):
            res_local["gpu_numa_node"] = open(numa).read().strip()
    except Exception:  # noqa: BLE001
        res_local = {"rank": rank}
    if rank == 0:
        res["host_cpus"] = os.cpu_count()
        res["rank0"] = res_local
        # what bench.py's end-to-end arm needs per round-trip pixel: 15 B up + 15 B down
        res["e2e_ceiling_mpx_s"] = res["both"]["aggregate_gbs"] * 1e9 / 30.0 / 1e6
        line = json.dumps(res)
        print(line, flush=True)
        if a.out:
            with open(a.out, "w") as f:
                f.write(line + "\n")
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
