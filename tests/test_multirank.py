"""World-size-2 host logic on CPU (gloo): the quantizer broadcast and the frame sharding bench.py uses for N > 1.

The data path has no collective (frames are independent); what ranks exchange is the packed
quantizer built by rank 0's libm.  Each rank then checks, with the oracle, that quantising ITS shard of the
frame stream with the received LUT gives the planes a single process would have produced for those frames."""
import os
import socket
import sys
from pathlib import Path

import numpy as np
import pytest

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _rank_main(rank, world, port, n_frames, out_dir):
    import torch
    import torch.distributed as dist

    import lumahdrv_b200 as L
    from lumahdrv_b200.shard import broadcast_quantizer, frame_shard, pack_quantizer, packed_quantizer_size, unpack_quantizer
    from oracle import pyoracle as po

    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        bits, cbits = 11, 8
        vec = None
        if rank == 0:
            lut = L.build_lut("PQ", bits, 1e4, 0.005)  # host libm, reference formula (lumacu_build_lut)
            vec = pack_quantizer(lut, (1 << cbits) - 1, L.CS_LUV, 1e4, 0.005, 1.0, 2, ptf=L.PTF_PQ)
        got = unpack_quantizer(broadcast_quantizer(vec, packed_quantizer_size(bits), torch.device("cpu"), src=0))
        assert got["ptf"] == L.PTF_PQ and got["ptf_bit_depth"] == bits and got["color_bit_depth"] == cbits
        assert got["max_val_color"] == 255 and got["color_space"] == L.CS_LUV and got["profile"] == 2
        assert got["max_lum"] == 1e4 and abs(got["min_lum"] - 0.005) < 1e-9 and got["pre_scaling"] == 1.0

        o = po.Oracle().setQuantizer("PQ", bits, "LUV", cbits)
        assert np.array_equal(np.asarray(o.getMapping(), dtype=np.float32).view(np.uint32), got["lut"].view(np.uint32))
        # this rank's shard, quantised with the RECEIVED table
        o.setMapping(got["lut"])
        mine = frame_shard(n_frames, rank, world)
        hashes = {}
        for f in mine:
            frame = po.noise_frame(64, 32, seed=0x9E3779B97F4A7C15 + f)
            planes, _ = o.encode(frame, 2, 1.0)
            hashes[f] = [int(h) for h in po.plane_hashes(planes, 64, 32, 2)]
        np.save(os.path.join(out_dir, f"rank{rank}.npy"), np.array(sorted(hashes.items()), dtype=object), allow_pickle=True)
        dist.barrier()
    finally:
        dist.destroy_process_group()


def test_world2_quantizer_broadcast_and_frame_shards(tmp_path, po):
    import torch.multiprocessing as mp

    if not hasattr(po.Oracle, "setMapping"):
        pytest.skip("oracle has no setMapping")
    world, n_frames = 2, 5
    port = _free_port()
    mp.start_processes(_rank_main, args=(world, port, n_frames, str(tmp_path)), nprocs=world, join=True, start_method="spawn")
    seen = {}
    for r in range(world):
        for f, h in np.load(tmp_path / f"rank{r}.npy", allow_pickle=True):
            assert f not in seen, "frame processed by two ranks"
            seen[f] = h
    assert sorted(seen) == list(range(n_frames))
    # single-process truth
    o = po.Oracle().setQuantizer("PQ", 11, "LUV", 8)
    for f in range(n_frames):
        planes, _ = o.encode(po.noise_frame(64, 32, seed=0x9E3779B97F4A7C15 + f), 2, 1.0)
        assert [int(h) for h in po.plane_hashes(planes, 64, 32, 2)] == seen[f]


def _parity_rank_main(rank, world, port, out_dir):
    """bench.py's parity helpers over gloo: each rank checks one frame (rank 1's carries one corrupted plane byte and a
    1-ulp float error); the reduced record must carry the SUM of mismatches and the MAX ulp on every rank."""
    import json

    import torch
    import torch.distributed as dist

    import bench
    from oracle import pyoracle as po

    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        w, h = 128, 64
        frame = po.noise_frame(w, h, seed=17 + rank)
        o = po.Oracle().setQuantizer("LOG", 12, "LUV", 8)
        fc = frame.copy()
        planes, _ = o.encode(fc, 2, 1.0)
        out = o.decode(planes, w, h, 2, 1.0)
        if rank == 1:
            planes[1][3, 8] ^= 0x40
            out[2, 5, 5] = np.nextafter(out[2, 5, 5], np.float32(np.inf))
        prm = {"w": w, "h": h, "profile": 2, "ptf": "LOG", "ptfBitDepth": 12, "colorSpace": "LUV", "colorBitDepth": 8,
               "preScaling": 1.0, "maxLum": 1e4, "minLum": 0.005,
               "stats": {"sum": float(fc[0].astype(np.float64).sum()), "max": float(fc[0].max()), "min": float(fc[0].min())}}
        local = bench.run_parity_check(prm, frame, planes, out)
        red = bench.parity_over_ranks(local, torch.device("cpu"), world)
        with open(os.path.join(out_dir, f"parity{rank}.json"), "w") as f:
            json.dump({"local": local, "reduced": red}, f)
        dist.barrier()
    finally:
        dist.destroy_process_group()


def test_world2_bench_parity_reduction(tmp_path, po):
    import json

    import torch.multiprocessing as mp

    world = 2
    mp.start_processes(_parity_rank_main, args=(world, _free_port(), str(tmp_path)), nprocs=world, join=True, start_method="spawn")
    recs = [json.loads((tmp_path / f"parity{r}.json").read_text()) for r in range(world)]
    assert recs[0]["local"]["plane_mismatch_bytes"] == 0 and recs[0]["local"]["max_ulp"] == 0
    assert recs[1]["local"]["plane_mismatch_bytes"] == 1 and recs[1]["local"]["max_ulp"] == 1
    for r in recs:
        red = r["reduced"]
        assert red["plane_mismatch_bytes"] == 1 and red["max_ulp"] == 1 and red["frames_checked"] == 2
        assert red["ranks"] == 2 and red["ranks_failed"] == 0 and red["pixels_checked"] == 2 * 128 * 64
        assert red["stats_max_mismatches"] == 0 and red["stats_sum_max_rel_err"] < 1e-12


def test_bench_arms_share_one_config_object():
    """the driver compares the `config` of the two arms key by key"""
    import bench
    a = bench.config_dict(4, 32)
    assert a == bench.config_dict(4, 32) and a["workload"] == bench.WORKLOAD
    assert set(a) == {"workload", "frames_per_gpu_per_step", "input", "l2", "parallelism"}
