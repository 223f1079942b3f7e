#!/usr/bin/env python
"""Multi-GPU parity check of the in-library sharding (development aid, run under torchrun on N GPUs):

  python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port 29511 tools/multi_gpu_check.py

Every rank: gpdb_comm_init -> gpdb_set_cloud_bcast (cloud from rank 0 only) -> gpdb_detect_sharded over the SAME sample
array; then every rank recomputes ALL samples alone (gpdb_detect on its own copy of the broadcast cloud) and compares the
all-gathered flags / scores bit for bit, plus its own slice's pose records. Prints one PASS / FAIL line per rank."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402

import bench  # noqa: E402
from gpd_b200 import lib  # noqa: E402


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    n_total = int(sys.argv[1]) if len(sys.argv) > 1 else 20001  # not divisible by 2 / 4 / 8: short last slots
    ok = True
    for config in (3, 5):
        cfg = bench.CONFIGS[config]
        cloud = bench.scenes.synthetic_table_scene(cfg["seed"], two_cameras=cfg["two_cameras"]) if rank == 0 else None
        sidx = np.random.default_rng(11).integers(0, 300000, n_total).astype(np.int32)
        p = lib.default_params(channels=cfg["channels"], relu_after_conv=cfg["relu"], device=local)
        ctx = lib.Context(p)
        ctx.set_weights(bench.load_weights(cfg["channels"]))
        uid = [lib.comm_unique_id() if rank == 0 else None]
        dist.broadcast_object_list(uid, src=0)
        ctx.comm_init(uid[0], rank, world)
        if rank == 0:
            n = ctx.set_cloud_bcast(0, cloud["xyz"], cloud["normals"], cloud["cam_source"], cloud["view_points"])
        else:
            n = ctx.set_cloud_bcast(0)
        sh = ctx.detect_sharded(sidx)
        one = ctx.detect(sidx)  # every rank alone, on the cloud it RECEIVED
        lo, hi, _ = lib.shard_bounds(n_total, rank, world)
        mine = one["candidates"][(one["candidates"]["sample_slot"] >= lo) & (one["candidates"]["sample_slot"] < hi)]
        checks = {
            "cloud_points": n == 300000,
            "flags": np.array_equal(sh["pose_flags"], one["pose_flags"]),
            "scores": np.array_equal(sh["pose_scores"].view(np.uint32), one["pose_scores"].view(np.uint32)),
            "total": sh["n_total_candidates"] == one["n_candidates"],
            "records": sh["candidates"].tobytes() == mine.tobytes(),
        }
        good = all(checks.values())
        ok = ok and good
        print(f"[multi_gpu_check] config {config} rank {rank}/{world}: {'PASS' if good else 'FAIL'} {checks} "
              f"({sh['n_candidates']} local of {sh['n_total_candidates']} candidates)", flush=True)
        ctx.close()
    t = torch.tensor([0 if ok else 1], device="cuda")
    dist.all_reduce(t)
    dist.barrier()
    dist.destroy_process_group()
    sys.exit(int(t.item() != 0))


if __name__ == "__main__":
    main()
