"""Multi-GPU sharding of the hot path: every sample is independent
(hand_search.cpp:172-182, image_generator.cpp:57-64,86-89, eigen_classifier.cpp:67-76), so the
sample-index array is cut into contiguous slices [r*n/R, (r+1)*n/R), one per rank, over a broadcast
copy of the cloud; the only exchange step is ONE all-gather of fixed-stride score slots
(n_samples x poses_per_sample float32, NaN = no candidate). torch.distributed is plumbing here:
backend "nccl" on GPUs (NVLink/NVSwitch), "gloo" in the CPU tests.
"""
import numpy as np


def slice_bounds(n, rank, world):
    """Contiguous slice of rank `rank` out of `world` over n samples."""
    return (rank * n) // world, ((rank + 1) * n) // world


def slot_stride(n, world):
    """Fixed per-rank slot count of the all-gather (largest slice)."""
    return max(slice_bounds(n, r, world)[1] - slice_bounds(n, r, world)[0] for r in range(world))


def gather_scores(local_scores, n, poses_per_sample, rank, world, dist=None, device=None):
    """All-gather the per-pose scores of every rank's slice into the full [n, P] array.

    local_scores: float32 array/tensor [n_local * P] (this rank's slice, NaN where no candidate).
    Returns a torch tensor [n, P] identical on all ranks.
    """
    import torch

    P = poses_per_sample
    lo, hi = slice_bounds(n, rank, world)
    t = torch.as_tensor(local_scores, dtype=torch.float32, device=device).reshape(-1)
    assert t.numel() == (hi - lo) * P, (t.numel(), hi - lo, P)
    if world == 1:
        return t.reshape(n, P)
    stride = slot_stride(n, world) * P
    slot = torch.full((stride,), float("nan"), dtype=torch.float32, device=t.device)
    slot[: t.numel()] = t
    out = torch.empty(world * stride, dtype=torch.float32, device=t.device)
    dist.all_gather_into_tensor(out, slot)
    full = torch.empty(n * P, dtype=torch.float32, device=t.device)
    for r in range(world):
        a, b = slice_bounds(n, r, world)
        full[a * P:b * P] = out[r * stride:r * stride + (b - a) * P]
    return full.reshape(n, P)


def select_global(local_candidates, num_selected, n, rank, world, dist=None, device=None):
    """Global selectGrasps (grasp_detector.cpp:405-420) over sharded samples: every rank contributes its local
    `num_selected` best pose records (gpdb_detect_select on its slice), ONE all-gather of fixed-stride record slots
    (num_selected x sizeof(gpdb_pose) bytes per rank; 17.6 KB for the default 100), then every rank merges them: descending
    score, ties in (rank, local order) = global candidate order, exactly like the single-GPU call. `sample_slot` is rebased
    from the slice to the full sample array. local_candidates: numpy structured array (abi.POSE_DTYPE), already sorted by
    descending score. Returns the global top records (numpy structured array), identical on all ranks."""
    import torch

    from . import abi

    k = int(num_selected)
    lo, _ = slice_bounds(n, rank, world)
    loc = np.array(local_candidates[:k], dtype=abi.POSE_DTYPE, copy=True)
    loc["sample_slot"] += lo
    if world == 1:
        return loc
    rec = abi.POSE_DTYPE.itemsize
    slot = np.zeros(k * rec + 8, np.uint8)
    slot[:8] = np.frombuffer(np.int64(len(loc)).tobytes(), np.uint8)
    slot[8:8 + len(loc) * rec] = np.frombuffer(loc.tobytes(), np.uint8)
    t = torch.from_numpy(slot).to(device) if device is not None else torch.from_numpy(slot)
    out = torch.empty(world * slot.size, dtype=torch.uint8, device=t.device)
    dist.all_gather_into_tensor(out, t)
    buf = out.cpu().numpy()
    parts = []
    for r in range(world):
        chunk = buf[r * slot.size:(r + 1) * slot.size]
        cnt = int(np.frombuffer(chunk[:8].tobytes(), np.int64)[0])
        parts.append(np.frombuffer(chunk[8:8 + cnt * rec].tobytes(), dtype=abi.POSE_DTYPE))
    allc = np.concatenate(parts) if parts else loc[:0]
    order = np.argsort(-allc["score"].astype(np.float64), kind="stable")[:k]
    return allc[order].copy()


def broadcast_cloud(cloud, rank, dist, device=None):
    """Broadcast the cloud arrays from rank 0 (ncclBroadcast over NVLink on GPUs)."""
    import torch

    out = {}
    for key in ("xyz", "normals", "cam_source", "view_points"):
        meta = [None]
        if rank == 0:
            meta = [(cloud[key].shape, str(cloud[key].dtype))]
        dist.broadcast_object_list(meta, src=0)
        shape, dtype = meta[0]
        if rank == 0:
            t = torch.from_numpy(np.ascontiguousarray(cloud[key])).to(device) if device else torch.from_numpy(
                np.ascontiguousarray(cloud[key]))
        else:
            t = torch.empty(shape, dtype=getattr(torch, dtype), device=device)
        dist.broadcast(t, 0)
        out[key] = t.cpu().numpy()
    return out
