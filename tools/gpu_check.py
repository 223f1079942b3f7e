#!/usr/bin/env python
"""Stage-by-stage GPU-vs-oracle comparison with verbose mismatch statistics (development aid;
the asserting version lives in tests/test_gpu_parity.py)."""
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from gpd_b200 import abi, lib, scenes  # noqa: E402
from oracle import oracle  # noqa: E402


def weights_for(ch):
    d = np.load(os.path.join(ROOT, "gpd_b200", "weights", f"lenet_{ch}ch.npz"))
    return [d[n] for n in oracle.WeightPack.NAMES], int(d["relu_after_conv"])


def compare(name, cloud, sidx, ch, **over):
    print(f"=== {name}: N={len(cloud['xyz'])} samples={len(sidx)} channels={ch} {over}")
    w, relu = weights_for(ch)
    p = lib.default_params(channels=ch, relu_after_conv=relu, keep_images=1, **over)
    oc = oracle.OracleCloud(cloud["xyz"], cloud["normals"], cloud["cam_source"], cloud["view_points"])
    t = time.time()
    ro = oc.detect(p, oracle.WeightPack(w), sidx)
    t_or = time.time() - t
    ctx = lib.Context(p)
    ctx.set_weights(w)
    ctx.set_cloud(cloud["xyz"], cloud["normals"], cloud["cam_source"], cloud["view_points"])
    t = time.time()
    rg = ctx.detect(sidx)
    t_g = time.time() - t
    t = time.time()
    rg = ctx.detect(sidx)
    t_g2 = time.time() - t
    print(f"oracle {t_or:.3f}s  gpu first {t_g:.3f}s second {t_g2:.3f}s  stage ms {rg['ms']} launches {rg['kernel_launches']}")
    print("frames: valid equal", np.array_equal(ro["frame_valid"], rg["frame_valid"]), "max abs diff",
          np.abs(ro["frames"] - rg["frames"]).max(), "bit-equal", np.array_equal(ro["frames"], rg["frames"]))
    fe = ro["pose_flags"] == rg["pose_flags"]
    print("flags equal:", fe.mean(), "mismatches", (~fe).sum(), "candidates", ro["n_candidates"], rg["n_candidates"])
    if not fe.all():
        bad = np.argwhere(~fe)[:10]
        for b in bad:
            print("   flag mismatch at", b, ro["pose_flags"][tuple(b)], rg["pose_flags"][tuple(b)])
    if ro["n_candidates"] == rg["n_candidates"] and ro["n_candidates"]:
        co, cg = ro["candidates"], rg["candidates"]
        for f in ("sample", "frame", "position", "top", "bottom", "center", "width"):
            print(f"   cand.{f}: max abs diff {np.abs(co[f] - cg[f]).max():.3e} bit-equal {np.array_equal(co[f], cg[f])}")
        for f in ("sample_index", "sample_slot", "pose_slot", "finger_idx", "half_antipodal", "full_antipodal"):
            print(f"   cand.{f}: equal {np.array_equal(co[f], cg[f])}")
        io, ig = ro["images"].astype(np.int32), rg["images"].astype(np.int32)
        d = np.abs(io - ig)
        print(f"   images: differing pixels {np.count_nonzero(d)} of {d.size} ({np.count_nonzero(d) / d.size:.2e}), max diff {d.max()}")
        dd = d.reshape(len(d), -1, ch)
        print("   per-channel differing pixel counts:", np.count_nonzero(dd, axis=(0, 1)))
        print("   per-channel max diff:", dd.max(axis=(0, 1)))
        so, sg = co["score"], cg["score"]
        scale = np.abs(so).max()
        print(f"   scores: max abs diff {np.abs(so - sg).max():.4e} rel-to-max {np.abs(so - sg).max() / scale:.3e}")
        # classifier alone on identical (oracle) images
        sc_g, lg_g = ctx.classify(ro["images"])
        sc_o, lg_o = oracle.classify(p, oracle.WeightPack(w), ro["images"])
        print(f"   classify(oracle images): max rel logit diff {np.abs(lg_g - lg_o).max() / np.abs(lg_o).max():.3e}")
        # images alone on identical (oracle) poses
        ig2 = ctx.images(co).reshape(io.shape).astype(np.int32)
        d2 = np.abs(io - ig2)
        print(f"   images(oracle poses): differing pixels {np.count_nonzero(d2)} max {d2.max()}")
    ctx.close()
    return ro, rg


if __name__ == "__main__":
    print(lib.lib().gpdb_build_info().decode())
    k = scenes.krylon_cloud()
    compare("krylon 15ch", k, scenes.sample_indices(2, len(k["xyz"]), 300), 15)
    compare("krylon 3ch", k, scenes.sample_indices(1, len(k["xyz"]), 100), 3)
    s = scenes.synthetic_table_scene(7, n_points=60000)
    compare("synthetic 60k 15ch", s, scenes.sample_indices(3, 60000, 300), 15)
    s2 = scenes.synthetic_table_scene(5, n_points=60000, two_cameras=True)
    compare("synthetic two-view 12ch", s2, scenes.sample_indices(5, 60000, 200), 12)
    compare("synthetic two-view 15ch all axes", s2, scenes.sample_indices(5, 60000, 60), 15, hand_axes=[0, 1, 2],
            num_orientations=4)
