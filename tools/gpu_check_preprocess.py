#!/usr/bin/env python
"""GPU-vs-oracle comparison of gpdb_preprocess with verbose statistics and timings (development aid; the asserting
version is tests/test_gpu_preprocess.py)."""
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from gpd_b200 import lib, scenes  # noqa: E402
from oracle import oracle  # noqa: E402


def compare(name, s, pp, normals=None):
    ctx = lib.Context(lib.default_params(channels=15))
    t = time.time()
    rg = ctx.preprocess(s["xyz"], s.get("cam_source"), s["view_points"], pp, normals=normals)
    tg = time.time() - t
    t = time.time()
    n2 = ctx.preprocess(s["xyz"], s.get("cam_source"), s["view_points"], pp, normals=normals, read_back=False)
    tg2 = time.time() - t
    ms = ctx.preprocess_timings()
    t = time.time()
    ro = oracle.preprocess(s["xyz"], s.get("cam_source"), s["view_points"], pp, normals=normals)
    to = time.time() - t
    print(f"=== {name}: raw {len(s['xyz'])} -> gpu {len(rg['xyz'])} / oracle {len(ro['xyz'])} points; gpu wall {tg:.3f}s, "
          f"second call {tg2:.3f}s (N'={n2}); device ms upload/filter/voxel/grid/normals/total = {np.round(ms, 3)}; "
          f"oracle {to:.3f}s {ro['seconds']} on {oracle.num_threads()} threads")
    if len(rg["xyz"]) != len(ro["xyz"]):
        return
    print("   src equal", np.array_equal(ro["src"], rg["src"]), "xyz bit-equal", np.array_equal(ro["xyz"], rg["xyz"]),
          "cam equal", np.array_equal(ro["cam_source"], rg["cam_source"]))
    no, ng = ro["normals"], rg["normals"]
    print("   nan pattern equal", np.array_equal(np.isnan(no), np.isnan(ng)), "nan rows", int(np.isnan(ng).any(1).sum()))
    d = np.abs(np.nan_to_num(no) - np.nan_to_num(ng)).max(1)
    print(f"   normals: max abs diff {d.max():.3e}, bit-equal rows {np.mean(d == 0):.5f}, rows > 1e-6: {int((d > 1e-6).sum())}, "
          f"sign flips {int((np.nan_to_num((no * ng).sum(1)) < 0).sum())}")
    if (d > 1e-6).any():
        for i in np.argsort(-d)[:5]:
            print("     worst", i, no[i], ng[i])
    ctx.close()


if __name__ == "__main__":
    g = np.load(os.path.join(ROOT, "tests", "golden", "krylon_preprocess.npz"))
    compare("krylon raw", {"xyz": g["raw"], "view_points": np.zeros((1, 3))}, lib.preprocess_params())
    s = scenes.synthetic_raw_scene(7, n_points=60000, two_cameras=True, nan_fraction=0.01)
    compare("two-view 60k-scale raw", s, lib.preprocess_params(workspace=[-0.6, 0.6, -0.5, 0.5, 0.2, 1.0]))
    s = scenes.synthetic_raw_scene(3)
    compare("config-3-size raw", s, lib.preprocess_params())
