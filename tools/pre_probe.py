#!/usr/bin/env python
"""Runs gpdb_preprocess on the config-3-size raw cloud a few times (target of the ncu captures of k_normals)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from gpd_b200 import lib, scenes
raw = scenes.synthetic_raw_scene(3)
ctx = lib.Context(lib.default_params(channels=15))
for _ in range(int(sys.argv[1]) if len(sys.argv) > 1 else 2):
    n = ctx.preprocess(raw["xyz"], raw["cam_source"], raw["view_points"], lib.preprocess_params(), read_back=False)
print("processed", n, "device ms", ctx.preprocess_timings())
