"""CPU tests of the preprocessing restatement (oracle/gpd_oracle.cpp "Cloud preprocessing"): SURVEY.md 8(f).1,
CandidatesGenerator::preprocessPointCloud (candidates_generator.cpp:14-37).

PCL is absent from this image, so parity against upstream binaries is UNPINNED; what is pinned here: the committed
golden (regression), an independent float64 PCA of the same neighbourhoods, the closed-form float32 eigen-solver
against numpy.linalg.eigh, and the voxel-set semantics against numpy.unique.
"""
import os

import numpy as np

from gpd_b200 import abi, scenes
from oracle import oracle


def test_krylon_preprocess_golden(golden_dir):
    g = np.load(os.path.join(golden_dir, "krylon_preprocess.npz"))
    r = oracle.preprocess(g["raw"], None, np.zeros((1, 3)), abi.default_preprocess_params())
    assert np.array_equal(r["xyz"], g["xyz"]) and np.array_equal(r["src"], g["src"])
    assert np.array_equal(r["cam_source"], g["cam_source"])
    # regression pin; libm's float atan2f / cosf / sinf may differ between hosts in the last bit
    assert np.abs(r["normals"] - g["normals"]).max() <= 1e-6
    # independent check: float64 PCA of the same balls (sign-free)
    c = np.abs((r["normals"] * g["pca64"]).sum(1))
    assert (1.0 - c).max() < 2e-4
    # viewpoint flip + reverseNormals: every normal points towards the camera at the origin
    assert ((r["xyz"].astype(np.float64) * r["normals"]).sum(1) < 0).all()
    # and the committed voxelised fixture used by the path tests is the same point set
    k = np.load(os.path.join(golden_dir, "krylon_voxel.npz"))
    assert set(map(tuple, k["xyz"])) == set(map(tuple, r["xyz"]))


def test_pcl_eigen33_is_the_smallest_eigenpair():
    rng = np.random.default_rng(0)
    for _ in range(200):
        a = rng.standard_normal((3, 8)) * rng.uniform(1e-3, 1.0, (3, 1))
        cov = (a @ a.T / 8).astype(np.float32)
        ev, vec = oracle.pcl_eigen33(cov)
        w, v = np.linalg.eigh(cov.astype(np.float64))
        scale = np.abs(cov).max()
        assert abs(ev - w[0]) <= 2e-5 * scale
        gap = (w[1] - w[0]) / scale
        if gap > 1e-2:
            assert 1.0 - abs(float(vec @ v[:, 0])) < 1e-3 / gap
        assert abs(np.linalg.norm(vec) - 1) < 1e-5


def test_voxelise_is_an_exact_set_in_reverse_first_occurrence_order():
    rng = np.random.default_rng(5)
    pts = rng.uniform(-0.05, 0.05, (5000, 3)).astype(np.float32)
    pts[100] = [np.nan, 0, 0]                      # removeNans
    pts[200] = [0.9, 0, 0]                         # outside the workspace below
    nrm = rng.standard_normal((5000, 3))
    cam = np.zeros((5000, 2), np.int32)
    cam[np.arange(5000), rng.integers(0, 2, 5000)] = 1
    pp = abi.default_preprocess_params(workspace=[-0.5, 0.5, -0.5, 0.5, -0.5, 0.5], voxel_size=0.01, estimate_normals=0)
    r = oracle.preprocess(pts, cam, np.zeros((2, 3)), pp, normals=nrm)
    keep = np.array([i for i in range(5000) if i not in (100, 200)])
    mn = pts[keep].min(0)
    vox = np.floor((pts[keep] - mn) / np.float32(0.01)).astype(np.int64)
    uniq, first, inv = np.unique(vox, axis=0, return_index=True, return_inverse=True)
    order = np.argsort(-first)                      # newest first
    assert np.array_equal(r["src"], keep[first[order]])
    assert np.array_equal(r["xyz"], (mn + np.float32(0.01) * uniq[order].astype(np.float32)).astype(np.float32))
    assert np.array_equal(r["cam_source"], cam[r["src"]])
    inv = inv.ravel()
    for o in (0, 7, len(order) - 1):               # voxel-averaged normals, summed in index order
        members = keep[np.nonzero(inv == order[o])[0]]
        acc = np.zeros(3)
        for i in members:
            acc += nrm[i]
        assert np.array_equal(r["normals"][o], acc / len(members))
    # no voxelisation: plain order-preserving filter
    pp2 = abi.default_preprocess_params(workspace=[-0.5, 0.5, -0.5, 0.5, -0.5, 0.5], voxelize=0, estimate_normals=0)
    r2 = oracle.preprocess(pts, cam, np.zeros((2, 3)), pp2, normals=nrm)
    assert np.array_equal(r2["src"], keep) and np.array_equal(r2["xyz"], pts[keep]) and np.array_equal(r2["normals"], nrm[keep])


def test_normals_two_cameras_and_degenerate_neighbourhoods():
    s = scenes.synthetic_raw_scene(7, n_points=20000, two_cameras=True, nan_fraction=0.01)
    r = oracle.preprocess(s["xyz"], s["cam_source"], s["view_points"], abi.default_preprocess_params())
    n = r["normals"]
    assert not np.isnan(r["xyz"]).any() and len(r["xyz"]) < len(s["xyz"])
    ok = ~np.isnan(n).any(1)
    assert ok.mean() > 0.999 and np.abs(np.linalg.norm(n[ok], axis=1) - 1).max() < 1e-5
    # reverseNormals: the normal points towards at least one camera that sees the point
    d = r["xyz"].astype(np.float64)[:, None, :] - s["view_points"][None]
    toward = ((d * n[:, None, :]).sum(2) < 0) & (r["cam_source"] == 1)
    assert toward.any(1)[ok].all()
    # isolated points (< 3 neighbours) get NaN normals like pcl::computePointNormal
    iso = np.array([[0.3, 0.3, 0.3], [0.0, 0.0, 0.5], [0.001, 0.0, 0.5], [0.0, 0.001, 0.5], [0.001, 0.001, 0.5]], np.float32)
    r3 = oracle.preprocess(iso, None, np.zeros((1, 3)), abi.default_preprocess_params(voxelize=0))
    assert np.isnan(r3["normals"][0]).all() and not np.isnan(r3["normals"][1:]).any()


def test_literal_std_set_emulation_quantifies_the_upstream_quirk(golden_dir):
    """The reference's voxel set uses a comparator that is not a strict weak ordering (cloud.h:105-122); emulating
    libstdc++'s red-black tree literally (oracle RbSim) shows what upstream really computes: duplicates are only
    recognised on the tree's left spine. The exact-set variant (product + oracle) equals the literal result whenever
    no duplicate is missed, order included; on the tutorial cloud upstream keeps 993 duplicate voxels."""
    raw = np.load(os.path.join(golden_dir, "krylon_preprocess.npz"))["raw"]
    pp = abi.default_preprocess_params(estimate_normals=0)
    zeros = np.zeros((len(raw), 3))
    exact = oracle.preprocess(raw, None, np.zeros((1, 3)), pp, normals=zeros)
    src, missed = oracle.voxelize_literal(raw)
    assert len(exact["src"]) == 2373 and len(src) == 2373 + missed and missed == 993
    # every literal element is a voxel of the exact set, and every exact voxel appears
    mn = raw.min(0)
    vox = lambda idx: set(map(tuple, np.floor((raw[idx] - mn) / np.float32(0.003)).astype(np.int64)))
    assert vox(src) == vox(exact["src"])
    # duplicates adjacent in the input (voxel-sorted scan): nothing is missed and the two agree, order included
    v = np.floor((raw - mn) / np.float32(0.003)).astype(np.int64)
    rs = raw[np.lexsort((v[:, 2], v[:, 1], v[:, 0]))]
    src2, missed2 = oracle.voxelize_literal(rs)
    ex2 = oracle.preprocess(rs, None, np.zeros((1, 3)), pp, normals=zeros)
    assert missed2 == 0 and np.array_equal(src2, ex2["src"])


def test_normals_against_a_float32_numpy_restatement(golden_dir):
    """Second restatement of pcl::NormalEstimation (PCL 1.9.1: computeMeanAndCovarianceMatrix float32 single pass in the
    kd-tree's (dist, index) order, solvePlaneParameters / pcl::eigen33 closed form, flipNormalTowardsViewpoint) with numpy
    float32 scalars, written separately from the C++ oracle: the float32 sums must agree bit for bit (same order, no
    FMA), the normals to ~1e-6 (numpy's float32 atan2 / cos / sin vs glibc's)."""
    f32 = np.float32
    g = np.load(os.path.join(golden_dir, "krylon_preprocess.npz"))
    xyz, ref = g["xyz"], g["normals"]
    oc = oracle.OracleCloud(xyz, np.zeros((len(xyz), 3)), None, np.zeros((1, 3)))

    def roots2(b, c):
        d = f32(float(f32(b * b)) - 4.0 * float(c))
        d = f32(0.0) if d < 0 else d
        sd = f32(np.sqrt(d))
        return [f32(0.0), f32(0.5) * (b - sd), f32(0.5) * (b + sd)]

    def roots(m):
        c0 = (m[0][0] * m[1][1] * m[2][2] + f32(2) * m[0][1] * m[0][2] * m[1][2] - m[0][0] * m[1][2] * m[1][2]
              - m[1][1] * m[0][2] * m[0][2] - m[2][2] * m[0][1] * m[0][1])
        c1 = m[0][0] * m[1][1] - m[0][1] * m[0][1] + m[0][0] * m[2][2] - m[0][2] * m[0][2] + m[1][1] * m[2][2] - m[1][2] * m[1][2]
        c2 = m[0][0] + m[1][1] + m[2][2]
        if abs(c0) < np.finfo(f32).eps:
            return roots2(c2, c1)
        inv3, sqrt3 = f32(1.0 / 3.0), f32(np.sqrt(f32(3.0)))
        c2o3 = c2 * inv3
        a3 = (c1 - c2 * c2o3) * inv3
        a3 = f32(0) if a3 > 0 else a3
        hb = f32(0.5) * (c0 + c2o3 * (f32(2) * c2o3 * c2o3 - c1))
        q = hb * hb + a3 * a3 * a3
        q = f32(0) if q > 0 else q
        rho = f32(np.sqrt(-a3))
        theta = f32(np.arctan2(f32(np.sqrt(-q)), hb)) * inv3
        ct, st = f32(np.cos(theta)), f32(np.sin(theta))
        r = [c2o3 + f32(2) * rho * ct, c2o3 - rho * (ct + sqrt3 * st), c2o3 - rho * (ct - sqrt3 * st)]
        if r[0] >= r[1]:
            r[0], r[1] = r[1], r[0]
        if r[1] >= r[2]:
            r[1], r[2] = r[2], r[1]
            if r[0] >= r[1]:
                r[0], r[1] = r[1], r[0]
        return roots2(c2, c1) if r[0] <= 0 else r

    checked, worst = 0, 0.0
    for i in range(0, len(xyz), 37):
        idx, _ = oc.radius_search(xyz[i], 0.03)
        acc = [f32(0)] * 9
        for j in idx:                                        # sorted (dist, index) order
            x, y, z = xyz[j]
            for k, v in enumerate((x * x, x * y, x * z, y * y, y * z, z * z, x, y, z)):
                acc[k] = f32(acc[k] + v)
        acc = [a / f32(len(idx)) for a in acc]
        cov = [[acc[0] - acc[6] * acc[6], acc[1] - acc[6] * acc[7], acc[2] - acc[6] * acc[8]],
               [None, acc[3] - acc[7] * acc[7], acc[4] - acc[7] * acc[8]], [None, None, acc[5] - acc[8] * acc[8]]]
        cov[1][0], cov[2][0], cov[2][1] = cov[0][1], cov[0][2], cov[1][2]
        scale = max(abs(v) for r in cov for v in r)
        sm = [[v / scale for v in r] for r in cov]
        ev = roots(sm)
        for d in range(3):
            sm[d][d] = sm[d][d] - ev[0]
        cr = lambda a, b: [a[1] * b[2] - a[2] * b[1], a[2] * b[0] - a[0] * b[2], a[0] * b[1] - a[1] * b[0]]
        vs = [cr(sm[0], sm[1]), cr(sm[0], sm[2]), cr(sm[1], sm[2])]
        ln = [v[0] * v[0] + v[1] * v[1] + v[2] * v[2] for v in vs]
        b = 0 if (ln[0] >= ln[1] and ln[0] >= ln[2]) else (1 if (ln[1] >= ln[0] and ln[1] >= ln[2]) else 2)
        n = np.array([c / f32(np.sqrt(ln[b])) for c in vs[b]], f32)
        vpv = f32(0) - xyz[i]                                # view point (0,0,0) - point, float32
        if vpv[0] * n[0] + vpv[1] * n[1] + vpv[2] * n[2] < 0:
            n = -n
        nd = n.astype(np.float64)
        if nd @ xyz[i].astype(np.float64) >= 0:              # reverseNormals: must point towards the camera at the origin
            nd = -nd
        worst = max(worst, float(np.abs(nd - ref[i]).max()))
        checked += 1
    assert checked >= 60 and worst <= 2e-6, (checked, worst)
