"""Seeded INPUT generators for the hot path (host-side, numpy): the tutorial cloud fixture and the
synthetic "cluttered table" clouds of BASELINE.json's configs.

This is input preparation (what util::Cloud preprocessing produces before the path starts,
candidates_generator.cpp:14-37) — NOT part of the accelerated path and not timed.
Outputs follow the C-ABI of gpdb_set_cloud: xyz float32 [N,3], normals float64 [N,3] (float32
values, as PCL normals are, cloud.cpp:531-532), cam_source int32 [N,K], view_points float64 [K,3].
"""
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
GOLDEN_DIR = os.path.join(os.path.dirname(_HERE), "tests", "golden")


def voxelize(points, cell=0.003):
    """Cloud::voxelizeCloud (cloud.cpp:286-348): voxel corner min_pt + cell*floor((p-min)/cell),
    one point per occupied voxel, ordered lexicographically by voxel index."""
    pts = np.asarray(points, np.float32)
    mn = pts.min(axis=0)
    vox = np.floor((pts - mn) / np.float32(cell)).astype(np.int64)
    _, first = np.unique(vox, axis=0, return_index=True)
    v = vox[first]
    order = np.lexsort((v[:, 2], v[:, 1], v[:, 0]))
    first = first[order]
    out = (mn + np.float32(cell) * vox[first].astype(np.float32)).astype(np.float32)
    return out, first


def pca_normals(points, radius, view_point, flip_away=False):
    """pcl::NormalEstimation restated loosely (PCA of the r-ball, flipped towards the view
    point); only used to prepare the small tutorial fixture."""
    from scipy.spatial import cKDTree

    pts = np.asarray(points, np.float64)
    tree = cKDTree(pts)
    nn = tree.query_ball_point(pts, radius)
    normals = np.zeros_like(pts)
    for i, idx in enumerate(nn):
        q = pts[idx]
        c = np.cov((q - q.mean(0)).T) if len(idx) >= 3 else np.eye(3)
        w, v = np.linalg.eigh(c)
        n = v[:, 0]
        if np.dot(n, view_point - pts[i]) < 0:
            n = -n
        normals[i] = n
    if flip_away:
        normals = -normals
    return normals.astype(np.float32).astype(np.float64)


def load_pcd_ascii(path):
    """Minimal ASCII PCD reader (x y z [rgb]) for tutorials/*.pcd."""
    with open(path) as f:
        lines = f.read().split("\n")
    k = next(i for i, l in enumerate(lines) if l.startswith("DATA"))
    rows = [l.split()[:3] for l in lines[k + 1:] if l.strip()]
    a = np.array(rows, dtype=np.float64)
    return a[np.isfinite(a).all(axis=1)].astype(np.float32)


def krylon_cloud(pcd_path=None):
    """Config 1/2 cloud: tutorials/krylon.pcd voxelised at 0.003 (2 373 points), normals PCA
    r=0.03 towards camera_position = origin then negated (test_grasp_image.cpp:117-118: the
    object surrounds the origin). Loaded from the committed fixture when present."""
    fx = os.path.join(GOLDEN_DIR, "krylon_voxel.npz")
    if pcd_path is None and os.path.exists(fx):
        d = np.load(fx)
        return {k: d[k] for k in ("xyz", "normals", "cam_source", "view_points")}
    pcd_path = pcd_path or "/root/reference/tutorials/krylon.pcd"
    pts = load_pcd_ascii(pcd_path)
    xyz, _ = voxelize(pts, 0.003)
    vp = np.zeros((1, 3))
    normals = pca_normals(xyz, 0.03, vp[0], flip_away=True)
    return {"xyz": xyz, "normals": normals, "cam_source": np.ones((len(xyz), 1), np.int32), "view_points": vp}


# ------------------------------------------------------------------------------------------------
# synthetic cluttered table (configs 3-5)
# ------------------------------------------------------------------------------------------------
def _lattice_rect(origin, eu, ev, lu, lv, step):
    nu, nv = max(int(lu / step), 1), max(int(lv / step), 1)
    u, v = np.meshgrid(np.arange(nu) * step, np.arange(nv) * step, indexing="ij")
    return origin + u.reshape(-1, 1) * eu + v.reshape(-1, 1) * ev


def _box(rng, cx, cy, z0, step):
    sx, sy, sz = rng.uniform(0.05, 0.25, 3)
    yaw = rng.uniform(0, np.pi)
    ex = np.array([np.cos(yaw), np.sin(yaw), 0.0])
    ey = np.array([-np.sin(yaw), np.cos(yaw), 0.0])
    ez = np.array([0.0, 0.0, -1.0])  # "up" is -z: camera looks along +z, table at larger z
    c = np.array([cx, cy, z0])
    P, Nn = [], []
    faces = [
        (c - ex * sx / 2 - ey * sy / 2 + ez * sz, ex, ey, sx, sy, ez),  # top
        (c - ex * sx / 2 - ey * sy / 2, ex, ez, sx, sz, -ey),
        (c - ex * sx / 2 + ey * sy / 2, ex, ez, sx, sz, ey),
        (c - ex * sx / 2 - ey * sy / 2, ey, ez, sy, sz, -ex),
        (c + ex * sx / 2 - ey * sy / 2, ey, ez, sy, sz, ex),
    ]
    for o, eu, ev, lu, lv, n in faces:
        p = _lattice_rect(o, eu, ev, lu, lv, step)
        P.append(p)
        Nn.append(np.tile(n, (len(p), 1)))
    return np.vstack(P), np.vstack(Nn)


def _cylinder(rng, cx, cy, z0, step):
    r, h = rng.uniform(0.025, 0.09), rng.uniform(0.05, 0.25)
    nth = max(int(2 * np.pi * r / step), 8)
    th = np.arange(nth) * (2 * np.pi / nth)
    zz = np.arange(max(int(h / step), 1)) * step
    T, Z = np.meshgrid(th, zz, indexing="ij")
    side = np.stack([cx + r * np.cos(T), cy + r * np.sin(T), z0 - Z], -1).reshape(-1, 3)
    nside = np.stack([np.cos(T), np.sin(T), np.zeros_like(T)], -1).reshape(-1, 3)
    g = np.arange(-r, r, step)
    X, Y = np.meshgrid(g, g, indexing="ij")
    m = X * X + Y * Y < r * r
    top = np.stack([cx + X[m], cy + Y[m], np.full(m.sum(), z0 - h)], -1)
    ntop = np.tile([0.0, 0.0, -1.0], (len(top), 1))
    return np.vstack([side, top]), np.vstack([nside, ntop])


def _sphere(rng, cx, cy, z0, step):
    r = rng.uniform(0.03, 0.10)
    n = max(int(4 * np.pi * r * r / (step * step)), 16)
    i = np.arange(n) + 0.5
    phi = np.arccos(1 - 2 * i / n)
    th = np.pi * (1 + 5 ** 0.5) * i
    d = np.stack([np.cos(th) * np.sin(phi), np.sin(th) * np.sin(phi), np.cos(phi)], -1)
    return np.array([cx, cy, z0 - r]) + r * d, d


def _visible(points, cam, ang_res):
    """Hidden-surface culling with a spherical z-buffer seen from `cam`."""
    d = points - cam
    rng_ = np.linalg.norm(d, axis=1)
    az = np.arctan2(d[:, 0], d[:, 2])
    el = np.arcsin(np.clip(d[:, 1] / rng_, -1, 1))
    iu = np.floor(az / ang_res).astype(np.int64)
    iv = np.floor(el / ang_res).astype(np.int64)
    key = (iu - iu.min()) * (iv.max() - iv.min() + 1) + (iv - iv.min())
    order = np.lexsort((rng_, key))
    ks = key[order]
    first = np.ones(len(ks), bool)
    first[1:] = ks[1:] != ks[:-1]
    grp = np.cumsum(first) - 1
    minr = rng_[order][first][grp]
    vis = np.zeros(len(points), bool)
    vis[order] = rng_[order] <= minr + 0.004
    return vis


def synthetic_raw_scene(seed, n_points=300000, two_cameras=False, step=0.002, nan_fraction=0.0):
    """RAW (unprocessed) cloud of the same cluttered-table scene family: the hidden-surface-culled noisy surface
    samples BEFORE voxelisation and normal estimation (several points per 3 mm voxel at the default 2 mm
    lattice), i.e. what CandidatesGenerator::preprocessPointCloud receives (candidates_generator.cpp:14-37).
    `n_points` sizes the scene like synthetic_table_scene (the voxelised cloud has roughly that many points);
    `nan_fraction` of the points get a NaN coordinate (depth-camera dropouts, cloud.cpp:154-164)."""
    d = synthetic_table_scene(seed, n_points=n_points, two_cameras=two_cameras, step=step, _raw=True)
    if nan_fraction > 0:
        rng = np.random.default_rng(seed + 1000)
        bad = rng.random(len(d["xyz"])) < nan_fraction
        d["xyz"][bad, rng.integers(0, 3, int(bad.sum()))] = np.nan
    return d


def synthetic_table_scene(seed, n_points=300000, two_cameras=False, step=0.003, _raw=False):
    """Config 3/4/5 cloud: a table plane at z ~ 0.9 m in front of a camera at the origin looking
    along +z, 30-60 boxes / cylinders / spheres (5-25 cm) resting on it, surfaces on a 3 mm
    lattice with sigma = 0.5 mm noise, hidden-surface culled per camera, voxelised at 0.003 and
    cut to exactly n_points. Normals are the analytic surface normals perturbed by ~3 degrees of
    noise (stand-in for PCA r=0.03), flipped towards the seeing camera, stored as float32 values.
    """
    rng = np.random.default_rng(seed)
    scale = (n_points / 300000.0) ** 0.5
    tx, ty, tz = 2.0 * scale, 1.5 * scale, 0.9
    cams = [np.zeros(3)] + ([np.array([0.6, 0.0, 0.0])] if two_cameras else [])
    table = _lattice_rect(np.array([-tx / 2, -ty / 2, tz]), np.array([1.0, 0, 0]), np.array([0, 1.0, 0]), tx, ty, step)
    P, Nn = [table], [np.tile([0.0, 0.0, -1.0], (len(table), 1))]
    n_obj = int(rng.integers(30, 61) * scale * scale) + 1
    for _ in range(n_obj):
        cx, cy = rng.uniform(-tx / 2 + 0.1, tx / 2 - 0.1), rng.uniform(-ty / 2 + 0.1, ty / 2 - 0.1)
        kind = rng.integers(0, 3)
        p, n = (_box, _cylinder, _sphere)[kind](rng, cx, cy, tz, step)
        P.append(p)
        Nn.append(n)
    pts = np.vstack(P)
    nrm = np.vstack(Nn)
    pts = pts + rng.normal(0, 0.0005, pts.shape)
    seen = np.zeros((len(pts), len(cams)), bool)
    for k, cam in enumerate(cams):
        facing = ((cam - pts) * nrm).sum(1) > 0
        seen[:, k] = facing & _visible(pts, cam, 0.6 * step / tz)
    keep = seen.any(1)
    pts, nrm, seen = pts[keep], nrm[keep], seen[keep]
    if _raw:
        order = rng.permutation(len(pts))  # scan order is not voxel order
        pts, seen = pts[order], seen[order]
        firstcam = np.argmax(seen, axis=1)
        cam_source = np.zeros((len(pts), len(cams)), np.int32)
        cam_source[np.arange(len(pts)), firstcam] = 1
        return {"xyz": np.ascontiguousarray(pts.astype(np.float32)), "cam_source": cam_source,
                "view_points": np.array(cams, dtype=np.float64)}
    xyz, first = voxelize(pts.astype(np.float32), step)
    nrm, seen = nrm[first], seen[first]
    if len(xyz) < n_points:
        raise RuntimeError(f"synthetic scene produced only {len(xyz)} < {n_points} points")
    sel = np.sort(rng.choice(len(xyz), n_points, replace=False))
    xyz, nrm, seen = xyz[sel], nrm[sel], seen[sel]
    nrm = nrm + rng.normal(0, 0.05, nrm.shape)
    nrm /= np.linalg.norm(nrm, axis=1, keepdims=True)
    firstcam = np.argmax(seen, axis=1)
    cam_source = np.zeros((n_points, len(cams)), np.int32)
    cam_source[np.arange(n_points), firstcam] = 1
    vp = np.array(cams, dtype=np.float64)
    to_cam = vp[firstcam] - xyz.astype(np.float64)
    flip = (to_cam * nrm).sum(1) < 0
    nrm[flip] *= -1
    normals = nrm.astype(np.float32).astype(np.float64)
    return {"xyz": np.ascontiguousarray(xyz), "normals": np.ascontiguousarray(normals),
            "cam_source": np.ascontiguousarray(cam_source), "view_points": vp}


def sample_indices(config, n_cloud, n_samples=None):
    """Seeded sample indices of BASELINE.json's configs."""
    if config == 1:
        return np.random.default_rng(1).choice(n_cloud, n_samples or 500, replace=False).astype(np.int32)
    if config == 2:
        return np.random.default_rng(2).integers(0, n_cloud, n_samples or 10000).astype(np.int32)
    if config == 3:
        return np.random.default_rng(3).choice(n_cloud, n_samples or 100000, replace=False).astype(np.int32)
    if config == 4:
        return np.random.default_rng(4).integers(0, n_cloud, n_samples or 1000000).astype(np.int32)
    if config == 5:
        return np.random.default_rng(5).integers(0, n_cloud, n_samples or 200000).astype(np.int32)
    raise ValueError(config)


def random_lenet_weights(channels, seed=0):
    """Random-init LeNet of the reference's architecture (A14) in the .bin layout, with weight
    scales close to the shipped models' (|w|max ~ 0.17 / 0.10 / 0.045 / 0.24)."""
    rng = np.random.default_rng(seed)
    f = np.float32
    return [
        (rng.standard_normal(20 * channels * 25) * 0.04).astype(f), (rng.standard_normal(20) * 0.1).astype(f),
        (rng.standard_normal(50 * 500) * 0.025).astype(f), (rng.standard_normal(50) * 0.1).astype(f),
        (rng.standard_normal(500 * 7200) * 0.008).astype(f), (rng.standard_normal(500) * 0.1).astype(f),
        (rng.standard_normal(2 * 500) * 0.05).astype(f), (rng.standard_normal(2) * 0.1).astype(f),
    ]


def load_weights_dir(d):
    """Read the reference's raw float32 .bin weight directory (eigen_classifier.cpp:185-205)."""
    names = ["conv1_weights", "conv1_biases", "conv2_weights", "conv2_biases", "ip1_weights", "ip1_biases",
             "ip2_weights", "ip2_biases"]
    return [np.fromfile(os.path.join(d, n + ".bin"), dtype=np.float32) for n in names]
