"""CPU tests: the oracle (oracle/gpd_oracle.cpp) against every pin available for this path.

The reference ships no asserting tests or golden vectors; the pins are
 * the known-answer conv example of src/tests/test_conv_layer.cpp:11-16,
 * real OpenCV (cv2.dilate / normalize / convertTo) outputs, committed in tests/golden/cv_pins.npz,
 * cv2.dnn forward passes of the reference's own .prototxt/.caffemodel (3 and 15 channels) and a
   float64 torch restatement of the OpenVINO IR 12-channel net (tests/golden/lenet_*.npz),
 * numpy.linalg for the eigen-solver, brute force for the radius search,
 * a regression fixture of the oracle itself on the tutorial cloud.
"""
import os
import zlib

import numpy as np
import pytest

from conftest import load_weights
from gpd_b200 import abi, scenes
from oracle import oracle


def test_conv_layer_known_answer():
    # src/tests/test_conv_layer.cpp:11-16 — 5x5 input, 3x3 kernel, zero bias
    X = np.array([1, 1, 1, 0, 0, 0, 1, 1, 1, 0, 0, 0, 1, 1, 1, 0, 0, 1, 1, 0, 0, 1, 1, 0, 0], np.float32).reshape(1, 5, 5)
    W = np.array([1, 0, 1, 0, 1, 0, 1, 0, 1], np.float32).reshape(1, 1, 3, 3)
    Y = oracle.conv_forward(X, W, np.zeros(1, np.float32), 3).reshape(3, 3)
    assert np.array_equal(Y, np.array([[4, 3, 4], [2, 4, 3], [2, 3, 4]], np.float32))


def test_opencv_pins(golden_dir):
    g = np.load(os.path.join(golden_dir, "cv_pins.npz"))
    bad = total = 0
    for img, ref, ch in zip(g["inputs"], g["outputs"], g["channels"]):
        out = oracle.dilate_normalize_u8(img[:, :, :ch])
        d = np.abs(out.astype(int) - ref[:, :, :ch].astype(int))
        assert d.max() <= 1
        bad += np.count_nonzero(d)
        total += d.size
    assert bad <= 2, f"{bad} of {total} pixels differ from cv2"


@pytest.mark.parametrize("ch,name", [(15, "lenet_caffe_15ch"), (3, "lenet_caffe_3ch"), (12, "lenet_ir_12ch")])
def test_lenet_against_reference_models(golden_dir, ch, name):
    g = np.load(os.path.join(golden_dir, name + ".npz"))
    w, relu = load_weights(ch)
    p = abi.default_params(ch, relu_after_conv=relu)
    scores, logits = oracle.classify(p, oracle.WeightPack(w), g["images"])
    ref = g["logits"]
    assert np.abs(logits - ref).max() <= 1e-5 * np.abs(ref).max()
    assert np.allclose(scores, ref[:, 1] - ref[:, 0], rtol=0, atol=2e-5 * np.abs(ref).max())


def test_eigen3_is_a_symmetric_eigensolver():
    rng = np.random.default_rng(0)
    for _ in range(500):
        N = rng.standard_normal((3, rng.integers(3, 60)))
        M = N @ N.T
        ev, evec = oracle.eigen3(M)
        w, v = np.linalg.eigh(M)
        assert np.all(np.diff(ev) >= 0)
        assert np.abs(ev - w).max() <= 1e-12 * max(w.max(), 1e-300)
        assert np.abs(evec.T @ evec - np.eye(3)).max() < 1e-12
        assert np.abs(M @ evec - evec * ev).max() <= 1e-11 * w.max()
    # degenerate inputs must not produce NaN
    for M in (np.zeros((3, 3)), np.eye(3), np.diag([1.0, 1.0, 0.0])):
        ev, evec = oracle.eigen3(M)
        assert np.isfinite(ev).all() and np.isfinite(evec).all()


def test_radius_search_flann_semantics():
    k = scenes.krylon_cloud()
    oc = oracle.OracleCloud(k["xyz"], k["normals"], k["cam_source"], k["view_points"])
    xyz = k["xyz"]
    for qi, r in [(10, 0.11), (500, 0.01), (2000, 0.10), (7, 0.0)]:
        q = xyz[qi]
        idx, d = oc.radius_search(q, r)
        dx = (q - xyz).astype(np.float32)
        dd = (dx[:, 0] * dx[:, 0] + dx[:, 1] * dx[:, 1]) + dx[:, 2] * dx[:, 2]
        bf = np.where(dd < np.float32(r * r))[0]
        assert set(idx.tolist()) == set(bf.tolist())
        order = np.lexsort((bf, dd[bf]))  # (dist, index)
        assert np.array_equal(idx, bf[order])
        assert np.array_equal(d, dd[bf][order])


def test_krylon_regression_fixture(golden_dir, weights15):
    g = np.load(os.path.join(golden_dir, "krylon_oracle_15ch.npz"))
    k = scenes.krylon_cloud()
    oc = oracle.OracleCloud(k["xyz"], k["normals"], k["cam_source"], k["view_points"])
    p = abi.default_params(15, keep_images=1)
    r = oc.detect(p, oracle.WeightPack(weights15), g["sample_idx"])
    assert np.array_equal(r["pose_flags"], g["pose_flags"])
    assert np.array_equal(r["frame_valid"], g["frame_valid"])
    assert np.allclose(r["frames"], g["frames"], atol=1e-12, rtol=0)
    for f in ("position", "frame", "top", "bottom", "center", "width"):
        assert np.allclose(r["candidates"][f], g["candidates"][f], atol=1e-12, rtol=0), f
    assert np.array_equal(r["candidates"]["finger_idx"], g["candidates"]["finger_idx"])
    crc = np.array([zlib.crc32(im.tobytes()) for im in r["images"]], np.uint32)
    assert np.array_equal(crc, g["image_crc32"])
    assert np.allclose(r["candidates"]["score"], g["candidates"]["score"], rtol=1e-5, atol=1e-2)


def test_structural_invariants_of_the_path():
    """Properties the reference semantics imply, checked on the oracle (and on the GPU in test_gpu_parity)."""
    s = scenes.synthetic_table_scene(11, n_points=30000)
    oc = oracle.OracleCloud(s["xyz"], s["normals"], s["cam_source"], s["view_points"])
    p = abi.default_params(3)
    sidx = scenes.sample_indices(3, 30000, 120)
    fr, valid = oc.frames(p, sidx)
    assert valid.all()
    F = fr.reshape(-1, 3, 3)  # rows: normal, binormal, curvature
    assert np.allclose(np.einsum("nij,nkj->nik", F, F), np.eye(3), atol=1e-9)
    poses, flags = oc.hand_search(p, sidx, fr, valid)
    v = (flags & 1) == 1
    assert v.any()
    pv = poses[v]
    R = pv["frame"].reshape(-1, 3, 3)
    assert np.allclose(np.einsum("nij,nkj->nik", R, R), np.eye(3), atol=1e-9)
    assert (pv["top"] - pv["bottom"] - 0.06 < 1e-12).all() and (pv["top"] >= 0.01 - 1e-15).all() and (pv["top"] <= 0.06 + 1e-12).all()
    assert (pv["width"] >= 0).all() and (pv["width"] <= 0.10 + 1e-9).all()
    assert ((flags & 8) <= ((flags & 4) << 1)).all()  # full antipodal implies half
    assert ((flags & 2) <= ((flags & 1) << 1)).all()  # filtered implies valid
    # sample order does not matter (each sample is independent)
    perm = np.random.default_rng(0).permutation(len(sidx))
    poses2, flags2 = oc.hand_search(p, sidx[perm], fr[perm], valid[perm])
    assert np.array_equal(flags2, flags[perm])
    assert np.array_equal(poses2["position"], poses["position"][perm])


def test_edge_cases():
    k = scenes.krylon_cloud()
    oc = oracle.OracleCloud(k["xyz"], k["normals"], k["cam_source"], k["view_points"])
    p = abi.default_params(3)
    r = oc.detect(p, None, np.zeros(0, np.int32))
    assert r["n_candidates"] == 0 and r["pose_flags"].shape == (0, 8)
    # a single isolated far-away point: the frame exists (the point itself) and the hand closes on that one
    # point (width 0), but the grasp lies outside workspace_grasps = [-1, 1]^3 -> valid, not filtered
    xyz = np.vstack([k["xyz"], [[5.0, 5.0, 5.0]]]).astype(np.float32)
    nrm = np.vstack([k["normals"], [[0.0, 0.0, 1.0]]])
    oc2 = oracle.OracleCloud(xyz, nrm, None, np.zeros((1, 3)))
    r = oc2.detect(p, None, np.array([len(xyz) - 1, 3], np.int32))
    assert r["frame_valid"].tolist() == [1, 1]
    assert ((r["pose_flags"][0] & 1) == 1).all() and ((r["pose_flags"][0] & 2) == 0).all()
    # duplicated sample indices give duplicated results
    r = oc.detect(p, None, np.array([100, 100, 7], np.int32))
    assert np.array_equal(r["pose_flags"][0], r["pose_flags"][1])


def test_shadow_variant_is_deterministic_and_documented():
    t = oracle.qtab()
    assert len(t) == 1024 and np.all(np.diff(t) > 0) and abs(t[511] + t[512]) < 1e-12 and 3.0 < t[-1] < 3.3
    k = scenes.krylon_cloud()
    oc = oracle.OracleCloud(k["xyz"], k["normals"], k["cam_source"], k["view_points"])
    p = abi.default_params(15, keep_images=1)
    sidx = np.array([1987, 42], np.int32)
    a = oc.detect(p, None, sidx, nthreads=1)
    b = oc.detect(p, None, sidx[::-1].copy(), nthreads=4)
    ia = {(c["sample_index"], c["pose_slot"]): im for c, im in zip(a["candidates"], a["images"])}
    ib = {(c["sample_index"], c["pose_slot"]): im for c, im in zip(b["candidates"], b["images"])}
    assert ia.keys() == ib.keys() and len(ia) > 0
    for key in ia:
        assert np.array_equal(ia[key], ib[key])
    # shadow channels are populated
    assert any(im.reshape(-1, 15)[:, 4].max() > 0 for im in a["images"])


def test_config1_krylon_500_samples_3ch_cpu_plumbing():
    """BASELINE config 1: tutorials/krylon.pcd (voxelised fixture), num_samples = 500 without replacement, 3-channel
    images, the reference's 3-channel LeNet weights — the CPU-only plumbing case, run through the oracle with the
    reference's three stage timers (grasp_detector.cpp:313-320)."""
    k = scenes.krylon_cloud()
    assert len(k["xyz"]) == 2373
    sidx = scenes.sample_indices(1, len(k["xyz"]))
    assert len(sidx) == 500 and len(np.unique(sidx)) == 500
    w, relu = load_weights(3)
    p = abi.default_params(3, relu_after_conv=relu)
    oc = oracle.OracleCloud(k["xyz"], k["normals"], k["cam_source"], k["view_points"])
    r = oc.detect(p, oracle.WeightPack(w), sidx)
    assert r["frame_valid"].all()
    assert r["n_candidates"] == ((r["pose_flags"] & 3) == 3).sum() > 1000
    assert np.isfinite(r["candidates"]["score"]).all()
    assert np.array_equal(np.isnan(r["pose_scores"]), (r["pose_flags"] & 3) != 3)
    assert (r["stage_seconds"][:3] > 0).all() and abs(r["stage_seconds"][:3].sum() - r["stage_seconds"][3]) < 0.05
    # top-5 selection as GraspDetector::selectGrasps would do it (grasp_detector.cpp:405-420)
    top = np.sort(r["candidates"]["score"])[::-1][:5]
    assert (np.diff(top) <= 0).all()
    r2 = oc.detect(p, oracle.WeightPack(w), sidx, nthreads=1)
    assert np.array_equal(r2["pose_flags"], r["pose_flags"]) and np.array_equal(r2["candidates"]["score"], r["candidates"]["score"])


def test_rotation_set_restatements_against_scipy_and_numpy():
    """Eigen::AngleAxisd::toRotationMatrix and VectorXd::LinSpaced restated in the oracle (hand_set.cpp:52-53,68-69,
    hand_search.cpp:151-155) against scipy's Rotation and numpy.linspace; the derived radii against the reference's
    formulas (hand_search.cpp:13-17, image_generator.cpp:43-46). The half-turn about y carries the +-1.22e-16
    off-diagonals of sin(pi) that Eigen produces."""
    import ctypes as C
    from scipy.spatial.transform import Rotation
    L = oracle.lib()
    rng = np.random.default_rng(3)
    for _ in range(100):
        axis = rng.standard_normal(3)
        axis /= np.linalg.norm(axis)
        ang = rng.uniform(-np.pi, np.pi)
        R = np.zeros(9)
        L.gpdo_angle_axis(C.c_double(ang), axis.ctypes.data_as(C.c_void_p), R.ctypes.data_as(C.c_void_p))
        assert np.allclose(R.reshape(3, 3).T, Rotation.from_rotvec(ang * axis).as_matrix(), atol=1e-15)  # column-major
    p = abi.default_params(15)
    out = np.zeros(4 + p.num_orientations + 9)
    L.gpdo_derived(C.byref(p), out.ctypes.data_as(C.c_void_p))
    assert out[0] == 8 and out[1] == 0.11 and out[2] == 0.10 and out[3] == 0.10
    angles = out[4:12]
    assert np.allclose(angles, np.linspace(-np.pi / 2, np.pi / 2, 9)[:8], atol=2e-16) and angles[0] == -np.pi / 2
    rb = out[12:21].reshape(3, 3).T
    assert np.allclose(rb, np.diag([-1.0, 1.0, -1.0]), atol=2e-16) and abs(rb[0, 2]) == abs(rb[2, 0]) == np.sin(np.pi)
    p2 = abi.default_params(15, num_orientations=5, hand_outer_diameter=0.2, finger_width=0.02, hand_depth=0.3, volume_width=0.05,
                            volume_depth=0.07)
    out2 = np.zeros(4 + 5 + 9)
    L.gpdo_derived(C.byref(p2), out2.ctypes.data_as(C.c_void_p))
    assert out2[0] == 5 and out2[1] == 0.3 and out2[2] == 0.07
    assert np.allclose(out2[4:9], np.linspace(-np.pi / 2, np.pi / 2, 6)[:5], atol=2e-16)


def test_sample_positions_extend_the_sample_indices():
    """Cloud::setSamples (cloud.cpp:662, used by the SIS outer loop): sample indices >= N address arbitrary float64
    positions. Positions that coincide with cloud points reproduce the index results (everything but the sample index
    and, for 15 channels, the index-seeded shadow draws); off-cloud positions keep their float64 value in the pose."""
    k = scenes.krylon_cloud()
    oc = oracle.OracleCloud(k["xyz"], k["normals"], k["cam_source"], k["view_points"])
    z = np.load(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "gpd_b200", "weights", "lenet_3ch.npz"))
    w = oracle.WeightPack([z[n] for n in oracle.WeightPack.NAMES])
    p = abi.default_params(3)
    sidx = scenes.sample_indices(1, len(k["xyz"]), 40)
    a = oc.detect(p, w, sidx)
    pidx = oc.set_samples(k["xyz"][sidx].astype(np.float64))
    assert pidx[0] == len(k["xyz"]) and len(pidx) == 40
    b = oc.detect(p, w, pidx)
    assert np.array_equal(a["frames"], b["frames"]) and np.array_equal(a["pose_flags"], b["pose_flags"])
    assert np.array_equal(a["pose_scores"], b["pose_scores"], equal_nan=True) and a["n_candidates"] == b["n_candidates"] > 0
    for f in a["candidates"].dtype.names:
        if f not in ("sample_index", "pad_"):
            assert np.array_equal(a["candidates"][f], b["candidates"][f]), f
    assert np.array_equal(b["candidates"]["sample_index"], pidx[b["candidates"]["sample_slot"]])
    # positions off the cloud: the pose keeps the float64 position, the search runs at its float32 image
    off = k["xyz"][sidx].astype(np.float64) + np.random.default_rng(0).normal(0, 1e-3, (40, 3))
    c = oc.detect(p, w, oc.set_samples(off))
    assert c["frame_valid"].all() and c["n_candidates"] > 0
    assert np.array_equal(c["candidates"]["sample"], off[c["candidates"]["sample_slot"]])
    # mixed indices in one call
    mix = np.concatenate([sidx[:5], oc.set_samples(off)[:5]]).astype(np.int32)
    d = oc.detect(p, w, mix)
    assert np.array_equal(d["pose_flags"][:5], a["pose_flags"][:5]) and np.array_equal(d["pose_flags"][5:], c["pose_flags"][:5])


def test_reevaluate_hypotheses_is_consistent_with_the_hand_search():
    """HandSearch::reevaluateHypotheses (hand_search.cpp:66-134) restated: re-labelling the hands of a hand search against the
    SAME cloud goes through another code path (evaluateFingers at the hand's own depth and finger index instead of the sweep
    + deepenHand) and must return the labels the search gave them; against a thinned cloud (the "other" cloud of
    GraspDetector::evalGroundTruth) labels may only change where points are missing, and an empty neighbourhood gives 0."""
    from conftest import load_weights
    k = scenes.krylon_cloud()
    p = abi.default_params(15)
    oc = oracle.OracleCloud(k["xyz"], k["normals"], k["cam_source"], k["view_points"])
    w, _ = load_weights(15)
    c = oc.detect(p, oracle.WeightPack(w), scenes.sample_indices(2, len(k["xyz"]), 150))["candidates"]
    labels, again = oc.reevaluate(p, c)
    assert len(c) > 500 and np.array_equal(labels, c["full_antipodal"].astype(np.int32))
    assert np.array_equal(again["half_antipodal"], c["half_antipodal"]) and np.array_equal(again["full_antipodal"], c["full_antipodal"])
    keep = np.arange(len(k["xyz"])) % 3 != 0
    thin = oracle.OracleCloud(k["xyz"][keep], k["normals"][keep], k["cam_source"][keep], k["view_points"])
    l2, h2 = thin.reevaluate(p, c)
    assert 0 < l2.sum() < labels.sum() + 200 and not np.array_equal(l2, labels)
    assert np.all(h2["full_antipodal"] <= h2["half_antipodal"])
    far = c[:4].copy()
    far["sample"] += 10.0
    l3, h3 = oc.reevaluate(p, far)
    assert not l3.any() and not h3["half_antipodal"].any()
