"""ctypes loader for oracle/libgpd_oracle.so — the CPU restatement of the reference path.

TEST INFRASTRUCTURE ONLY: imported by tests/, __graft_entry__.smoke() and bench.py's
cpu_baseline / --impl reference legs. Nothing under gpd_b200/ imports this module.
"""
import ctypes as C
import os
import subprocess

import numpy as np

from gpd_b200 import abi

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None


def build(force=False):
    so = os.path.join(_HERE, "libgpd_oracle.so")
    src = os.path.join(_HERE, "gpd_oracle.cpp")
    if force or not os.path.exists(so) or os.path.getmtime(so) < os.path.getmtime(src):
        subprocess.check_call(["make", "-C", _HERE, "-s"], env={**os.environ, "CXX": "g++"})
    return so


def lib():
    global _LIB
    if _LIB is not None:
        return _LIB
    so = os.path.join(_HERE, "libgpd_oracle.so")
    try:
        if not os.path.exists(so):
            build()
        L = C.CDLL(so)
    except OSError:
        build(force=True)
        L = C.CDLL(so)
    vp = C.c_void_p
    L.gpdo_cloud_create.restype = vp
    L.gpdo_cloud_create.argtypes = [vp, vp, vp, C.c_int32, vp, C.c_int32]
    L.gpdo_cloud_destroy.argtypes = [vp]
    L.gpdo_cloud_set_samples.argtypes = [vp, vp, C.c_int32]
    L.gpdo_radius_search.argtypes = [vp, vp, C.c_double, vp, vp, C.c_int32]
    L.gpdo_eigen3.argtypes = [vp, vp, vp]
    L.gpdo_derived.argtypes = [C.POINTER(abi.Params), vp]
    L.gpdo_frames.argtypes = [vp, C.POINTER(abi.Params), vp, C.c_int32, vp, vp, C.c_int32]
    L.gpdo_hand_search.argtypes = [vp, C.POINTER(abi.Params), vp, C.c_int32, vp, vp, vp, vp, C.c_int32]
    L.gpdo_images.argtypes = [vp, C.POINTER(abi.Params), vp, C.c_int32, vp, C.c_int32]
    L.gpdo_classify.argtypes = [C.POINTER(abi.Params), vp, vp, C.c_int32, vp, vp, C.c_int32]
    L.gpdo_detect.argtypes = [vp, C.POINTER(abi.Params), vp, vp, C.c_int32, C.POINTER(abi.Result), C.c_int32, vp]
    L.gpdo_free_result.argtypes = [C.POINTER(abi.Result)]
    L.gpdo_reevaluate.argtypes = [vp, C.POINTER(abi.Params), vp, C.c_int32, vp, C.c_int32]
    L.gpdo_dilate_normalize_u8.argtypes = [vp, C.c_int32, C.c_int32, vp]
    L.gpdo_conv_forward.argtypes = [vp, C.c_int32, C.c_int32, C.c_int32, vp, vp, C.c_int32, C.c_int32, vp]
    L.gpdo_angle_axis.argtypes = [C.c_double, vp, vp]
    L.gpdo_qtab.argtypes = [vp]
    L.gpdo_num_threads.restype = C.c_int
    L.gpdo_preprocess.argtypes = [vp, vp, vp, C.c_int32, vp, C.c_int32, C.POINTER(abi.PreprocessParams), vp, vp, vp, vp, vp,
                                  C.c_int32]
    L.gpdo_pcl_eigen33.argtypes = [vp, vp, vp]
    L.gpdo_voxelize_literal.argtypes = [vp, C.c_int32, C.c_double, vp, vp]
    _LIB = L
    return L


def _p(a):
    return a.ctypes.data_as(C.c_void_p)


def num_threads():
    return int(lib().gpdo_num_threads())


class WeightPack:
    """The 8 LeNet arrays in the reference's .bin layout (eigen_classifier.cpp:24-47)."""

    NAMES = ["conv1_weights", "conv1_biases", "conv2_weights", "conv2_biases", "ip1_weights", "ip1_biases",
             "ip2_weights", "ip2_biases"]

    def __init__(self, arrays):
        self.arrays = [np.ascontiguousarray(a, dtype=np.float32).ravel() for a in arrays]
        self.ptrs = (C.c_void_p * 8)(*[a.ctypes.data for a in self.arrays])


class OracleCloud:
    def __init__(self, xyz, normals, cam_source=None, view_points=None):
        """xyz [N,3] f32; normals [N,3] f64 (stored 3xN column-major = same memory);
        cam_source [N,K] int32 (k x N column-major = same memory); view_points [K,3] f64."""
        self.xyz = np.ascontiguousarray(xyz, dtype=np.float32)
        self.normals = np.ascontiguousarray(normals, dtype=np.float64)
        self.view_points = np.ascontiguousarray(
            view_points if view_points is not None else np.zeros((1, 3)), dtype=np.float64)
        self.K = self.view_points.shape[0]
        self.N = self.xyz.shape[0]
        self.cam = None if cam_source is None else np.ascontiguousarray(cam_source, dtype=np.int32)
        self.h = lib().gpdo_cloud_create(_p(self.xyz), _p(self.normals), None if self.cam is None else _p(self.cam),
                                         self.N, _p(self.view_points), self.K)

    def __del__(self):
        if getattr(self, "h", None):
            lib().gpdo_cloud_destroy(self.h)
            self.h = None

    def set_samples(self, samples):
        """Cloud::setSamples: arbitrary float64 sample positions [n, 3]; sample indices N .. N+n-1 address them."""
        sm = np.ascontiguousarray(samples, dtype=np.float64)
        lib().gpdo_cloud_set_samples(self.h, _p(sm), len(sm))
        return np.arange(self.N, self.N + len(sm), dtype=np.int32)

    def radius_search(self, q, radius, cap=1 << 20):
        q = np.ascontiguousarray(q, dtype=np.float32)
        idx = np.zeros(cap, np.int32)
        dist = np.zeros(cap, np.float32)
        n = lib().gpdo_radius_search(self.h, _p(q), float(radius), _p(idx), _p(dist), cap)
        return idx[:n].copy(), dist[:n].copy()

    def frames(self, params, sample_idx, nthreads=0):
        sidx = np.ascontiguousarray(sample_idx, dtype=np.int32)
        n = len(sidx)
        frames = np.zeros((n, 9))
        valid = np.zeros(n, np.uint8)
        lib().gpdo_frames(self.h, C.byref(params), _p(sidx), n, _p(frames), _p(valid), nthreads or num_threads())
        return frames, valid

    def hand_search(self, params, sample_idx, frames, valid, nthreads=0):
        sidx = np.ascontiguousarray(sample_idx, dtype=np.int32)
        n = len(sidx)
        P = params.num_hand_axes * params.num_orientations
        poses = np.zeros(n * P, dtype=abi.POSE_DTYPE)
        flags = np.zeros(n * P, np.uint8)
        frames = np.ascontiguousarray(frames)
        valid = np.ascontiguousarray(valid, dtype=np.uint8)
        lib().gpdo_hand_search(self.h, C.byref(params), _p(sidx), n, _p(frames), _p(valid), _p(poses), _p(flags),
                               nthreads or num_threads())
        return poses.reshape(n, P), flags.reshape(n, P)

    def images(self, params, poses, nthreads=0):
        poses = np.ascontiguousarray(poses, dtype=abi.POSE_DTYPE)
        n = len(poses)
        S, Cc = params.image_size, params.image_num_channels
        out = np.zeros((n, S, S, Cc), np.uint8)
        lib().gpdo_images(self.h, C.byref(params), _p(poses), n, _p(out), nthreads or num_threads())
        return out

    def reevaluate(self, params, hands, nthreads=0):
        """HandSearch::reevaluateHypotheses: labels (1 = full antipodal against THIS cloud) + the re-labelled records."""
        hands = np.array(hands, dtype=abi.POSE_DTYPE, copy=True)
        labels = np.zeros(len(hands), np.int32)
        lib().gpdo_reevaluate(self.h, C.byref(params), _p(hands), len(hands), _p(labels), nthreads or num_threads())
        return labels, hands

    def images_literal_shadow(self, params, poses, lcg_seed, mt_seed):
        """Grasp images with the occlusion channels in the reference's LITERAL semantics (sequential shared LCG stream +
        mt19937 Gaussian jitter; hand_set.cpp:187-233,263-266) for the two given seeds — the statistical yardstick of the
        deterministic variant."""
        poses = np.ascontiguousarray(poses, dtype=abi.POSE_DTYPE)
        n = len(poses)
        S, Cc = params.image_size, params.image_num_channels
        out = np.zeros((n, S, S, Cc), np.uint8)
        lib().gpdo_images_literal_shadow(self.h, C.byref(params), _p(poses), n, _p(out), C.c_uint32(lcg_seed), C.c_uint32(mt_seed))
        return out

    def detect(self, params, weights, sample_idx, nthreads=0):
        sidx = np.ascontiguousarray(sample_idx, dtype=np.int32)
        res = abi.Result()
        st = np.zeros(4)
        nc = lib().gpdo_detect(self.h, C.byref(params), weights.ptrs if weights is not None else None, _p(sidx),
                               len(sidx), C.byref(res), nthreads or num_threads(), _p(st))
        S, Cc = params.image_size, params.image_num_channels
        out = abi.result_to_numpy(res, S * S * Cc)
        out["stage_seconds"] = st
        lib().gpdo_free_result(C.byref(res))
        assert nc == out["n_candidates"]
        return out


def classify(params, weights, images, nthreads=0):
    images = np.ascontiguousarray(images, dtype=np.uint8)
    n = images.shape[0]
    scores = np.zeros(n, np.float32)
    logits = np.zeros((n, 2), np.float32)
    lib().gpdo_classify(C.byref(params), weights.ptrs, _p(images), n, _p(scores), _p(logits),
                        nthreads or num_threads())
    return scores, logits


def eigen3(M):
    M = np.asfortranarray(M, dtype=np.float64)
    ev = np.zeros(3)
    evec = np.zeros((3, 3), order="F")
    lib().gpdo_eigen3(_p(M), _p(ev), _p(evec))
    return ev, evec


def dilate_normalize_u8(img):
    """img [S,S,ch] float32 -> cv::dilate(3x3) -> cv::normalize(MINMAX) -> convertTo(CV_8U,255)."""
    img = np.ascontiguousarray(img, dtype=np.float32)
    S, ch = img.shape[0], (img.shape[2] if img.ndim == 3 else 1)
    out = np.zeros((S, S, ch), np.uint8)
    lib().gpdo_dilate_normalize_u8(_p(img), S, ch, _p(out))
    return out


def qtab():
    t = np.zeros(1024)
    lib().gpdo_qtab(_p(t))
    return t


def conv_forward(x, w, b, k):
    """ConvLayer::forward: x [C,H,W], w [M,C,k,k], b [M] -> [M, oh*ow]."""
    x = np.ascontiguousarray(x, np.float32)
    w = np.ascontiguousarray(w, np.float32)
    b = np.ascontiguousarray(b, np.float32)
    Cc, H, Wd = x.shape
    M = w.shape[0]
    out = np.zeros((M, (H - k + 1) * (Wd - k + 1)), np.float32)
    lib().gpdo_conv_forward(_p(x), Cc, H, Wd, _p(w), _p(b), M, k, _p(out))
    return out


def preprocess(xyz, cam_source, view_points, pp, normals=None, nthreads=0):
    """CandidatesGenerator::preprocessPointCloud restated on the CPU (removeNans, filterWorkspace, voxelizeCloud,
    calculateNormalsOMP, reverseNormals). Returns a cloud dict + `src` (raw index of each output point) +
    `seconds` (voxelise, normals)."""
    xyz = np.ascontiguousarray(xyz, dtype=np.float32)
    M = len(xyz)
    vp_ = np.ascontiguousarray(view_points, dtype=np.float64)
    K = vp_.shape[0]
    cam = None if cam_source is None else np.ascontiguousarray(cam_source, dtype=np.int32)
    nrm = None if normals is None else np.ascontiguousarray(normals, dtype=np.float64)
    xo = np.zeros((M, 3), np.float32)
    no = np.zeros((M, 3), np.float64)
    co = np.zeros((M, K), np.int32)
    so = np.zeros(M, np.int32)
    sec = np.zeros(2)
    n = lib().gpdo_preprocess(_p(xyz), None if nrm is None else _p(nrm), None if cam is None else _p(cam), M, _p(vp_), K,
                              C.byref(pp), _p(xo), _p(no), _p(co), _p(so), _p(sec), nthreads or num_threads())
    if n < 0:
        raise RuntimeError(f"gpdo_preprocess failed: {n}")
    return {"xyz": xo[:n].copy(), "normals": no[:n].copy(), "cam_source": co[:n].copy(), "view_points": vp_,
            "src": so[:n].copy(), "seconds": sec}


def pcl_eigen33(cov):
    """pcl::eigen33 (smallest eigenvalue + eigenvector) of a 3x3 float32 symmetric matrix."""
    cov = np.ascontiguousarray(cov, dtype=np.float32)
    ev = np.zeros(1, np.float32)
    vec = np.zeros(3, np.float32)
    lib().gpdo_pcl_eigen33(_p(cov), _p(ev), _p(vec))
    return float(ev[0]), vec


def voxelize_literal(xyz, voxel_size=0.003):
    """Cloud::voxelizeCloud with the literal libstdc++ behaviour of its std::set (non-strict-weak comparator):
    returns (src indices in the reference's iteration order, number of duplicate voxels that slipped in)."""
    xyz = np.ascontiguousarray(xyz, dtype=np.float32)
    src = np.zeros(len(xyz), np.int32)
    missed = np.zeros(1, np.int32)
    n = lib().gpdo_voxelize_literal(_p(xyz), len(xyz), float(voxel_size), _p(src), _p(missed))
    return src[:n].copy(), int(missed[0])
