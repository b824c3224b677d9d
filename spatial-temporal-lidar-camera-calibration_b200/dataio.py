"""Readers for the reference's on-disk inputs that do not need ORB-SLAM2, and the pieces of the
KeyFramePack they determine (first half).

The keyframe side of the pack (keypoints, map points, covisibility, ``Tcw``) lives in ORB-SLAM2's own
``Map.yml`` / ``KeyFrames/*.{yml,bin}`` dump and is exported from inside the reference
(``INTEGRATION.md`` §1); everything else the hot path reads comes from plain files:

* calibration config (OpenCV-flavoured YAML)          -> :func:`params_from_config`
  ``src/examples/iba_global.cpp:413-468``, ``iba_local.cpp:358-377``
* KITTI velodyne ``.bin`` scans                       -> :func:`read_pointcloud_bin`
  ``include/io_tools.h:142-196`` (incl. its ``skip`` behaviour, see below)
* pose lists (12 numbers per line)                    -> :func:`read_pose_list`
  ``include/kitti_tools.h:66-87``
* Sim3 files (12 numbers + scale)                     -> :func:`read_sim3`, :func:`write_sim3`
  ``include/kitti_tools.h:118-158``
* LiDAR poses of the keyframes and the hand-eye pairs -> :func:`lidar_poses_for_keyframes`,
  :func:`hand_eye_lidar_motions`   (``iba_global.cpp:473-484,264-270``)
* initial estimate and search box                     -> :func:`sim3_to_x`, :func:`search_box`
  (``iba_global.cpp:513-544``)
"""
from __future__ import annotations

import os
import re

import numpy as np

from . import _abi
from .pack import default_params


# ----------------------------------------------------------------------------- config
def _load_yaml(text_or_path: str) -> dict:
    import yaml
    text = text_or_path
    if "\n" not in text_or_path and os.path.exists(text_or_path):
        with open(text_or_path) as f:
            text = f.read()
    # OpenCV FileStorage header ("%YAML:1.0" + "---") is not valid YAML 1.1 for PyYAML
    text = re.sub(r"^%YAML[^\n]*\n", "", text.lstrip())
    return yaml.safe_load(text) or {}


def params_from_config(text_or_path: str):
    """``runtime:`` block of an ``iba_calib_global.yml`` / iba_local config -> (stl_params_t, extras).

    Keys the evaluator does not consume (bounds, NOMAD / Ceres settings, covisibility selection used
    while building the pack, KD-tree leaf sizes) are returned in ``extras`` untouched.  Missing keys keep
    the defaults of ``IBAGlobalParams`` / ``IBALocalParams`` as shipped in ``stl_default_params``."""
    cfg = _load_yaml(text_or_path)
    rt = dict(cfg.get("runtime", {}))
    p = default_params()
    direct = {  # config key -> params field (iba_global.cpp:440-461)
        "max_pixel_dist": "max_pixel_dist", "corr_3d_2d_threshold": "corr_3d_2d_threshold",
        "corr_3d_3d_threshold": "corr_3d_3d_threshold", "norm_max_pts": "norm_max_pts", "norm_min_pts": "norm_min_pts",
        "norm_radius": "norm_radius", "norm_reg_threshold": "norm_reg_threshold", "min_diff_dist": "min_diff_dist",
        "he_threshold": "he_threshold", "valid_rate": "valid_rate",
        # iba_local.cpp:363-371 uses other names for the same quantities
        "neigh_radius": "norm_radius", "neigh_max_pts": "norm_max_pts", "robust_kernel_delta": "robust_kernel_delta",
        "robust_kernel_3ddelta": "robust_kernel_3ddelta", "max_3d_dist": "max_3d_dist",
        "init_sigma": "gpr_sigma", "init_l": "gpr_l", "sigma_noise": "gpr_sigma_noise",
    }
    used = set()
    for key, field in direct.items():
        if key in rt:
            cur = getattr(p, field)
            setattr(p, field, type(cur)(rt[key]))
            used.add(key)
    if "err_weight" in rt:
        p.err_weight[0], p.err_weight[1] = float(rt["err_weight"][0]), float(rt["err_weight"][1])
        used.add("err_weight")
    if "use_plane" in rt:
        p.use_plane = int(bool(rt["use_plane"]))
        used.add("use_plane")
    extras = {k: v for k, v in rt.items() if k not in used}
    extras["io"] = cfg.get("io", {})
    extras["orb"] = cfg.get("orb", {})
    return p, extras


# ----------------------------------------------------------------------------- scans
def read_pointcloud_bin(path: str, skip: int = 1, only_positive_x: bool = False) -> np.ndarray:
    """KITTI ``.bin`` (x, y, z, intensity as float32) -> [M,3] float32, as ``readPointCloud`` does
    (io_tools.h:154-187).  Two behaviours are reproduced on purpose: with ``skip > 1`` the reference reads
    the FIRST ``(n - skip) // skip + 1`` points consecutively (it advances its counter by ``skip`` but never
    seeks), and ``only_positive_x`` drops points with ``x <= 0`` among those."""
    raw = np.fromfile(path, dtype=np.float32)
    n = raw.size // 4
    pts = raw[: n * 4].reshape(n, 4)
    if skip < 1:
        raise ValueError("skip must be >= 1")
    take = (n - skip) // skip + 1 if n >= skip else 0
    pts = pts[:take, :3]
    if only_positive_x:
        pts = pts[pts[:, 0] > 0]
    return np.ascontiguousarray(pts, dtype=np.float32)


# ----------------------------------------------------------------------------- poses
def read_pose_list(path: str) -> np.ndarray:
    """One pose per line, 12 numbers = the first three rows of a 4x4 (kitti_tools.h:66-87) -> [n,4,4] fp64."""
    vals = np.loadtxt(path, dtype=np.float64, ndmin=2)
    if vals.shape[1] != 12:
        raise ValueError(f"{path}: expected 12 numbers per line, found {vals.shape[1]}")
    T = np.tile(np.eye(4), (len(vals), 1, 1))
    T[:, :3, :] = vals.reshape(-1, 3, 4)
    return T


def read_sim3(path: str):
    """12 numbers (3x4 rigid part, row-major) + scale (kitti_tools.h:147-158) -> (4x4, scale)."""
    v = np.loadtxt(path, dtype=np.float64).reshape(-1)
    T = np.eye(4)
    T[:3, :] = v[:12].reshape(3, 4)
    return T, float(v[12]) if v.size > 12 else 1.0


def write_sim3(path: str, T: np.ndarray, scale: float) -> None:
    """The inverse of :func:`read_sim3` with max_digits10 precision (kitti_tools.h:118-140)."""
    with open(path, "w") as f:
        f.write(" ".join(repr(float(x)) for x in np.asarray(T)[:3, :].reshape(-1)) + " " + repr(float(scale)))


def lidar_poses_for_keyframes(raw_poses: np.ndarray, frame_ids) -> np.ndarray:
    """``PointCloudPoses`` (iba_global.cpp:473-484): the LiDAR odometry poses of the keyframes' frames,
    re-based on the first keyframe unless that is frame 0."""
    ids = np.asarray(frame_ids, dtype=np.int64)
    P = raw_poses[ids]
    if ids[0] != 0:
        P = np.linalg.inv(raw_poses[ids[0]])[None] @ P
    return P


def hand_eye_lidar_motions(Twl: np.ndarray):
    """``Tl = vTwl[i+1].inverse() * vTwl[i]`` (iba_global.cpp:269) for every keyframe that has a successor
    -> (he_Tl [F,12] fp64, he_valid [F] uint8) in KeyFramePack layout."""
    F = len(Twl)
    he_Tl = np.tile(np.eye(4)[:3].reshape(-1), (F, 1)).astype(np.float64)
    he_valid = np.zeros(F, np.uint8)
    for i in range(F - 1):
        he_Tl[i] = (np.linalg.inv(Twl[i + 1]) @ Twl[i])[:3].reshape(-1)
        he_valid[i] = 1
    return he_Tl, he_valid


# ----------------------------------------------------------------------------- estimate
def sim3_to_x(T: np.ndarray, scale: float) -> np.ndarray:
    """(rigid 4x4, scale) -> the 7 parameters [omega, upsilon, s] the evaluator takes:
    ``g2o::SE3Quat(R, t).log()`` and the scale (iba_global.cpp:519-522).  The logarithm is evaluated by
    the same routine the hand-eye term uses, through the oracle-independent closed form below."""
    R, t = np.asarray(T)[:3, :3], np.asarray(T)[:3, 3]
    c = np.clip((np.trace(R) - 1.0) * 0.5, -1.0, 1.0)
    th = np.arccos(c)
    w_hat = np.array([R[2, 1] - R[1, 2], R[0, 2] - R[2, 0], R[1, 0] - R[0, 1]]) * 0.5
    if th < 1e-7:
        omega = w_hat
        Vinv = np.eye(3) - 0.5 * _skew(omega)
    else:
        omega = w_hat * (th / np.sin(th))
        K = _skew(omega)
        Vinv = np.eye(3) - 0.5 * K + (1.0 - th * np.cos(th * 0.5) / (2.0 * np.sin(th * 0.5))) / (th * th) * (K @ K)
    return np.concatenate([omega, Vinv @ t, [float(scale)]])


def _skew(w):
    return np.array([[0, -w[2], w[1]], [w[2], 0, -w[0]], [-w[1], w[0], 0]], dtype=np.float64)


def search_box(x0: np.ndarray, lb, ub):
    """NOMAD's bounds: ``offset + lb``, ``offset + ub`` (iba_global.cpp:535-538)."""
    x0 = np.asarray(x0, dtype=np.float64)
    return x0 + np.asarray(lb, dtype=np.float64), x0 + np.asarray(ub, dtype=np.float64)
