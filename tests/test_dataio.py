"""Readers of the reference's plain on-disk inputs (dataio.py): config YAML, KITTI .bin scans, pose
lists, Sim3 files, and the pack pieces derived from them."""
import importlib

import numpy as np
import pytest

from conftest import PKG

CONFIG = """%YAML:1.0
---
io:
  BaseDir: ../KITTI-00/
  PointCloudskip: 1
  PointCloudOnlyPositiveX: true
runtime:
  max_pixel_dist: 1.25
  num_best_covis: 3 # set negative to use min_covis_weight
  kdtree3d_max_leaf_size: 30
  corr_3d_2d_threshold: 40
  corr_3d_3d_threshold: 10
  norm_max_pts: 20
  norm_min_pts: 6
  norm_radius: 0.5
  norm_reg_threshold: 0.03
  min_diff_dist: 0.25
  he_threshold: 0.094
  err_weight: [1.0, 0.5]
  lb: [-0.1,-0.1,-0.1,-0.3,-0.3,-0.3,-1.0]
  ub: [0.1,0.1,0.1,0.3,0.3,0.3,1.0]
  max_bbeval: 5000
  valid_rate: 0.9
  use_plane: false
"""


def test_params_from_reference_style_config(pkg, tmp_path):
    dio = importlib.import_module(PKG + ".dataio")
    p, extras = dio.params_from_config(CONFIG)
    assert (p.max_pixel_dist, p.norm_max_pts, p.norm_min_pts, p.norm_radius) == (1.25, 20, 6, 0.5)
    assert (p.norm_reg_threshold, p.min_diff_dist, p.valid_rate, p.use_plane) == (0.03, 0.25, 0.9, 0)
    assert (p.err_weight[0], p.err_weight[1]) == (1.0, 0.5) and p.corr_3d_2d_threshold == 40.0
    assert extras["max_bbeval"] == 5000 and extras["lb"][6] == -1.0 and extras["io"]["PointCloudOnlyPositiveX"] is True
    f = tmp_path / "cfg.yml"
    f.write_text(CONFIG)
    p2, _ = dio.params_from_config(str(f))
    assert bytes(p2) == bytes(p)
    # an iba_local style block uses other key names for the same quantities (iba_local.cpp:363-371)
    p3, _ = dio.params_from_config("runtime:\n  neigh_radius: 0.7\n  neigh_max_pts: 25\n  robust_kernel_delta: 2.0\n  init_sigma: 5.0\n")
    assert (p3.norm_radius, p3.norm_max_pts, p3.robust_kernel_delta, p3.gpr_sigma) == (0.7, 25, 2.0, 5.0)


def test_kitti_bin_reader_follows_the_reference_loop(pkg, tmp_path):
    dio = importlib.import_module(PKG + ".dataio")
    rng = np.random.default_rng(0)
    pts = rng.normal(0, 10, (1001, 4)).astype(np.float32)
    f = tmp_path / "000000.bin"
    pts.tofile(f)
    a = dio.read_pointcloud_bin(str(f))
    assert a.dtype == np.float32 and np.array_equal(a, pts[:, :3])
    b = dio.read_pointcloud_bin(str(f), only_positive_x=True)
    assert np.array_equal(b, pts[pts[:, 0] > 0, :3])
    # skip = 4: the reference's loop runs for i = 0, 4, ..., <= n - 4 but reads consecutive records
    c = dio.read_pointcloud_bin(str(f), skip=4)
    assert len(c) == (1001 - 4) // 4 + 1 and np.array_equal(c, pts[: len(c), :3])


def test_pose_and_sim3_files(pkg, oracle_mod, tmp_path):
    dio = importlib.import_module(PKG + ".dataio")
    rng = np.random.default_rng(1)
    poses = []
    T = np.eye(4)
    for _ in range(8):
        x = np.r_[rng.normal(0, 0.05, 3), rng.normal(0, 1.0, 3), 1.0]
        R, t, _ = oracle_mod.sim3exp(x)
        S = np.eye(4); S[:3, :3] = R; S[:3, 3] = t
        T = T @ S
        poses.append(T.copy())
    f = tmp_path / "floam.txt"
    np.savetxt(f, np.array([p[:3].reshape(-1) for p in poses]), fmt="%.17g")
    P = dio.read_pose_list(str(f))
    assert P.shape == (8, 4, 4) and np.array_equal(P, np.array(poses))
    # keyframes 2, 4, 5, 7: re-based on the first keyframe's frame (iba_global.cpp:480-483)
    Twl = dio.lidar_poses_for_keyframes(P, [2, 4, 5, 7])
    assert np.allclose(Twl[0], np.eye(4), atol=1e-14) and np.allclose(Twl[2], np.linalg.inv(P[2]) @ P[5])
    assert np.array_equal(dio.lidar_poses_for_keyframes(P, [0, 3])[1], P[3])
    he_Tl, he_valid = dio.hand_eye_lidar_motions(Twl)
    assert he_valid.tolist() == [1, 1, 1, 0]
    assert np.allclose(he_Tl[1].reshape(3, 4), (np.linalg.inv(Twl[2]) @ Twl[1])[:3])
    # Sim3 file round trip and the 7-vector handed to the evaluator
    x = np.array([0.31, -1.2, 0.8, 0.1, -0.4, 0.25, 17.5])
    R, t, s = oracle_mod.sim3exp(x)
    # Sim3Exp's translation is V*upsilon and its scale a plain multiplier, so log(R, t) returns x[:6]
    S = np.eye(4); S[:3, :3] = R; S[:3, 3] = t
    g = tmp_path / "he_calib.txt"
    dio.write_sim3(str(g), S, s)
    S2, s2 = dio.read_sim3(str(g))
    assert np.array_equal(S2, S) and s2 == s
    x_back = dio.sim3_to_x(S2, s2)
    assert np.allclose(x_back, x, atol=1e-12)
    assert np.allclose(x_back[:6], oracle_mod.se3log(R, t), atol=1e-12)
    lb, ub = dio.search_box(x_back, [-0.1] * 3 + [-0.3] * 3 + [-1.0], [0.1] * 3 + [0.3] * 3 + [1.0])
    assert np.allclose(ub - lb, [0.2] * 3 + [0.6] * 3 + [2.0])
