"""N2, ORB-SLAM2 half: a map dump written in the reference's own on-disk format (cv::FileStorage YAML:
Map.yml, KeyFrames/*.yml, FrameId.yml; KITTI .bin scans; pose list) is read back and flattened into a
KeyFramePack that equals a brute-force construction following BAError's own loops."""
import importlib
import os

import numpy as np

from conftest import PKG


def _world(rng, F=5, M=400, N=260):
    """A small consistent SLAM world: M map points, F keyframes observing random subsets at keypoint slots."""
    pts = {int(i * 3 + 7): rng.normal(0, 5, 3).astype(np.float32) for i in range(M)}
    ids = sorted(pts)
    kfs = []
    for f in range(F):
        th = 0.05 * f
        Tcw = np.eye(4, dtype=np.float32)
        Tcw[:3, :3] = np.array([[np.cos(th), 0, np.sin(th)], [0, 1, 0], [-np.sin(th), 0, np.cos(th)]], np.float32)
        Tcw[:3, 3] = rng.normal(0, 1, 3).astype(np.float32)
        keys = np.stack([rng.uniform(0, 1241, N), rng.uniform(0, 376, N)], 1).astype(np.float32)
        seen = rng.choice(ids, size=N // 2, replace=False)
        slots = rng.choice(N, size=N // 2, replace=False)
        kfs.append(dict(mnId=2 * f + 1, mnFrameId=3 * f + 2, Tcw=Tcw, keys=keys, obs={int(m): int(k) for m, k in zip(seen, slots)}))
    for a in kfs:      # covisibility: shared map points, best first
        w = sorted(((len(set(a["obs"]) & set(b["obs"])), b["mnId"]) for b in kfs if b is not a), reverse=True)
        a["cov"] = [(i, c) for c, i in w if c > 0]
    return pts, kfs


def _dump(tmp, pts, kfs, io):
    os.makedirs(tmp / "KeyFrames")
    for k in kfs:
        kp7 = np.zeros((len(k["keys"]), 7))
        kp7[:, :2] = k["keys"]; kp7[:, 2] = 31.0; kp7[:, 3] = -1.0; kp7[:, 5] = 2; kp7[:, 6] = -1
        io.write_cv_yaml(str(tmp / "KeyFrames" / f"{k['mnId']:06d}.yml"), {
            "mnId": k["mnId"], "mnFrameId": k["mnFrameId"], "mTimeStamp": 0.1 * k["mnId"], "mnGridCols": 64, "mnGridRows": 48,
            "fx": 718.856, "fy": 718.856, "cx": 607.1928, "cy": 185.2157, "N": len(k["keys"]),
            "mvKeysUn": kp7.reshape(-1), "mnMinX": 0, "mnMinY": 0, "mnMaxX": 1241, "mnMaxY": 376,
            "mvInvLevelSigma2": np.array([1.0, 0.694, 0.482], np.float32),
            "mK": np.array([[718.856, 0, 607.1928], [0, 718.856, 185.2157], [0, 0, 1]], np.float32), "Pose": k["Tcw"],
            "mvpMapPointsId": list(k["obs"].keys()), "mvpCorrKeyPointsId": list(k["obs"].values()),
            "mvpOrderedConnectedKeyFramesId": [i for i, _ in k["cov"]], "mvOrderedWeights": [c for _, c in k["cov"]],
            "mbFirstConnection": 0, "mpParentId": -1, "mspChildrensId": [], "mspLoopEdgesId": []})
    io.write_cv_yaml(str(tmp / "Map.yml"), {
        "mspMapPoints": {f"MapPoint_{i}": {"mnId": i, "mWorldPos": p.reshape(3, 1), "mnVisible": 3, "nObs": 2} for i, p in pts.items()},
        "mspKeyFrameId": [k["mnId"] for k in kfs]})
    io.write_cv_yaml(str(tmp / "FrameId.yml"), {"mnId": [k["mnId"] for k in kfs], "mnFrameId": [k["mnFrameId"] for k in kfs]})


def test_cv_yaml_roundtrip(tmp_path):
    io = importlib.import_module(PKG + ".orbslam_io")
    M = np.arange(12, dtype=np.float32).reshape(3, 4) / 7
    io.write_cv_yaml(str(tmp_path / "a.yml"), {"n": 3, "x": 1.5, "v": [1, 2, 3], "e": [], "M": M, "node": {"k": 2, "P": M[:2]}})
    txt = open(tmp_path / "a.yml").read()
    assert txt.startswith("%YAML:1.0\n---\n") and "!!opencv-matrix" in txt and "dt: f" in txt
    d = io.read_cv_yaml(str(tmp_path / "a.yml"))
    assert d["n"] == 3 and d["x"] == 1.5 and d["v"] == [1, 2, 3] and d["e"] == []
    assert d["M"].dtype == np.float32 and np.array_equal(d["M"], M) and np.array_equal(d["node"]["P"], M[:2])


def test_export_equals_brute_force_flattening(tmp_path, oracle_mod, pkg):
    io = importlib.import_module(PKG + ".orbslam_io")
    dataio = importlib.import_module(PKG + ".dataio")
    rng = np.random.default_rng(3)
    pts, kfs = _world(rng)
    _dump(tmp_path, pts, kfs, io)
    nframes = max(k["mnFrameId"] for k in kfs) + 1
    os.makedirs(tmp_path / "velodyne")
    files, poses = [], []
    for i in range(nframes):
        sc = np.concatenate([rng.normal(0, 10, (500 + i, 3)), rng.uniform(0, 1, (500 + i, 1))], 1).astype(np.float32)
        fn = str(tmp_path / "velodyne" / f"{i:06d}.bin")
        sc.tofile(fn); files.append(fn)
        T = np.eye(4); T[:3, 3] = [0.5 * i, 0.01 * i, 0]
        poses.append(T[:3].reshape(-1))
    np.savetxt(tmp_path / "lo.txt", np.asarray(poses))
    for nb, mw in ((2, 150), (0, 60)):
        pack = io.export_pack(str(tmp_path / "KeyFrames"), str(tmp_path / "Map.yml"), str(tmp_path / "FrameId.yml"), files,
                              str(tmp_path / "lo.txt"), num_best_covis=nb, min_covis_weight=mw)
        F = len(kfs)
        assert pack.n_kf == F and pack.n_keypoints == sum(len(k["keys"]) for k in kfs)
        raw = dataio.read_pose_list(str(tmp_path / "lo.txt"))
        ref0 = np.linalg.inv(raw[kfs[0]["mnFrameId"]])
        by_id = {k["mnId"]: k for k in kfs}
        for f, k in enumerate(kfs):                           # the reference's loops, object by object
            o = int(pack.kp_offset[f])
            assert np.array_equal(pack.kp_xy[o:o + len(k["keys"])], k["keys"])
            assert np.array_equal(pack.scan_xyz[pack.scan_offset[f]:pack.scan_offset[f + 1]],
                                  np.fromfile(files[k["mnFrameId"]], np.float32).reshape(-1, 4)[:, :3])
            kp2mp = {kp: m for m, kp in k["obs"].items()}
            for kp in range(len(k["keys"])):
                if kp in kp2mp:
                    assert np.array_equal(pack.kp_mappoint[o + kp], pts[kp2mp[kp]])
                else:
                    assert np.isnan(pack.kp_mappoint[o + kp, 0])
            if nb > 0:
                cov = [i for i, _ in k["cov"]][:nb]
            else:
                w = [c for _, c in k["cov"]]
                n = sum(1 for c in w if c >= mw)
                cov = [] if n == len(w) else [i for i, _ in k["cov"]][:n]
            assert int(pack.covis_valid[f].sum()) == len(cov)
            Twc = np.linalg.inv(k["Tcw"].astype(np.float64))
            for s, cid in enumerate(cov):
                other = by_id[cid]
                want = other["Tcw"].astype(np.float64) @ Twc
                assert np.allclose(pack.covis_relpose[f, s].reshape(3, 4), want[:3], atol=2e-6)
                for m, kp in k["obs"].items():
                    if m in other["obs"]:
                        assert np.array_equal(pack.covis_uv[o + kp, s], other["keys"][other["obs"][m]])
                n_match = len(set(k["obs"]) & set(other["obs"]))
                assert int((~np.isnan(pack.covis_uv[o:o + len(k["keys"]), s, 0])).sum()) == n_match
            if f + 1 < F:
                assert pack.he_valid[f] == 1
                assert np.allclose(pack.he_Tc[f].reshape(3, 4), (kfs[f + 1]["Tcw"].astype(np.float64) @ Twc)[:3], atol=2e-6)
                Twl_f, Twl_n = ref0 @ raw[k["mnFrameId"]], ref0 @ raw[kfs[f + 1]["mnFrameId"]]
                assert np.allclose(pack.he_Tl[f].reshape(3, 4), (np.linalg.inv(Twl_n) @ Twl_f)[:3], atol=1e-12)
            else:
                assert pack.he_valid[f] == 0
        # the flattened pack is a valid input of the evaluation
        s, _, _ = oracle_mod.Oracle(pack, kind="port").ba_error_sums(np.array([[1.2, -1.2, 1.2, 0.0, -0.08, -0.27, 12.0]]), mode=0)
        assert np.isfinite(s[:, 3:]).all()
