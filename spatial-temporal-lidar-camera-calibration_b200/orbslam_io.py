"""KeyFramePack exporter, ORB-SLAM2 half: reads the map dump the reference's ORB-SLAM2 fork
writes — ``Map.yml``, ``KeyFrames/NNNNNN.yml`` and ``FrameId.yml`` — WITHOUT linking ORB-SLAM2, and flattens it
into the ``stl_pack_t`` the GPU path consumes, following what BAError does with the restored objects.

On-disk format (OpenCV ``cv::FileStorage`` YAML 1.0):

* ``KeyFrames/<name>.yml`` — ``KeyFrame::saveData`` (src/orb_slam/src/KeyFrame.cc:169-258), read back by
  ``KeyFrameConstInfo`` (KeyFrame.cc:31-80): ``mnId, mnFrameId, fx, fy, cx, cy, N, mvKeysUn`` (cv::KeyPoint = 7 numbers:
  x, y, size, angle, response, octave, class_id), ``mnMinX..mnMaxY, Pose`` (4x4 CV_32F ``Tcw``),
  ``mvpMapPointsId`` / ``mvpCorrKeyPointsId`` (map point id -> keypoint index, KeyFrame.cc:76-79,108-132),
  ``mvpOrderedConnectedKeyFramesId`` / ``mvOrderedWeights`` (covisibility, best first), ``mvInvLevelSigma2``.
  The ``.bin`` twin (boost archive: grid, BoW) is not needed by the cost evaluation.
* ``Map.yml`` — ``operator<<(FileStorage, Map)`` (Map.cc:214-230): ``mspMapPoints: {MapPoint_<id>: {mnId, mWorldPos (3x1
  CV_32F), ...}}`` (MapPoint.cc:435-456) and ``mspKeyFrameId``.
* ``FrameId.yml`` — ``mnId`` / ``mnFrameId`` of the keyframes (System.cc:612-624), which picks the LiDAR scans and
  odometry poses (iba_global.cpp:469-499).

What the flattening mirrors (src/examples/iba_global.cpp):
  keyframes sorted by ``mnId`` (:505), ``mmapMpt2Kpt`` inverted to keypoint -> map point (:209-212),
  covisible keyframes ``GetBestCovisibilityKeyFramesSafe(num_best_covis)`` or ``GetCovisiblesByWeightSafe(min_covis_weight)``
  (:255-259; KeyFrame.cc:417-439 incl. its "all weights pass => empty" behaviour), keypoint-to-keypoint tables
  ``GetMatchedKptIds`` (KeyFrame.cc:528-538), ``relPose = Tcw_j * Twc_ref`` as a float32 product (:280), the hand-eye
  camera motion ``Tcw_{i+1} * Twc_i`` (:267).  ``Twc`` is formed as ``KeyFrame::SetPose`` does (Rwc = Rcw^T,
  Ow = -Rwc tcw, float32).  OpenCV's small-matrix gemm accumulates CV_32F products in double and rounds once; that is
  what ``_matmul_f32`` does ("parity-unpinned": OpenCV is not available here to confirm bit equality).
"""
from __future__ import annotations

import glob
import os
import re
from dataclasses import dataclass, field

import numpy as np

from .pack import KeyFramePack

_DT = {"f": np.float32, "d": np.float64, "i": np.int32, "u": np.uint8, "c": np.int8, "w": np.uint16, "s": np.int16}


# ----------------------------------------------------------------------------- cv::FileStorage YAML
def read_cv_yaml(path: str) -> dict:
    """OpenCV FileStorage YAML 1.0 -> dict; ``!!opencv-matrix`` nodes become numpy arrays."""
    import yaml
    Loader = getattr(yaml, "CSafeLoader", yaml.SafeLoader)

    class L(Loader):
        pass

    def mat(loader, node):
        m = loader.construct_mapping(node, deep=True)
        dt = str(m["dt"])
        ch = int(dt[:-1]) if len(dt) > 1 else 1
        a = np.asarray(m["data"], dtype=_DT[dt[-1]])
        return a.reshape(int(m["rows"]), int(m["cols"]), ch) if ch > 1 else a.reshape(int(m["rows"]), int(m["cols"]))
    L.add_constructor("tag:yaml.org,2002:opencv-matrix", mat)
    with open(path) as f:
        text = f.read()
    text = re.sub(r"^%YAML[^\n]*\n", "", text.lstrip())   # "%YAML:1.0" is not a directive PyYAML accepts
    text = re.sub(r"^---\s*\n", "", text)
    return yaml.load(text, Loader=L) or {}


def _fmt(v) -> str:
    if isinstance(v, (bool, np.bool_)):
        return "1" if v else "0"
    if isinstance(v, (int, np.integer)):
        return str(int(v))
    f = float(v)
    if f == int(f) and abs(f) < 1e15:
        return f"{int(f)}."            # FileStorage writes "1." for 1.0
    return repr(f)


def write_cv_yaml(path: str, data: dict) -> None:
    """Writes `data` the way cv::FileStorage lays a YAML file out (header, flow sequences, opencv-matrix nodes):
    used to produce fixtures in the reference's own format."""
    def emit(key, v, ind, out):
        pad = " " * ind
        if isinstance(v, np.ndarray) and v.ndim == 2:
            code = {np.dtype(np.float32): "f", np.dtype(np.float64): "d", np.dtype(np.int32): "i", np.dtype(np.uint8): "u"}[v.dtype]
            out.append(f"{pad}{key}: !!opencv-matrix")
            out.append(f"{pad}   rows: {v.shape[0]}")
            out.append(f"{pad}   cols: {v.shape[1]}")
            out.append(f"{pad}   dt: {code}")
            out.append(f"{pad}   data: [ " + ", ".join(_fmt(x) for x in v.reshape(-1)) + " ]")
        elif isinstance(v, dict):
            out.append(f"{pad}{key}:")
            for k2, v2 in v.items():
                emit(k2, v2, ind + 3, out)
        elif isinstance(v, (list, tuple, np.ndarray)):
            out.append(f"{pad}{key}: [ " + ", ".join(_fmt(x) for x in np.asarray(v).reshape(-1)) + " ]" if len(v) else f"{pad}{key}: []")
        else:
            out.append(f"{pad}{key}: {_fmt(v)}")
    out = ["%YAML:1.0", "---"]
    for k, v in data.items():
        emit(k, v, 0, out)
    with open(path, "w") as f:
        f.write("\n".join(out) + "\n")


# ----------------------------------------------------------------------------- restored objects
@dataclass
class KeyFrameInfo:
    """What KeyFrameConstInfo restores and BAError reads (KeyFrame.cc:31-80)."""
    mnId: int
    mnFrameId: int
    fx: float
    fy: float
    cx: float
    cy: float
    mnMaxX: int
    mnMaxY: int
    keys_un: np.ndarray                 # [N,2] float32 (mvKeysUn[i].pt)
    octave: np.ndarray                  # [N] int32
    Tcw: np.ndarray                     # [4,4] float32
    mpt2kpt: dict = field(default_factory=dict)   # map point id -> keypoint index (mmapMpt2KptId)
    covis_ids: list = field(default_factory=list)  # mvpOrderedConnectedKeyFramesId, best first
    covis_weights: list = field(default_factory=list)
    inv_level_sigma2: np.ndarray | None = None


def read_keyframe(path: str) -> KeyFrameInfo:
    d = read_cv_yaml(path)
    kp = np.asarray(d.get("mvKeysUn", []), dtype=np.float64).reshape(-1, 7)
    ids, kps = list(d.get("mvpMapPointsId", []) or []), list(d.get("mvpCorrKeyPointsId", []) or [])
    if len(ids) != len(kps):
        raise ValueError(f"{path}: mvpMapPointsId and mvpCorrKeyPointsId differ in length (KeyFrame.cc:75)")
    m = {}
    for a, b in zip(ids, kps):          # unordered_map::insert keeps the FIRST value of a repeated key (KeyFrame.cc:76-79)
        m.setdefault(int(a), int(b))
    inv = d.get("mvInvLevelSigma2")
    return KeyFrameInfo(
        mnId=int(d["mnId"]), mnFrameId=int(d["mnFrameId"]), fx=float(d["fx"]), fy=float(d["fy"]), cx=float(d["cx"]), cy=float(d["cy"]),
        mnMaxX=int(d["mnMaxX"]), mnMaxY=int(d["mnMaxY"]), keys_un=kp[:, :2].astype(np.float32), octave=kp[:, 5].astype(np.int32),
        Tcw=np.asarray(d["Pose"], dtype=np.float32).reshape(4, 4), mpt2kpt=m,
        covis_ids=[int(x) for x in (d.get("mvpOrderedConnectedKeyFramesId", []) or [])],
        covis_weights=[int(x) for x in (d.get("mvOrderedWeights", []) or [])],
        inv_level_sigma2=None if inv is None else np.asarray(inv, dtype=np.float32))


def read_keyframe_dir(dirname: str) -> list:
    """Every ``*.yml`` / ``*.yaml`` of the directory (System.cc:646-651), sorted by mnId (System.cc:676)."""
    files = sorted(glob.glob(os.path.join(dirname, "*.yml")) + glob.glob(os.path.join(dirname, "*.yaml")))
    kfs = [read_keyframe(f) for f in files]
    kfs.sort(key=lambda k: k.mnId)
    return kfs


def read_map(path: str):
    """Map.yml -> ({map point id: world position float32[3]}, [keyframe ids])  (Map.cc:162-171, MapPoint.cc:435-449)."""
    d = read_cv_yaml(path)
    pts = {}
    for node in (d.get("mspMapPoints") or {}).values():
        pts[int(node["mnId"])] = np.asarray(node["mWorldPos"], dtype=np.float32).reshape(3)
    return pts, [int(x) for x in (d.get("mspKeyFrameId", []) or [])]


def read_frame_ids(path: str):
    d = read_cv_yaml(path)
    return [int(x) for x in d["mnId"]], [int(x) for x in d["mnFrameId"]]


# ----------------------------------------------------------------------------- flattening
def _matmul_f32(A: np.ndarray, B: np.ndarray) -> np.ndarray:
    """CV_32F * CV_32F as cv::gemm evaluates small matrices: products accumulated in double, one rounding."""
    return (A.astype(np.float64) @ B.astype(np.float64)).astype(np.float32)


def pose_inverse_f32(Tcw: np.ndarray) -> np.ndarray:
    """``Twc`` as KeyFrame::SetPose builds it: Rwc = Rcw^T, Ow = -Rwc * tcw, all CV_32F."""
    Rcw, tcw = Tcw[:3, :3], Tcw[:3, 3:4]
    Rwc = Rcw.T.copy()
    Ow = -_matmul_f32(Rwc, tcw)
    Twc = np.eye(4, dtype=np.float32)
    Twc[:3, :3] = Rwc
    Twc[:3, 3:4] = Ow
    return Twc


def select_covisibles(kf: KeyFrameInfo, num_best_covis: int, min_covis_weight: int) -> list:
    """iba_global.cpp:255-259 -> KeyFrame.cc:417-439."""
    ids = kf.covis_ids
    if num_best_covis > 0:
        return list(ids[:num_best_covis])
    if not ids:
        return []
    w = kf.covis_weights
    n = 0
    while n < len(w) and not (min_covis_weight > w[n]):   # upper_bound with weightComp (a > b) on the descending weights
        n += 1
    if n == len(w):                                       # "it == end()" returns an EMPTY vector (KeyFrame.cc:432-433)
        return []
    return list(ids[:n])


def build_pack(keyframes: list, map_points: dict, scans: list, Twl: np.ndarray, num_best_covis: int = 3, min_covis_weight: int = 150,
               n_covis_slots: int | None = None) -> KeyFramePack:
    """Restored keyframes (sorted by mnId) + map points + one LiDAR scan and odometry pose per keyframe -> KeyFramePack.

    ``scans[i]`` is the [M,3] float32 scan of keyframe i (dataio.read_pointcloud_bin), ``Twl[i]`` its 4x4 fp64 LiDAR pose
    (dataio.lidar_poses_for_keyframes)."""
    from .dataio import hand_eye_lidar_motions
    F = len(keyframes)
    if not (len(scans) == F and len(Twl) == F):
        raise ValueError("one scan and one LiDAR pose per keyframe are required")
    by_id = {k.mnId: i for i, k in enumerate(keyframes)}
    covs = [[c for c in select_covisibles(k, num_best_covis, min_covis_weight) if c in by_id] for k in keyframes]
    C = n_covis_slots if n_covis_slots is not None else max([len(c) for c in covs] + [1])
    if max([len(c) for c in covs] + [0]) > C:
        raise ValueError("more covisible keyframes than slots")
    scan_offset = np.zeros(F + 1, np.int64)
    kp_offset = np.zeros(F + 1, np.int64)
    for i, k in enumerate(keyframes):
        scan_offset[i + 1] = scan_offset[i] + len(scans[i])
        kp_offset[i + 1] = kp_offset[i] + len(k.keys_un)
    NK = int(kp_offset[-1])
    kp_xy = np.zeros((NK, 2), np.float32)
    kp_mp = np.full((NK, 3), np.nan, np.float32)
    covis_uv = np.full((NK, C, 2), np.nan, np.float32)
    covis_valid = np.zeros((F, C), np.uint8)
    relpose = np.tile(np.eye(4, dtype=np.float32)[:3].reshape(-1), (F, C, 1))
    Tcw = np.zeros((F, 12), np.float32)
    he_Tc = np.tile(np.eye(4, dtype=np.float32)[:3].reshape(-1), (F, 1))
    intr = np.zeros((F, 4), np.float32)
    wh = np.zeros((F, 2), np.int32)
    for i, k in enumerate(keyframes):
        o = int(kp_offset[i])
        kp_xy[o:o + len(k.keys_un)] = k.keys_un
        intr[i] = (k.fx, k.fy, k.cx, k.cy)
        wh[i] = (k.mnMaxX, k.mnMaxY)
        Tcw[i] = k.Tcw[:3].reshape(-1)
        Twc = pose_inverse_f32(k.Tcw)
        # keypoint -> map point: the map is walked and inverted (iba_global.cpp:209-212); when two map points claim one
        # keypoint the reference keeps whichever its unordered_map visits last — here the larger map point id
        for mp_id in sorted(k.mpt2kpt):
            kp = k.mpt2kpt[mp_id]
            if mp_id in map_points and 0 <= kp < len(k.keys_un):
                kp_mp[o + kp] = map_points[mp_id]
        if i + 1 < F:
            he_Tc[i] = _matmul_f32(keyframes[i + 1].Tcw, Twc)[:3].reshape(-1)          # iba_global.cpp:267
        for s, cid in enumerate(covs[i]):
            j = by_id[cid]
            other = keyframes[j]
            covis_valid[i, s] = 1
            relpose[i, s] = _matmul_f32(other.Tcw, Twc)[:3].reshape(-1)                # iba_global.cpp:280
            for mp_id, kp in k.mpt2kpt.items():                                        # GetMatchedKptIds (KeyFrame.cc:528-538)
                kp2 = other.mpt2kpt.get(mp_id)
                if kp2 is not None and 0 <= kp < len(k.keys_un) and 0 <= kp2 < len(other.keys_un):
                    covis_uv[o + kp, s] = other.keys_un[kp2]
    he_Tl, he_valid = hand_eye_lidar_motions(np.asarray(Twl, dtype=np.float64))
    scan_xyz = np.concatenate([np.asarray(s, np.float32).reshape(-1, 3) for s in scans]) if F else np.zeros((0, 3), np.float32)
    return KeyFramePack(n_kf=F, n_covis=C, scan_offset=scan_offset, scan_xyz=scan_xyz, intrinsics=intr, image_wh=wh, kp_offset=kp_offset,
                        kp_xy=kp_xy, kp_mappoint=kp_mp, Tcw=Tcw, covis_relpose=relpose, covis_valid=covis_valid, covis_uv=covis_uv,
                        he_Tc=he_Tc, he_Tl=he_Tl, he_valid=he_valid)


def export_pack(keyframe_dir: str, map_file: str, frame_id_file: str, pointcloud_files: list, lidar_pose_file: str,
                num_best_covis: int = 3, min_covis_weight: int = 150, skip: int = 1, only_positive_x: bool = False) -> KeyFramePack:
    """The whole loading stage of iba_global's main (iba_global.cpp:469-511) -> KeyFramePack.
    ``pointcloud_files`` is the sorted listing of the scan directory (``listdir``, io_tools.h)."""
    from .dataio import lidar_poses_for_keyframes, read_pointcloud_bin, read_pose_list
    kfs = read_keyframe_dir(keyframe_dir)
    pts, _ = read_map(map_file)
    _, frame_ids = read_frame_ids(frame_id_file)
    if len(frame_ids) != len(kfs):
        raise ValueError(f"{len(kfs)} keyframes restored but FrameId.yml lists {len(frame_ids)}")
    Twl = lidar_poses_for_keyframes(read_pose_list(lidar_pose_file), frame_ids)
    scans = [read_pointcloud_bin(pointcloud_files[fid], skip=skip, only_positive_x=only_positive_x) for fid in frame_ids]
    return build_pack(kfs, pts, scans, Twl, num_best_covis, min_covis_weight)
