"""Generates the committed golden fixtures (tests/golden/*.npz).

Run HERE (needs /root/reference for oracle/_ref): every expected value is produced by the oracle
whose neighbour searches go through the reference's real vendored nanoflann v1.5.0.  The fixtures
carry their own inputs, so the GPU box needs neither /root/reference nor a bit-identical libm.

    python tests/golden/make_golden.py
"""
import importlib
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
PKG = "spatial-temporal-lidar-camera-calibration_b200"
synth = importlib.import_module(PKG + ".synth")
from oracle import oracle as O  # noqa: E402

CASES = {
    # name: generator overrides, number of candidates, spread, stl_params_t overrides
    "kitti_small": (dict(n_kf=3, beams=32, az_steps=900, n_kp=500, seed=21), 3, 0.5, {}),
    "sparse_ragged": (dict(n_kf=4, beams=16, az_steps=600, n_kp=300, seed=22, anchored_frac=0.9, mappoint_frac=0.8), 3, 0.3, {}),
    # iba_global_stable.cpp flavour (no LM part: iba_local has no such variant)
    "stable_small": (dict(n_kf=3, beams=32, az_steps=900, n_kp=500, seed=23), 3, 0.4, {"variant": 1, "min_diff_dist": 0.45}),
    # IBA_GPRFactor blocks in the LM problem
    "gpr_small": (dict(n_kf=2, beams=32, az_steps=900, n_kp=500, seed=24), 2, 0.3, {"use_gpr": 1}),
}
PARAM_KEYS = ("variant", "min_diff_dist", "use_gpr")  # stored in every fixture as p_<key>


def build_case(name, cfg, B, spread, pover):
    pkg = importlib.import_module(PKG)
    params = pkg.default_params()
    for k, v in pover.items():
        setattr(params, k, v)
    pack, x_gt, _ = synth.generate(**cfg)
    if name == "sparse_ragged":  # ragged edge cases: an empty scan, a keyframe without keypoints' map points
        so = pack.scan_offset.copy()
        n1 = int(so[2] - so[1])
        keep = np.ones(pack.n_points, bool)
        keep[so[1]:so[2]] = False
        pack.scan_xyz = pack.scan_xyz[keep]
        so[2:] -= n1
        pack.scan_offset = so
        k0, k1 = int(pack.kp_offset[2]), int(pack.kp_offset[3])
        pack.kp_mappoint[k0:k1] = np.nan
    X = synth.candidates(x_gt, B, spread)
    orc = O.Oracle(pack, params=params, kind="ref")
    sums, ties, cnt = orc.ba_error_sums(X, mode=0)
    assert ties.sum() == 0, "golden data must be tie-free"
    out = dict(pack.to_npz_dict())
    out.update(X=X, x_gt=x_gt, sums=sums, counters=cnt)
    for k in PARAM_KEYS:
        out["p_" + k] = np.float64(getattr(params, k))
    for b in range(2):
        for kf in range(pack.n_kf):
            d = orc.frame_debug(X[b], kf)
            fs = orc.frame_sums(X[b], kf)
            pre = f"b{b}_kf{kf}_"
            for key in ("corr_kp", "corr_pt", "align_kp", "align_nn", "align_m", "align_is_plane", "align_dist", "align_knn"):
                out[pre + key] = d[key]
            out[pre + "frame"] = np.array(list(fs.values()))
    if params.variant == 0:
        nb, lties = orc.associate(X[0])
        assert lties.sum() == 0
        out["lm_nblocks"] = nb
        out["lm_keys"] = orc.block_keys()
        out["lm_lin"] = orc.linearize(X)
    else:
        nb = None
    np.savez_compressed(os.path.join(HERE, name + ".npz"), **out)
    print(name, "points", pack.n_points, "sums[0]", sums[0][:4], "blocks", nb)


if __name__ == "__main__":
    assert O.have_ref(), "build oracle/_ref first (make -C oracle)"
    for name, (cfg, B, spread, pover) in CASES.items():
        build_case(name, cfg, B, spread, pover)
