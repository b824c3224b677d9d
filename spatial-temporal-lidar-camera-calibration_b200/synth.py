"""Seeded synthetic KITTI-shaped packs — wrapper of csrc/synth.cpp."""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import _abi
from .pack import KeyFramePack

# BASELINE.json configs -> generator settings (candidates B are the caller's business)
CONFIGS = {
    "c1": dict(n_kf=50, beams=64, az_steps=1875, n_kp=2000),      # 50-KF slice, CPU-runnable
    "c2": dict(n_kf=1500, beams=64, az_steps=1875, n_kp=2000),    # full KITTI-00 shape
    "c5": dict(n_kf=4000, beams=128, az_steps=2048, n_kp=2000),   # GPR stress shape
}


def make_cfg(**kw) -> _abi.SynthCfg:
    lib = _abi.load_synth()
    cfg = _abi.SynthCfg()
    lib.stl_synth_default_cfg(C.byref(cfg))
    if "n_kf" in kw and "n_kf_total" not in kw:
        kw["n_kf_total"] = kw["n_kf"] + kw.get("kf_begin", 0)
    for k, v in kw.items():
        if k == "x_gt":
            for i in range(7):
                cfg.x_gt[i] = float(v[i])
        else:
            setattr(cfg, k, v)
    return cfg


def generate(**kw):
    """Returns (KeyFramePack, x_gt[7], Twl[F,12]).  Keyword args override stl_synth_cfg fields."""
    lib = _abi.load_synth()
    cfg = make_cfg(**kw)
    h = lib.stl_synth_create(C.byref(cfg))
    if not h:
        raise ValueError("stl_synth_create rejected the configuration")
    try:
        pack = KeyFramePack.from_c(lib.stl_synth_pack(h).contents, copy=True)
        x_gt = np.ctypeslib.as_array(lib.stl_synth_x_gt(h), (7,)).copy()
        twl = np.ctypeslib.as_array(lib.stl_synth_Twl(h), (cfg.n_kf * 12,)).copy().reshape(cfg.n_kf, 12)
    finally:
        lib.stl_synth_destroy(h)
    return pack, x_gt, twl


def candidates(x_gt, B: int, spread: float = 0.5, seed: int = 42) -> np.ndarray:
    """[B,7] candidates: row 0 = x_gt, others x_gt + U(lb,ub)*spread
    (lb/ub of config/calib/00/iba_calib_global.yml:38-39)."""
    lib = _abi.load_synth()
    xg = np.ascontiguousarray(x_gt, dtype=np.float64)
    out = np.empty((B, 7), dtype=np.float64)
    lib.stl_synth_candidates(xg.ctypes.data_as(_abi._dp), seed, B, spread, out.ctypes.data_as(_abi._dp))
    return out
