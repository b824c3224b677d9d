"""A/B of the stand-alone k-NN kernels: warp-per-query (STL_KNN_WARP=1) vs thread-per-query."""
import importlib, sys, time, os
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
PKG = "spatial-temporal-lidar-camera-calibration_b200"
synth = importlib.import_module(PKG + ".synth"); capi = importlib.import_module(PKG + ".capi")
pack, xgt, _ = synth.generate(n_kf=2)
ctx = capi.Context(); ctx.upload(pack)
rng = np.random.default_rng(0)
P = pack.scan_xyz[: int(pack.scan_offset[1])].astype(np.float64)
nq = 400000
q = P[rng.choice(len(P), nq)] + rng.normal(0, 0.03, (nq, 3))
for k, r2 in ((1, 0.0), (30, 0.36)):
    ctx.knn3d(0, q[:1000], k, r2)
    t = time.time(); idx, d2, cnt = ctx.knn3d(0, q, k, r2); dt = time.time() - t
    print("k=%d r2=%.2f: %.1f ms total (incl. H2D/D2H) -> %.1f M queries/s ; mean count %.1f checksum %d" % (k, r2, dt * 1e3, nq / dt / 1e6, cnt.mean(), int(idx[:, 0].astype(np.int64).sum())))
