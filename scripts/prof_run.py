"""Small driver for ncu / timing runs: python scripts/prof_run.py NKF B REPS"""
import importlib, sys, time, os
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
PKG = "spatial-temporal-lidar-camera-calibration_b200"
synth = importlib.import_module(PKG + ".synth")
capi = importlib.import_module(PKG + ".capi")
nkf = int(sys.argv[1]); B = int(sys.argv[2]); reps = int(sys.argv[3]) if len(sys.argv) > 3 else 3
t = time.time(); pack, xgt, _ = synth.generate(n_kf=nkf); print("synth %.2fs, %d pts" % (time.time() - t, pack.n_points))
X = synth.candidates(xgt, B, 0.5)
ctx = capi.Context(); t = time.time(); ctx.upload(pack); print("upload %.3fs" % (time.time() - t))
ctx.set_profiling(True)
ctx.eval_sums(X); ctx.stage_stats()
t = time.time()
for _ in range(reps): s = ctx.eval_sums(X)
dt = (time.time() - t) / reps
st = ctx.stage_stats()
print("eval %.3f ms/call  -> %.1f evals/s" % (dt * 1e3, B / dt))
for k, (ms, n) in st.items():
    if n: print("  %-10s %.3f ms/launch x %d" % (k, ms / n, n))
wc = ctx.work_counters()
k1 = st["assoc2d"][0] / max(st["assoc2d"][1], 1) * 1e-3
print("K1 achieved GB/s (algorithmic): %.1f" % (wc["k1_bytes"] / k1 / 1e9))
k2 = st["knn3d"][0] / max(st["knn3d"][1], 1) * 1e-3
print("K2 queries/s: %.3e (nn %.0f + knn %.0f per call)" % ((wc["q3d_nn"] + wc["q3d_knn"]) / k2, wc["q3d_nn"], wc["q3d_knn"]))
print(s[0])
ctx.associate(X[0]); ctx.linearize(X[:1]); ctx.stage_stats()
t = time.time()
for _ in range(reps):
    nb = ctx.associate(X[0]); L = ctx.linearize(X[:1])
dt = (time.time() - t) / reps
st = ctx.stage_stats()
print("associate+linearize %.3f ms/call, blocks %s" % (dt * 1e3, nb.tolist()))
for k, (ms, n) in st.items():
    if n: print("  %-10s %.3f ms/launch x %d" % (k, ms / n, n))
print(L[0][:3])
