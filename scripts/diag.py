"""First-light diagnostic: GPU vs oracle on a small pack (run under gpurun)."""
import importlib, sys, time, os
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
PKG = "spatial-temporal-lidar-camera-calibration_b200"
synth = importlib.import_module(PKG + ".synth")
capi = importlib.import_module(PKG + ".capi")
from oracle import oracle as O

nkf = int(sys.argv[1]) if len(sys.argv) > 1 else 6
B = int(sys.argv[2]) if len(sys.argv) > 2 else 4
pack, xgt, _ = synth.generate(n_kf=nkf)
X = synth.candidates(xgt, B, 0.5)
orc = O.Oracle(pack, kind="best")
t = time.time(); so, ties, cnt = orc.ba_error_sums(X, mode=1); print("oracle %.3fs ties %s" % (time.time() - t, ties))
ctx = capi.Context(); t = time.time(); ctx.upload(pack); print("upload %.3fs" % (time.time() - t))
ctx.set_profiling(True)
t = time.time(); sg = ctx.eval_sums(X); print("gpu eval %.4fs" % (time.time() - t))
t = time.time(); sg = ctx.eval_sums(X); print("gpu eval(2) %.4fs" % (time.time() - t))
print(ctx.stage_stats())
np.set_printoptions(linewidth=220, precision=9)
print("oracle\n", so); print("gpu\n", sg)
rel = np.abs(sg - so) / np.maximum(np.abs(so), 1e-300)
print("max rel", rel.max(0))
bad = 0
for b in range(min(B, 2)):
    for kf in range(nkf):
        d = orc.frame_debug(X[b], kf)
        kp, pt = ctx.debug_corrset(b, kf)
        ok = np.array_equal(kp, d["corr_kp"]) and np.array_equal(pt, d["corr_pt"])
        al = ctx.debug_align(b, kf)
        ok2 = np.array_equal(al["nn"], d["align_nn"]) and np.array_equal(al["m"], d["align_m"]) and np.array_equal(al["is_plane"], d["align_is_plane"])
        ok3 = all(np.array_equal(al["knn"][i][:al["m"][i]], d["align_knn"][i][:d["align_m"][i]]) for i in range(len(al["m"]))) if ok2 else False
        derr = np.abs(al["dist"] - d["align_dist"]).max() if ok2 and len(al["dist"]) else -1
        print(f"b{b} kf{kf}: corr n={len(kp)}/{len(d['corr_kp'])} ok={ok}  align n={len(al['nn'])}/{len(d['align_nn'])} idx={ok2} knn={ok3} max|ddist|={derr:.3e}")
        bad += (not ok) + (not ok2) + (not ok3)
print("BAD", bad)
for kf in range(min(nkf, 3)):
    print("frame", kf, "gpu", ctx.debug_frame(0, kf)); print("   oracle", orc.frame_sums(X[0], kf))
