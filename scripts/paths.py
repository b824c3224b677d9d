"""Which search path the K2a queries take: STL_DEBUG_STATS=1 python scripts/paths.py NKF [spread]"""
import importlib, sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
PKG = "spatial-temporal-lidar-camera-calibration_b200"
synth = importlib.import_module(PKG + ".synth")
capi = importlib.import_module(PKG + ".capi")
nkf = int(sys.argv[1]); spread = float(sys.argv[2]) if len(sys.argv) > 2 else 0.2
pack, xgt, _ = synth.generate(n_kf=nkf)
X = synth.candidates(xgt, 4, spread, seed=42)
ctx = capi.Context(); ctx.upload(pack)
ctx.eval_sums(X[1:2]); ctx.debug_frame(0, 0)      # allocates the debug buffers (counters start at zero)
for i in range(1, 4):
    ctx.eval_sums(X[i:i + 1])
ctx.debug_frame(0, 0)
