"""Per-phase clocks of K1 under full load: STL_K1_CLK=1 python scripts/k1_clk.py NKF [spread]"""
import importlib, sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
PKG = "spatial-temporal-lidar-camera-calibration_b200"
synth = importlib.import_module(PKG + ".synth")
capi = importlib.import_module(PKG + ".capi")
nkf = int(sys.argv[1]); spread = float(sys.argv[2]) if len(sys.argv) > 2 else 0.2
pack, xgt, _ = synth.generate(n_kf=nkf)
X = synth.candidates(xgt, 4, spread, seed=42)
ctx = capi.Context(); ctx.upload(pack)
for i in range(1, 4):
    ctx.eval_sums(X[i:i + 1])
    ctx.debug_frame(0, 0)
