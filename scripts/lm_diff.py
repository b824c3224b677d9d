"""Diagnostic (run under gpurun): per-block difference between the GPU's frozen LM problem and the oracle's at
the bench candidates.  python scripts/lm_diff.py NKF NCAND [strict]"""
import importlib, sys, os, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
PKG = "spatial-temporal-lidar-camera-calibration_b200"
pkg = importlib.import_module(PKG)
synth = importlib.import_module(PKG + ".synth")
capi = importlib.import_module(PKG + ".capi")
from oracle import oracle as O

nkf = int(sys.argv[1]) if len(sys.argv) > 1 else 200
ncand = int(sys.argv[2]) if len(sys.argv) > 2 else 6
strict = len(sys.argv) > 3
pack, xgt, _ = synth.generate(n_kf=nkf, n_kf_total=1500, seed=1000)
X = synth.candidates(xgt, ncand + 1, 0.2, seed=42)[1:]
orc = O.Oracle(pack, kind="best")
with capi.Context() as c:
    c.upload(pack)
    for b in range(ncand):
        x = X[b]
        nb_o, ties = orc.associate(x, strict=strict)
        Lo = orc.linearize(x[None], nthreads=0)[0]
        nb_g = c.associate(x)
        Lg = c.linearize(x[None])[0]
        rel = abs(Lg[0] - Lo[0]) / abs(Lo[0])
        print(f"cand {b}: blocks gpu {nb_g.tolist()} oracle {nb_o.tolist()} ties {ties.tolist()} cost rel {rel:.2e}")
        if rel < 1e-11:
            continue
        B = c.eval_blocks(x)
        keys_o = orc.block_keys()
        idx = {(int(t), int(f), int(k)): i for i, (t, f, k) in enumerate(zip(B["type"], B["kf"], B["kp"]))}
        worst = []
        for i, (t, f, k) in enumerate(keys_o):
            e, J = orc.block_eval(i, x)
            j = idx.get((int(t), int(f), int(k)))
            if j is None:
                print("   block missing on the GPU:", t, f, k)
                continue
            eg = B["residuals"][j][: len(e)]
            d = np.abs(eg - e).max()
            if d > 1e-9 * max(1.0, np.abs(e).max()):
                worst.append((d, int(t), int(f), int(k), e.copy(), eg.copy()))
        worst.sort(key=lambda w: -w[0])
        print(f"   {len(worst)} blocks differ")
        for w in worst[:6]:
            print("   type %d kf %d kp %d  maxdiff %.3e\n      oracle %s\n      gpu    %s" % (w[1], w[2], w[3], w[0], w[4], w[5]))
        break
