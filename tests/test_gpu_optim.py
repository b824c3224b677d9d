"""Final-extrinsic parity (BASELINE north star: 0.01 deg / 0.1 cm after the unchanged optimiser).

NOMAD and Ceres are absent from this image, so the optimiser is the deterministic stand-in of
optim.py, run once over the GPU evaluator (through the host mirror and the C-ABI) and once over the
CPU oracle.  The bar is on the FINAL estimate, not on the individual evaluations."""
import importlib

import numpy as np
import pytest

PKG = "spatial-temporal-lidar-camera-calibration_b200"
pytestmark = pytest.mark.gpu

LB = np.array([-0.1, -0.1, -0.1, -0.3, -0.3, -0.3, -1.0])   # config/calib/00/iba_calib_global.yml:38-39
UB = -LB
ROT_TOL, TRANS_TOL = np.deg2rad(0.01), 1e-3                   # 0.01 deg, 0.1 cm


@pytest.fixture(scope="module")
def orc(oracle_mod, small_pack):
    return oracle_mod.Oracle(small_pack[0], kind="best")


class OracleLoss:
    """BALoss over the CPU oracle (iba_global.cpp:377-396)."""

    def __init__(self, orc, params):
        self.orc, self.p = orc, params

    def eval_block(self, X):
        sums, _, _ = self.orc.ba_error_sums(np.atleast_2d(X), mode=1)
        out = []
        for r in sums:
            f1, f2, C, valid, cnt = self.orc.finalize(r)
            out.append([self.p.err_weight[0] * f1 + self.p.err_weight[1] * f2, C - self.p.he_threshold,
                        -C - self.p.he_threshold, self.p.valid_rate - valid / (cnt + 1)])
        return out


class OracleProblem:
    def __init__(self, orc):
        self.orc = orc

    def build(self, x):
        return self.orc.associate(x)[0]

    def evaluate(self, x):
        s = self.orc.linearize(np.asarray(x, dtype=np.float64).reshape(1, 7))
        return float(s[0, 0]), s[0, 1:8].copy(), s[0, 8:57].reshape(7, 7).copy()


def _close(xa, xb):
    return np.abs(xa[:3] - xb[:3]).max() < ROT_TOL and np.abs(xa[3:6] - xb[3:6]).max() < TRANS_TOL and abs(xa[6] - xb[6]) < 1e-3


def test_poll_search_reaches_the_same_extrinsic(pkg, oracle_mod, small_pack, small_candidates):
    capi = importlib.import_module(PKG + ".capi")
    host = importlib.import_module(PKG + ".host")
    optim = importlib.import_module(PKG + ".optim")
    pack, x_gt = small_pack[0].shard(0, 3), small_pack[1]
    p = pkg.default_params()
    x0 = small_candidates[2]
    lb, ub = x_gt + LB, x_gt + UB
    orc = oracle_mod.Oracle(pack, params=p, kind="best")
    xo, bo, no, ho = optim.poll_search(OracleLoss(orc, p), x0, lb, ub, max_bb_eval=71, init_frame=0.25 * (UB - LB))
    with capi.Context(params=p) as c:
        c.upload(pack)
        xg, bg, ng, hg = optim.poll_search(host.BALoss(c), x0, lb, ub, max_bb_eval=71, init_frame=0.25 * (UB - LB))
    assert ng == no and _close(xg, xo)
    assert np.allclose(bg, bo, rtol=1e-6, atol=1e-9)
    assert hg[-1][1] < hg[0][1] or hg[-1][2] < hg[0][2]      # the search made progress


def test_lm_refine_reaches_the_same_extrinsic(gpu_ctx, orc, small_candidates):
    host = importlib.import_module(PKG + ".host")
    optim = importlib.import_module(PKG + ".optim")
    x0 = small_candidates[1]
    xg, cg = optim.lm_refine(host.LMProblem(gpu_ctx), x0, max_iba_iter=3, max_num_iterations=8)
    xo, co = optim.lm_refine(OracleProblem(orc), x0, max_iba_iter=3, max_num_iterations=8)
    assert len(cg) == len(co) and _close(xg, xo)
    assert np.allclose(cg, co, rtol=1e-6)
    assert cg[-1] < cg[0] * 1.0 + 1e-9
