"""N4 — the two problems that precede the cost evaluation on the same 7 parameters: the hand-eye initialisation
(include/NLHECalib.hpp) and the calibration bundle adjustment (src/orb_slam/src/Optimizer.cc:65-205,1399-1744).
CPU: known-answer tests of the oracle's restatement.  GPU: the CUDA kernels against the oracle, and the same final
extrinsic out of the stand-in optimiser loop whichever evaluator feeds it."""
import importlib

import numpy as np
import pytest

from conftest import PKG, has_cuda


def _exp(x):
    from oracle import oracle as O
    R, t, s = O.sim3exp(np.asarray(x, dtype=np.float64))
    T = np.eye(4); T[:3, :3] = R; T[:3, 3] = t
    return T, s


def _motions(rng, x_gt, n, noise=0.0, outliers=0):
    """Camera motions Ta (monocular: translation up to scale) and LiDAR motions Tb with Ta X = X Tb, X = Sim3Exp(x_gt)."""
    X, s = _exp(x_gt)
    Ta, Tb = [], []
    for i in range(n):
        w = rng.normal(0, 0.2, 3); u = rng.normal(0, 1.0, 3)
        B, _ = _exp(np.concatenate([w, u, [1.0]]))
        A = X @ B @ np.linalg.inv(X)          # metric camera motion
        A[:3, 3] /= s                          # ... seen by a monocular SLAM at scale 1/s
        if i < outliers:
            A[:3, 3] += rng.normal(0, 0.5, 3)
        A[:3, 3] += rng.normal(0, noise, 3)
        Ta.append(A[:3].reshape(-1)); Tb.append(B[:3].reshape(-1))
    return np.asarray(Ta), np.asarray(Tb)


def _calib_problem(rng, x_gt, n_kf=6, per_kf=200, px_noise=0.3, outlier_frac=0.05):
    """Map points in the first camera frame (monocular scale), LiDAR poses Tlw, observations by the calibEdge model."""
    from oracle import oracle as O
    capi = importlib.import_module(PKG + ".capi")
    intr = np.tile(np.array([718.856, 718.856, 607.1928, 185.2157], np.float32), (n_kf, 1))
    Tq, Xw, obs, off = [], [], [], [0]
    for f in range(n_kf):
        q = np.concatenate([rng.normal(0, 0.05, 3), [0.8 * f, 0.02 * f, 0.0]])
        Tq.append(q)
        for _ in range(per_kf):
            for _try in range(50):
                X = np.array([rng.uniform(-6, 6), rng.uniform(-2, 2), rng.uniform(4, 30)]) / x_gt[6]
                e = O.calib_edge(x_gt, X, q, intr[f], np.zeros(2))     # err = obs - pre with obs = 0  ->  pre = -err
                uv = -e
                if 0 < uv[0] < 1241 and 0 < uv[1] < 376:
                    break
            o = uv + rng.normal(0, px_noise, 2)
            if rng.uniform() < outlier_frac:
                o = o + rng.normal(0, 40, 2)
            Xw.append(X); obs.append(o)
        off.append(len(Xw))
    return capi.CalibBAEdges(off, np.asarray(Tq), intr, np.asarray(Xw), np.asarray(obs), np.full(len(Xw), 0.694, np.float32))


X_GT = np.array([1.21, -1.19, 1.2, 0.05, -0.08, -0.27, 9.0])


def test_rotation_vector_and_hand_eye_edge_known_answers(oracle_mod):
    O = oracle_mod
    rng = np.random.default_rng(0)
    for _ in range(20):
        w = rng.normal(0, 1.0, 3)
        R, _, _ = O.sim3exp(np.concatenate([w, np.zeros(3), [1.0]]))
        assert np.allclose(O.rotvec(R), w, atol=1e-12)
    # exact hand-eye pairs: the residual of EdgeHE vanishes at the true extrinsic (rotation part: R r_b = r_a; translation
    # part: (Ra - I) t + ta s = R tb) and the Jacobian is [hat(R r_b) | Ra - I | ta] as coded
    Ta, Tb = _motions(rng, X_GT, 5)
    for a, b in zip(Ta, Tb):
        e, J = O.he_edge(a, b, X_GT)
        assert np.abs(e).max() < 1e-10
        A = a.reshape(3, 4)
        assert np.allclose(J[:, 3:6], A[:, :3] - np.eye(3)) and np.allclose(J[:, 6], A[:, 3])
        assert np.allclose(J[:, :3], -J[:, :3].T)                      # a skew matrix
    # away from the solution the translation columns ARE the derivative with respect to (tab, s)
    x = X_GT + np.array([0, 0, 0, 0.1, -0.2, 0.05, 0.7])
    e0, J = O.he_edge(Ta[0], Tb[0], x)
    R, t, s = O.sim3exp(x)
    for k in range(3):                                                 # finite difference in tab through the residual formula
        d = np.zeros(3); d[k] = 1e-6
        A = Ta[0].reshape(3, 4)
        fd = ((A[:, :3] - np.eye(3)) @ d) / 1e-6
        assert np.allclose(J[:, 3 + k], fd, atol=1e-9)


def test_calib_edge_model_and_autodiff(oracle_mod):
    """calibEdge = project(Tcl * Tlw * Tcl^-1 * (s Xw)): checked against matrices, and its Dual<7> Jacobian against central
    differences of the plain-double evaluation."""
    O = oracle_mod
    rng = np.random.default_rng(1)
    ed = _calib_problem(rng, X_GT, n_kf=2, per_kf=5, px_noise=0.0, outlier_frac=0.0)
    Tcl, s = _exp(X_GT)
    Tcl[:3, 3] = X_GT[3:6]            # calibEdge's vertex holds the translation itself (p_tcl.segment<3>(3) = tcl), not upsilon
    L, chi2 = O.calib_linearize(ed, X_GT)
    assert chi2.max() < 1e-16 and L[0, 0] < 1e-14                       # noise-free observations: zero error at the truth
    for f in range(2):
        Tlw, _ = _exp(np.concatenate([ed.Tlw_quat[f], [1.0]]))
        Tlw[:3, 3] = ed.Tlw_quat[f, 3:]                                  # Tlw_quat carries the translation itself, not upsilon
        for i in range(ed.edge_offset[f], ed.edge_offset[f + 1]):
            P = Tcl @ Tlw @ np.linalg.inv(Tcl) @ np.append(s * ed.Xw[i], 1.0)
            uv = np.array([718.856 * P[0] / P[2] + 607.1928, 718.856 * P[1] / P[2] + 185.2157])
            assert np.allclose(uv, ed.obs[i], atol=2e-4)                 # float32 intrinsics
    x = X_GT + rng.normal(0, 0.01, 7)
    L, _ = O.calib_linearize(ed, x)
    g = L[0, 1:8]
    num = np.zeros(7)
    for k in range(7):
        h = 1e-6
        xp, xm = x.copy(), x.copy(); xp[k] += h; xm[k] -= h
        cp = O.calib_linearize(ed, xp)[0][0, 0]; cm = O.calib_linearize(ed, xm)[0][0, 0]
        num[k] = (cp - cm) / (2 * h) / 2.0                               # cost = sum rho(chi2), g = J^T rho' Omega e = d(cost)/dx / 2
    assert np.allclose(g, num, rtol=1e-4, atol=1e-5 * np.abs(num).max())


def test_hand_eye_problem_properties(oracle_mod, pkg):
    """What holds for the hand-eye problem AS CODED in NLHECalib.hpp: with exact motion pairs the true extrinsic has zero
    cost and zero gradient; the stand-in loop never increases the cost; the line-process weights are mu / (mu + chi2).
    (Its Jacobian [hat(R r_b) | Ra - I | ta] is not the derivative of the residual for the additive Sim3 update, so
    Gauss-Newton on it stalls short of the truth from a coarse start — in the reference as well; the drop-in reproduces
    the edges, it does not repair them.)"""
    host = importlib.import_module(PKG + ".host")
    capi = importlib.import_module(PKG + ".capi")
    rng = np.random.default_rng(2)
    Ta, Tb = _motions(rng, X_GT, 120)
    ed = capi.HandEyeEdges(Ta, Tb, huber_delta=0.1, regulation=0.0)
    L, chi2 = oracle_mod.he_linearize(ed, X_GT)
    assert L[0, 0] < 1e-18 and np.abs(L[0, 1:8]).max() < 1e-9 and chi2.max() < 1e-18 and L[0, 57] == 120 and L[0, 60] == 360
    he = host.HandEyeInit(lambda e, x: oracle_mod.he_linearize(e, x))
    x0 = X_GT + np.array([0.01, -0.01, 0.01, 0.05, 0.05, -0.05, 1.0])
    c0 = oracle_mod.he_linearize(ed, x0)[0][0, 0]
    x1, c1 = he.robust_kernel(ed, x0, iters=30)
    assert c1 < 0.5 * c0
    x2, c2 = he.line_process(ed, x0)
    assert np.isfinite(x2).all() and c2 <= c0
    # regularisation edge: adds regulation * |x[3:6]|^2 to the cost and regulation * I to the translation block of H
    ed_r = capi.HandEyeEdges(Ta, Tb, huber_delta=0.1, regulation=0.6)
    Lr, _ = oracle_mod.he_linearize(ed_r, x0)
    L0, _ = oracle_mod.he_linearize(ed, x0)
    assert np.isclose(Lr[0, 0] - L0[0, 0], 0.6 * (x0[3:6] ** 2).sum(), rtol=1e-12)
    dH = (Lr[0, 8:57] - L0[0, 8:57]).reshape(7, 7)
    want = np.zeros((7, 7)); want[3:6, 3:6] = 0.6 * np.eye(3)
    assert np.allclose(dH, want, atol=1e-9)


@pytest.mark.gpu
@pytest.mark.skipif(not has_cuda(), reason="no CUDA device")
def test_gpu_edges_match_the_oracle_and_give_the_same_extrinsic(oracle_mod, pkg):
    host = importlib.import_module(PKG + ".host")
    capi = importlib.import_module(PKG + ".capi")
    rng = np.random.default_rng(3)
    Ta, Tb = _motions(rng, X_GT, 300, noise=0.002, outliers=40)
    he_ed = capi.HandEyeEdges(Ta, Tb, info=rng.uniform(0.2, 1.0, 300), huber_delta=0.1, regulation=300 * 0.005)
    ca_ed = _calib_problem(rng, X_GT, n_kf=8, per_kf=400)
    ca_ed.level = (rng.uniform(size=ca_ed.n_edges) < 0.1).astype(np.uint8)
    X = X_GT + rng.normal(0, 0.02, (5, 7))
    X[4, :3] = 0.0                                                       # theta == 0: the first-order branches of calibEdge
    with capi.Context() as c:
        for ed, lin_g, lin_o in ((he_ed, c.he_linearize, oracle_mod.he_linearize), (ca_ed, c.calib_linearize, oracle_mod.calib_linearize)):
            got, chi_g = lin_g(ed, X, want_chi2=True)
            want, chi_o = lin_o(ed, X)
            assert np.array_equal(got[:, 57:], want[:, 57:])
            assert np.allclose(chi_g, chi_o, rtol=1e-9, atol=1e-12)
            assert np.allclose(got[:, 0], want[:, 0], rtol=1e-10)
            sg = np.abs(want[:, 1:8]).max(axis=1, keepdims=True); sh = np.abs(want[:, 8:57]).max(axis=1, keepdims=True)
            assert np.allclose(got[:, 1:8], want[:, 1:8], rtol=1e-8, atol=1e-10 * sg)
            assert np.allclose(got[:, 8:57], want[:, 8:57], rtol=1e-8, atol=1e-10 * sh)
        # final extrinsic: same loop, GPU edges vs oracle edges (0.01 deg / 0.1 cm)
        x0 = X_GT + np.array([0.04, -0.03, 0.02, 0.1, 0.05, -0.1, 1.5])
        he_ed.info = None
        xg, _ = host.HandEyeInit(lambda e, x: c.he_linearize(e, x, want_chi2=True)).line_process(he_ed, x0)
        xo, _ = host.HandEyeInit(lambda e, x: oracle_mod.he_linearize(e, x)).line_process(he_ed, x0)
        assert np.abs(xg[:3] - xo[:3]).max() < np.deg2rad(0.01) and np.abs(xg[3:6] - xo[3:6]).max() < 1e-3
        ca_ed.level = None
        x0 = X_GT + np.array([0.004, -0.003, 0.002, 0.02, 0.01, -0.02, 0.2])
        yg, ng = host.calib_ba(lambda e, x: c.calib_linearize(e, x, want_chi2=True), ca_ed, x0)
        yo, no = host.calib_ba(lambda e, x: oracle_mod.calib_linearize(e, x), ca_ed, x0)
        assert ng == no and ng > 0.9 * ca_ed.n_edges * 0.9
        assert np.abs(yg[:3] - yo[:3]).max() < np.deg2rad(0.01) and np.abs(yg[3:6] - yo[3:6]).max() < 1e-3
        assert np.abs(yg[:3] - X_GT[:3]).max() < 0.01                      # and it is the right extrinsic
