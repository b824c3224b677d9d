"""Analytic known-answer tests of the oracle, independent of any implementation:
(i) points on a known plane, (ii) a pixel lattice under the identity extrinsic, (iii) finite
differences of the autodiff Jacobians; plus Lie-group identities for Sim3Exp / SE3Log."""
import importlib

import numpy as np
import pytest

PKG = "spatial-temporal-lidar-camera-calibration_b200"


def test_plane_fit_recovers_known_normal(oracle_mod):
    rng = np.random.default_rng(0)
    n = np.array([0.3, -0.5, 0.8]); n /= np.linalg.norm(n)
    a = np.cross(n, [1, 0, 0]); a /= np.linalg.norm(a); b = np.cross(n, a)
    c = np.array([12.0, -7.0, 1.5])
    pts = c + rng.uniform(-0.5, 0.5, (30, 1)) * a + rng.uniform(-0.5, 0.5, (30, 1)) * b
    nn, reg = oracle_mod.plane_fit(pts)
    assert abs(abs(nn @ n) - 1) < 1e-9 and abs(np.linalg.norm(nn) - 1) < 1e-12
    assert reg < 1e-9                      # on-plane points: zero regression error
    noisy = pts + rng.normal(0, 0.05, pts.shape) * n
    _, reg2 = oracle_mod.plane_fit(noisy)
    assert 0.01 < reg2 < 0.2


def test_plane_fit_degenerate_axis_aligned(oracle_mod):
    """All off-diagonal covariances exactly zero -> axis of the smallest diagonal (pointcloud.h:452-460)."""
    pts = np.array([[x, y, 0.0] for x in (-1.0, 1.0) for y in (-2.0, 2.0)])
    n, _ = oracle_mod.plane_fit(pts)
    assert np.allclose(np.abs(n), [0, 0, 1])


def test_sim3exp_is_a_rotation_and_taylor_branch_is_continuous(oracle_mod):
    x = np.array([0.3, -1.1, 0.7, 0.2, -0.4, 0.9, 12.5])
    R, t, s = oracle_mod.sim3exp(x)
    assert np.allclose(R @ R.T, np.eye(3), atol=1e-14) and abs(np.linalg.det(R) - 1) < 1e-14
    assert s == 12.5                                   # the scale is a plain multiplier (g2o_tools.h:138)
    th = np.linalg.norm(x[:3])
    K = np.array([[0, -x[2], x[1]], [x[2], 0, -x[0]], [-x[1], x[0], 0]])
    Rr = np.eye(3) + np.sin(th) / th * K + (1 - np.cos(th)) / th ** 2 * K @ K
    assert np.allclose(R, Rr, atol=1e-14)
    # across the 1e-4 switch (g2o_tools.h:119)
    d = np.array([1.0, 2.0, -2.0]) / 3
    Ra, ta, _ = oracle_mod.sim3exp(np.r_[d * 0.99e-4, 1, 2, 3, 1])
    Rb, tb, _ = oracle_mod.sim3exp(np.r_[d * 1.01e-4, 1, 2, 3, 1])
    assert np.abs(Ra - Rb).max() < 3e-6 and np.abs(ta - tb).max() < 1e-5


def test_se3log_inverts_exp(oracle_mod):
    rng = np.random.default_rng(1)
    for _ in range(50):
        w = rng.normal(size=3); w *= rng.uniform(0.01, 3.0) / np.linalg.norm(w)
        x = np.r_[w, rng.normal(size=3), 1.0]
        R, t, _ = oracle_mod.sim3exp(x)
        assert np.allclose(oracle_mod.se3log(R, t), x[:6], atol=1e-9)
    # near-identity branch |d| > 0.99999
    x = np.r_[1e-4, -2e-4, 1.5e-4, 0.3, 0.2, -0.1, 1.0]
    R, t, _ = oracle_mod.sim3exp(x)
    assert np.allclose(oracle_mod.se3log(R, t), x[:6], atol=1e-9)


def _lattice_pack(pkg):
    """Camera-frame points that project exactly onto an integer pixel lattice under the identity."""
    fx = np.float32(512.0); cx = np.float32(320.0); cy = np.float32(240.0)
    us, vs = np.meshgrid(np.arange(8, 640, 16), np.arange(8, 480, 16))
    us, vs = us.ravel().astype(np.float64), vs.ravel().astype(np.float64)
    z = 4.0 + (np.arange(len(us)) % 5)                          # exactly representable depths
    pts = np.stack([(us - 320.0) / 512.0 * z, (vs - 240.0) / 512.0 * z, z], 1).astype(np.float32)
    behind = pts * np.float32(-1)                                # z < 0: must be culled
    xyz = np.concatenate([pts, behind])
    K = len(us)
    kp = np.stack([us, vs], 1).astype(np.float32)
    kp[::3] += np.float32(1.0)                                  # d^2 = 1   <= 1.5^2 : kept
    kp[1::3] += np.float32(2.0)                                 # d^2 >= 4  > 2.25  : dropped
    nan3 = np.full((K, 3), np.nan, np.float32)
    eye = np.tile(np.array([1, 0, 0, 0, 0, 1, 0, 0, 0, 0, 1, 0], np.float32), (1, 1))
    pack = pkg.KeyFramePack(
        n_kf=1, n_covis=1, scan_offset=[0, len(xyz)], scan_xyz=xyz, intrinsics=[[fx, fx, cx, cy]], image_wh=[[640, 480]],
        kp_offset=[0, K], kp_xy=kp, kp_mappoint=nan3, Tcw=eye, covis_relpose=eye.reshape(1, 1, 12), covis_valid=[[0]],
        covis_uv=np.full((K, 1, 2), np.nan, np.float32), he_Tc=eye, he_Tl=eye.astype(np.float64), he_valid=[0])
    return pack, K


def test_lattice_correspondences_are_known_by_construction(pkg, oracle_mod):
    pack, K = _lattice_pack(pkg)
    o = oracle_mod.Oracle(pack)
    d = o.frame_debug(np.array([0, 0, 0, 0, 0, 0, 1.0]), 0)   # identity extrinsic (Taylor branch)
    want_kp = np.array([k for k in range(K) if k % 3 != 1], np.uint32)
    assert np.array_equal(d["corr_kp"], want_kp)
    assert np.array_equal(d["corr_pt"], want_kp)              # lattice point k matches keypoint k
    assert d["ties"].sum() == 0


def test_stable_variant_rules_on_the_lattice(pkg, oracle_mod):
    """iba_global_stable.cpp: (a) only keypoints that observe a map point are queried, at the map
    point's re-projection (:67-80); (b) a scan point whose ROUNDED pixel is inside stays (:92-94)."""
    pack, K = _lattice_pack(pkg)
    # two extra scan points at u = -0.25 (rounds to 0: kept by the variant only) and u = -0.75 (rounds to -1)
    z = np.float32(4.0)
    extra = np.array([[(-0.25 - 320.0) / 512.0 * 4.0, (100.0 - 240.0) / 512.0 * 4.0, z],
                      [(-0.75 - 320.0) / 512.0 * 4.0, (200.0 - 240.0) / 512.0 * 4.0, z]], np.float32)
    n0 = int(pack.scan_offset[1])
    pack.scan_xyz = np.concatenate([pack.scan_xyz, extra]); pack.scan_offset = np.array([0, n0 + 2])
    # map points: keypoint k (k % 4 == 0) observes the world point that IS lattice point k (Tcw = identity),
    # so its query pixel is the lattice pixel itself whatever kp_xy says; keypoints K-2, K-1 look at the extras
    mp = np.full((K, 3), np.nan, np.float32)
    mp[::4] = pack.scan_xyz[:K][::4]
    mp[K - 2], mp[K - 1] = extra[0], extra[1]
    pack.kp_mappoint = mp
    p = pkg.default_params(); p.variant = 1
    d = oracle_mod.Oracle(pack, params=p).frame_debug(np.array([0, 0, 0, 0, 0, 0, 1.0]), 0)
    want_kp = sorted(set(range(0, K, 4)) | {K - 2})
    assert list(d["corr_kp"]) == want_kp
    assert list(d["corr_pt"]) == [k if k != K - 2 else n0 for k in want_kp]
    # the same pack under iba_global.cpp: kp_xy is used as is and the extras are outside the image
    d0 = oracle_mod.Oracle(pack).frame_debug(np.array([0, 0, 0, 0, 0, 0, 1.0]), 0)
    assert list(d0["corr_kp"]) == [k for k in range(K) if k % 3 != 1]


@pytest.fixture(scope="module")
def assoc(oracle_mod, small_pack, small_candidates):
    pack, _ = small_pack
    o = oracle_mod.Oracle(pack.shard(0, 2))
    nb, _ = o.associate(small_candidates[0])
    assert nb[:3].min() > 0 and nb[3] == 0  # no GPR blocks unless use_gpr
    return o, small_candidates


def test_dual_jacobians_match_central_differences(assoc):
    o, X = assoc
    keys = o.block_keys()
    x = X[1].copy()
    picks = [int(np.flatnonzero(keys[:, 0] == t)[i]) for t in (0, 1, 2) for i in (0, 5)]
    for bi in picks:
        e, J = o.block_eval(bi, x)
        Jn = np.zeros_like(J)
        for a in range(7):
            h = 1e-6 * max(1.0, abs(x[a]))
            xp, xm = x.copy(), x.copy(); xp[a] += h; xm[a] -= h
            ep, _ = o.block_eval(bi, xp, plain=True); em, _ = o.block_eval(bi, xm, plain=True)
            Jn[:, a] = (ep - em) / (2 * h)
        e0, _ = o.block_eval(bi, x, plain=True)
        assert np.allclose(e, e0, rtol=1e-13, atol=1e-13)         # Jet division is f*(1/g): value part within an ulp of plain doubles
        assert np.allclose(J, Jn, rtol=2e-6, atol=2e-6 * max(1.0, np.abs(J).max())), (bi, keys[bi])


def test_linearize_is_sum_of_huber_corrected_blocks(assoc, pkg):
    """cost / g / H re-assembled in numpy from per-block residuals and Jacobians (Ceres semantics)."""
    o, X = assoc
    x = X[2]
    L = o.linearize(x)[0]
    keys = o.block_keys()
    p = pkg.default_params()
    cost, g, H, nres = 0.0, np.zeros(7), np.zeros((7, 7)), 0
    for bi in range(len(keys)):
        e, J = o.block_eval(bi, x)
        delta = p.robust_kernel_delta if keys[bi, 0] == 0 else p.robust_kernel_3ddelta
        s = float(e @ e)
        if s > delta * delta:
            r = np.sqrt(s); rho0, rho1 = 2 * delta * r - delta * delta, delta / r
        else:
            rho0, rho1 = s, 1.0
        cost += 0.5 * rho0
        Jc, ec = np.sqrt(rho1) * J, np.sqrt(rho1) * e
        g += Jc.T @ ec; H += Jc.T @ Jc; nres += len(e)
    assert np.isclose(L[0], cost, rtol=1e-12)
    assert np.allclose(L[1:8], g, rtol=1e-10, atol=1e-9) and np.allclose(L[8:57].reshape(7, 7), H, rtol=1e-10, atol=1e-9)
    assert L[60] == nres and L[57] + L[58] + L[59] == len(keys)
    assert np.allclose(H, H.T) and np.linalg.eigvalsh(H).min() > -1e-6 * np.abs(H).max()


def test_cost_landscape_has_its_minimum_at_the_ground_truth(oracle_mod, small_pack, synth):
    """The 'special points' sanity check of BALoss's constructor (iba_global.cpp:369-374): GT scores lowest."""
    pack, x_gt = small_pack
    o = oracle_mod.Oracle(pack)
    X = synth.candidates(x_gt, 6, spread=1.0)
    s, _, _ = o.ba_error_sums(X)
    f = np.array([sum(o.finalize(r)[:2]) for r in s])
    assert f.argmin() == 0 and (f[1:] > 2 * f[0]).all()


def test_gpr_factor_jacobian_and_depth(oracle_mod, pkg, small_pack, small_candidates):
    """IBA_GPRFactor (IBACalib2.hpp:472-507, TGPR::fit_predict GPR.hpp:449-491): dual-number Jacobian
    through the Cholesky solve vs central differences, and the GP depth of an on-surface keypoint."""
    p = pkg.default_params(); p.use_gpr = 1
    o = oracle_mod.Oracle(small_pack[0].shard(0, 2), params=p)
    nb, _ = o.associate(small_candidates[0])
    assert nb[3] > 10 and nb[0] > 100                     # GPR blocks only where the plane test failed
    keys = o.block_keys()
    x = small_candidates[1].copy()
    for bi in np.flatnonzero(keys[:, 0] == 3)[:4]:
        e, J = o.block_eval(int(bi), x)
        Jn = np.zeros_like(J)
        for a in range(7):
            h = 1e-4 * max(1.0, abs(x[a]))   # K (sigma_n = 1e-10) is ill-conditioned: the function itself carries ~1e-8 noise
            xp, xm = x.copy(), x.copy(); xp[a] += h; xm[a] -= h
            ep, _ = o.block_eval(int(bi), xp, plain=True); em, _ = o.block_eval(int(bi), xm, plain=True)
            Jn[:, a] = (ep - em) / (2 * h)
        assert np.allclose(J, Jn, rtol=5e-3, atol=2e-3 * max(1.0, np.abs(J).max()))
    L = o.linearize(x)[0]
    assert L[61] == nb[3] and L[57] == nb[0]
