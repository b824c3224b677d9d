"""GPU parity of the LM path: stl_associate (BuildProblem, iba_local.cpp:145-323) and
stl_linearize_batch (plane / point-to-point / point-to-plane factors + Huber, cost, J^T r, J^T J)."""
import importlib

import numpy as np
import pytest

from conftest import PKG

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def orc(oracle_mod, small_pack):
    return oracle_mod.Oracle(small_pack[0], kind="best")


def test_association_block_counts_and_linearisation(gpu_ctx, orc, small_candidates):
    x0 = small_candidates[0]
    nb_o, ties = orc.associate(x0)
    assert ties.sum() == 0
    nb_g = gpu_ctx.associate(x0)
    assert np.array_equal(nb_g, nb_o), (nb_g, nb_o)
    want = orc.linearize(small_candidates)
    got = gpu_ctx.linearize(small_candidates)
    assert np.array_equal(got[:, 57:], want[:, 57:])  # block / residual counts
    assert np.allclose(got[:, 0], want[:, 0], rtol=1e-9, atol=0)  # cost
    scale_g = np.abs(want[:, 1:8]).max(axis=1, keepdims=True)
    scale_h = np.abs(want[:, 8:57]).max(axis=1, keepdims=True)
    assert np.allclose(got[:, 1:8], want[:, 1:8], rtol=1e-6, atol=1e-9 * scale_g)     # J^T r
    assert np.allclose(got[:, 8:57], want[:, 8:57], rtol=1e-6, atol=1e-9 * scale_h)   # J^T J
    H = got[0, 8:57].reshape(7, 7)
    assert np.array_equal(H, H.T)


def test_reassociation_at_another_estimate(gpu_ctx, orc, small_candidates):
    x1 = small_candidates[2]
    nb_o, _ = orc.associate(x1)
    nb_g = gpu_ctx.associate(x1)
    assert np.array_equal(nb_g, nb_o)
    want, got = orc.linearize(x1), gpu_ctx.linearize(x1)
    assert np.allclose(got[:, 0], want[:, 0], rtol=1e-9) and np.array_equal(got[:, 57:], want[:, 57:])


def test_gradient_matches_finite_difference_of_gpu_cost(gpu_ctx, small_candidates):
    """Blocks stay frozen (like inside ceres::Solve): d cost / d x == J^T r for residuals inside the
    Huber band; checked on the GPU numbers themselves."""
    gpu_ctx.associate(small_candidates[0])
    x = small_candidates[0].copy()
    g = gpu_ctx.linearize(x)[0, 1:8]
    num = np.zeros(7)
    for a in range(7):
        h = 1e-6 * max(1.0, abs(x[a]))
        xp, xm = x.copy(), x.copy(); xp[a] += h; xm[a] -= h
        num[a] = (gpu_ctx.linearize(xp)[0, 0] - gpu_ctx.linearize(xm)[0, 0]) / (2 * h)
    assert np.allclose(g, num, rtol=2e-4, atol=1e-4 * np.abs(g).max())


def test_lm_steps_reduce_the_cost_like_the_oracle(gpu_ctx, orc, small_candidates, small_pack):
    """A deterministic in-repo stand-in for the (absent) Ceres LM loop, run identically over the GPU
    and the oracle evaluators: same iterates to 1e-6 => same final extrinsic (0.01 deg / 0.1 cm)."""
    host = importlib.import_module(PKG + ".host")
    x0 = small_candidates[1].copy()
    prob = host.LMProblem(gpu_ctx)
    prob.build(x0)
    orc.associate(x0)
    xg, xo = x0.copy(), x0.copy()
    costs = []
    for _ in range(4):
        xg, cg = prob.lm_step(xg, 1e-3)
        L = orc.linearize(xo)[0]
        H, g = L[8:57].reshape(7, 7), L[1:8]
        xo = xo - np.linalg.solve(H + 1e-3 * np.diag(np.maximum(np.diag(H), 1e-12)), g)
        costs.append(cg)
    assert costs[-1] < costs[0]
    assert np.abs(xg[:3] - xo[:3]).max() < np.deg2rad(0.01) and np.abs(xg[3:6] - xo[3:6]).max() < 1e-3


def test_gpr_factor_blocks(pkg, oracle_mod, small_pack, small_candidates):
    """use_gpr: non-planar neighbourhoods get an IBA_GPRFactor (the branch the reference keeps commented
    out, iba_local.cpp:272-280); GPU derivative by the adjoint identity vs the oracle's Jet-Cholesky."""
    capi = importlib.import_module(PKG + ".capi")
    p = pkg.default_params(); p.use_gpr = 1
    pack = small_pack[0].shard(0, 3)
    orc = oracle_mod.Oracle(pack, params=p, kind="best")
    nb_o, _ = orc.associate(small_candidates[0])
    want = orc.linearize(small_candidates[:3])
    with capi.Context(params=p) as c:
        c.upload(pack)
        nb_g = c.associate(small_candidates[0])
        got = c.linearize(small_candidates[:3])
    assert np.array_equal(nb_g, nb_o) and nb_o[3] > 20
    assert np.array_equal(got[:, 57:], want[:, 57:])
    # the kernel matrix (sigma^2 = 100, sigma_n = 1e-10, neighbours a few pixels apart) has a condition number
    # around 1e12, so two correct evaluations agree to ~1e-8: the north-star tolerance (1e-6 relative) applies
    assert np.allclose(got[:, 0], want[:, 0], rtol=1e-6, atol=0)
    sg = np.abs(want[:, 1:8]).max(axis=1, keepdims=True); sh = np.abs(want[:, 8:57]).max(axis=1, keepdims=True)
    assert np.allclose(got[:, 1:8], want[:, 1:8], rtol=1e-5, atol=1e-6 * sg)
    assert np.allclose(got[:, 8:57], want[:, 8:57], rtol=1e-5, atol=1e-6 * sh)
