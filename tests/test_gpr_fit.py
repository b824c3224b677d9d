"""a21 — GPR::fit (GPR.hpp:350-387) and its objective GPRHyperLoss::Evaluate (GPR.hpp:154-174).

Host side of the boundary (the reference keeps the fit on the CPU): the objective is reproduced exactly, its
as-coded gradient expressions too; the derivative of the objective is checked by finite differences; the fit
itself (the reference runs Ceres' L-BFGS, which is not available) is checked for what any minimiser of this
objective must do.  The GPU test runs the fit per factor inside stl_associate (params.gpr_optimize)."""
import importlib

import numpy as np
import pytest

from conftest import PKG, has_cuda


def _sample(seed, n=24):
    """Neighbour pixels a few pixels apart and depths on a gently curved surface + noise."""
    rng = np.random.default_rng(seed)
    X = np.array([600.0, 180.0]) + rng.uniform(-12, 12, (n, 2))
    y = 9.0 + 0.03 * (X[:, 0] - 600) - 0.02 * (X[:, 1] - 180) + 0.002 * (X[:, 0] - 600) ** 2 / 10 + rng.normal(0, 0.01, n)
    return X, y


@pytest.mark.parametrize("seed", [1, 2, 3])
def test_objective_matches_the_restatement_of_gpr_hpp(pkg, oracle_mod, seed):
    capi = importlib.import_module(PKG + ".capi")
    X, y = _sample(seed)
    for sigma, l, noise in ((10.0, 10.0, 1e-10), (3.0, 25.0, 1e-6), (0.7, 4.0, 1e-4)):
        want = oracle_mod.gpr_hyper_loss(X, y, sigma, l, noise)
        assert want is not None
        cost, g_coded = capi.gpr_nlml(X, y, sigma, l, noise, flavour=1)
        # the kernel matrix is conditioned like 1e12 at sigma_noise = 1e-10: two evaluations agree to cond * eps
        tol = 1e-4 if noise < 1e-8 else 1e-9
        assert np.isclose(cost, want[0], rtol=tol, atol=tol)
        assert np.allclose(g_coded, want[1], rtol=1e-3 if noise < 1e-8 else 1e-7, atol=1e-6 * np.abs(want[1]).max())


@pytest.mark.parametrize("seed", [1, 4])
def test_analytic_gradient_is_the_derivative_of_the_objective(pkg, seed):
    capi = importlib.import_module(PKG + ".capi")
    X, y = _sample(seed, n=16)
    noise = 1e-4   # well conditioned, so that finite differences mean something
    for sigma, l in ((2.0, 8.0), (0.5, 15.0)):
        _, g = capi.gpr_nlml(X, y, sigma, l, noise, flavour=0)
        h = 1e-6
        fd = np.array([
            (capi.gpr_nlml(X, y, sigma + h, l, noise)[0] - capi.gpr_nlml(X, y, sigma - h, l, noise)[0]) / (2 * h),
            (capi.gpr_nlml(X, y, sigma, l + h, noise)[0] - capi.gpr_nlml(X, y, sigma, l - h, noise)[0]) / (2 * h)])
        assert np.allclose(g, fd, rtol=1e-5, atol=1e-6 * np.abs(fd).max()), (g, fd)
        # ... and the expressions as coded in GPR.hpp:166-171 are NOT (matrix product for dK/dl, full Kff for dK/dsigma)
        _, gc = capi.gpr_nlml(X, y, sigma, l, noise, flavour=1)
        assert not np.allclose(gc, fd, rtol=1e-2)


def test_fit_minimises_the_objective(pkg):
    capi = importlib.import_module(PKG + ".capi")
    X, y = _sample(7, n=20)
    noise = 1e-4
    r = capi.gpr_fit(X, y, 10.0, 10.0, noise, max_iter=15)
    assert r["cost"] < r["cost0"] and 1 <= r["iterations"] <= 15
    assert np.isclose(r["cost0"], capi.gpr_nlml(X, y, 10.0, 10.0, noise)[0], rtol=1e-12)
    assert np.isclose(r["cost"], capi.gpr_nlml(X, y, r["sigma"], r["l"], noise)[0], rtol=1e-12)
    long = capi.gpr_fit(X, y, 10.0, 10.0, noise, max_iter=400)          # given room, it reaches a stationary point
    _, g = capi.gpr_nlml(X, y, long["sigma"], long["l"], noise)
    assert long["cost"] <= r["cost"] + 1e-9 and np.abs(g).max() < 1e-3 * max(1.0, abs(long["cost"]))
    zero = capi.gpr_fit(X, y, 10.0, 10.0, noise, max_iter=0)            # optimize = false: hyper-parameters unchanged
    assert (zero["sigma"], zero["l"], zero["iterations"]) == (10.0, 10.0, 0)
    with pytest.raises(pkg._abi.StlError):
        capi.gpr_nlml(X, y, 10.0, 10.0, -1e3)                             # not positive definite -> Evaluate returns false


@pytest.mark.gpu
@pytest.mark.skipif(not has_cuda(), reason="no CUDA device")
def test_per_factor_fit_inside_the_association(pkg, oracle_mod, small_pack, small_candidates):
    """params.gpr_optimize: every GPR factor gets its own (sigma, l) at association time (IBA_GPRFactor's constructor,
    IBACalib2.hpp:441-461).  The training data formed on the device equals the oracle's, the stored pair equals the
    host fit on that data, and the linearisation with per-factor hyper-parameters matches the oracle given the same pairs."""
    capi = importlib.import_module(PKG + ".capi")
    p = pkg.default_params(); p.use_gpr = 1; p.gpr_optimize = 1; p.gpr_sigma_noise = 1e-6
    pack = small_pack[0].shard(0, 3)
    x0 = small_candidates[0]
    orc = oracle_mod.Oracle(pack, params=p, kind="best")
    nb_o, _ = orc.associate(x0)
    with capi.Context(params=p) as c:
        c.upload(pack)
        nb = c.associate(x0)
        assert np.array_equal(nb, nb_o) and nb[3] > 20
        hyp = c.gpr_hyper()
        assert hyp.shape == (nb[3], 2) and not np.allclose(hyp, 10.0)     # the fit moved them
        for g in (0, int(nb[3]) // 2, int(nb[3]) - 1):
            X, y = orc.gpr_train(g, x0)
            r = capi.gpr_fit(X, y, p.gpr_sigma, p.gpr_l, p.gpr_sigma_noise, 15, 0)
            assert np.allclose(hyp[g], [r["sigma"], r["l"]], rtol=1e-6), (g, hyp[g], r)
            assert r["cost"] <= r["cost0"]
        orc.set_gpr_hyper(hyp)
        want = orc.linearize(small_candidates[:2])
        got = c.linearize(small_candidates[:2])
        assert np.array_equal(got[:, 57:], want[:, 57:])
        sh = np.abs(want[:, 8:57]).max(axis=1, keepdims=True)
        assert np.allclose(got[:, 0], want[:, 0], rtol=1e-6) and np.allclose(got[:, 8:57], want[:, 8:57], rtol=1e-6, atol=1e-6 * sh)
