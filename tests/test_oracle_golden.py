"""The self-contained oracle (in-repo KD-tree port) reproduces the committed golden vectors,
which were generated through the reference's real nanoflann (tests/golden/make_golden.py)."""
import glob
import os

import numpy as np
import pytest

GOLDEN = sorted(glob.glob(os.path.join(os.path.dirname(__file__), "golden", "*.npz")))


def _params(pkg, g):
    """stl_params_t of a fixture: the defaults plus the overrides stored as p_<field>."""
    p = pkg.default_params()
    for key in g.files:
        if key.startswith("p_"):
            cur = getattr(p, key[2:])
            setattr(p, key[2:], type(cur)(g[key]))
    return p


@pytest.mark.parametrize("path", GOLDEN, ids=[os.path.basename(p) for p in GOLDEN])
def test_port_reproduces_golden(path, pkg, oracle_mod):
    g = np.load(path)
    pack = pkg.KeyFramePack.from_npz_dict(g)
    params = _params(pkg, g)
    orc = oracle_mod.Oracle(pack, params=params, kind="port")
    sums, ties, cnt = orc.ba_error_sums(g["X"], mode=0)
    assert ties.sum() == 0
    assert np.array_equal(sums, g["sums"]) and np.array_equal(cnt, g["counters"])
    # OpenMP-over-keyframes mode (iba_func) only re-associates the fp64 sums
    s1, _, _ = orc.ba_error_sums(g["X"], mode=1)
    assert np.allclose(s1, g["sums"], rtol=1e-13, atol=0)
    for b in range(2):
        for kf in range(pack.n_kf):
            d = orc.frame_debug(g["X"][b], kf)
            for key in ("corr_kp", "corr_pt", "align_nn", "align_m", "align_is_plane", "align_knn", "align_dist"):
                assert np.array_equal(d[key], g[f"b{b}_kf{kf}_{key}"]), (b, kf, key)
    if params.variant == 0:
        nb, _ = orc.associate(g["X"][0])
        assert np.array_equal(nb, g["lm_nblocks"]) and np.array_equal(orc.block_keys(), g["lm_keys"])
        assert np.array_equal(orc.linearize(g["X"]), g["lm_lin"])
        assert (nb[3] > 0) == bool(params.use_gpr)


def test_golden_covers_the_edge_cases():
    g = np.load([p for p in GOLDEN if "sparse_ragged" in p][0])
    n = np.diff(g["scan_offset"])
    assert (n == 0).any(), "an empty scan"
    assert g["sums"][0][10] < g["n_kf"], "some keyframe is skipped by the num_min_corr gate"
    assert any(len(g[f"b0_kf{k}_align_nn"]) == 0 and len(g[f"b0_kf{k}_corr_kp"]) > 0 for k in range(int(g["n_kf"])))
