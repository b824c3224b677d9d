"""include/stlcalib_host.hpp (the C++ mirror of BAError / BALoss / BuildProblem) compiles against the
C-ABI, fails loudly without a GPU, and on the GPU box returns the same numbers as the ctypes path."""
import importlib
import os
import subprocess

import numpy as np
import pytest

from conftest import PKG, ROOT, has_cuda

PKGDIR = os.path.join(ROOT, PKG)
EXE = os.path.join(ROOT, "tests", "cpp", "host_shim_test")


def _build():
    src = os.path.join(ROOT, "tests", "cpp", "host_shim_test.cpp")
    if os.path.exists(EXE) and os.path.getmtime(EXE) > max(os.path.getmtime(src), os.path.getmtime(os.path.join(ROOT, "include", "stlcalib_host.hpp"))):
        return
    cmd = ["/usr/bin/g++", "-std=c++17", "-O2", "-I" + os.path.join(ROOT, "include"), src, "-o", EXE,
           "-L" + PKGDIR, "-l:libstlcalib.so", "-l:libstlsynth.so", "-Wl,-rpath," + PKGDIR,
           "-L/usr/local/cuda/lib64", "-Wl,-rpath,/usr/local/cuda/lib64"]
    r = subprocess.run(cmd, capture_output=True, text=True)
    assert r.returncode == 0, r.stderr


def test_cpp_shim_compiles_and_links():
    _build()
    assert os.path.exists(EXE)


@pytest.mark.skipif(has_cuda(), reason="only meaningful without a GPU")
def test_cpp_shim_fails_loudly_without_gpu():
    _build()
    r = subprocess.run([EXE, "1"], capture_output=True, text=True)
    assert r.returncode == 3 and "sm_100" in r.stderr


@pytest.mark.gpu
def test_cpp_shim_matches_ctypes_path(pkg, synth):
    _build()
    r = subprocess.run([EXE, "3"], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    capi = importlib.import_module(PKG + ".capi")
    host = importlib.import_module(PKG + ".host")
    pack, x_gt, _ = synth.generate(n_kf=3, beams=32, az_steps=900, n_kp=500, seed=21)
    X = synth.candidates(x_gt, 3, 0.4)
    with capi.Context() as ctx:
        ctx.upload(pack)
        lines = [ln for ln in r.stdout.splitlines() if ln.startswith("BA ")]
        assert len(lines) == 3
        for b, ln in enumerate(lines):
            left, right = ln.split("|")
            f = left.split()
            ba = host.BAError(X[b], ctx)
            got = np.array([float(f[2]), float(f[3]), float(f[4]), int(f[5]), int(f[6])])
            assert np.array_equal(got, np.array(ba, dtype=np.float64), equal_nan=True)  # candidate 1 has no valid edge: DBL_MAX / NaN sentinels
            bb = right.split()
            assert bb[1:3] == ["1", "1"]
            assert np.array_equal(np.array([float(v) for v in bb[3:7]]), np.array(ctx.bbo(ba)), equal_nan=True)
        nb = ctx.associate(X[0])
        L = ctx.linearize(X[1])[0]
        lm = [ln for ln in r.stdout.splitlines() if ln.startswith("LM ")][0].split()
        assert [int(v) for v in lm[1:4]] == nb.tolist()[:3]
        assert (float(lm[4]), float(lm[5]), float(lm[6])) == (L[0], L[1], L[8])
        Bk = ctx.eval_blocks(X[1], rmax=6)
        blk = [ln for ln in r.stdout.splitlines() if ln.startswith("BLK ")][0].split()
        assert int(blk[1]) == len(Bk["type"]) == nb.sum() and [int(v) for v in blk[2:5]] == [Bk["type"][0], Bk["n_res"][0], Bk["kp"][0]]
        assert float(blk[5]) == Bk["residuals"][0, 0] and np.isclose(float(blk[6]), (Bk["residuals"] ** 2).sum(), rtol=1e-12)


# ---- N1: the optimiser adapters (include/adapters/*.hpp) compiled against stand-in NOMAD / Ceres / g2o headers
ADAPT_EXE = os.path.join(ROOT, "tests", "cpp", "adapters_test")


def _build_adapters():
    src = os.path.join(ROOT, "tests", "cpp", "adapters_test.cpp")
    deps = [src] + [os.path.join(ROOT, "include", "adapters", f) for f in ("stl_nomad.hpp", "stl_ceres.hpp", "stl_g2o.hpp")] + \
           [os.path.join(ROOT, "include", "stlcalib_host.hpp")]
    if os.path.exists(ADAPT_EXE) and os.path.getmtime(ADAPT_EXE) > max(os.path.getmtime(d) for d in deps):
        return
    cmd = ["/usr/bin/g++", "-std=c++17", "-O2", "-Wall", "-I" + os.path.join(ROOT, "include"), "-I" + os.path.join(ROOT, "tests", "cpp", "stubs"),
           src, "-o", ADAPT_EXE, "-L" + PKGDIR, "-l:libstlcalib.so", "-l:libstlsynth.so", "-Wl,-rpath," + PKGDIR,
           "-L/usr/local/cuda/lib64", "-Wl,-rpath,/usr/local/cuda/lib64"]
    r = subprocess.run(cmd, capture_output=True, text=True)
    assert r.returncode == 0, r.stderr


def test_adapters_compile_against_optimiser_interfaces():
    """NomadBALoss : NOMAD::Evaluator, StlEvaluationCallback / StlBlockCost : ceres::*, StlPlaneEdge : g2o::BaseUnaryEdge<20,...>
    override the virtuals of the (stand-in) third-party interfaces: a signature mismatch is a compile error."""
    _build_adapters()
    assert os.path.exists(ADAPT_EXE)


@pytest.mark.skipif(has_cuda(), reason="only meaningful without a GPU")
def test_adapters_fail_loudly_without_gpu():
    _build_adapters()
    r = subprocess.run([ADAPT_EXE, "1"], capture_output=True, text=True)
    assert r.returncode == 3 and "sm_100" in r.stderr


@pytest.mark.gpu
def test_adapters_match_ctypes_path(pkg, synth):
    _build_adapters()
    r = subprocess.run([ADAPT_EXE, "3"], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    capi = importlib.import_module(PKG + ".capi")
    host = importlib.import_module(PKG + ".host")
    pack, x_gt, _ = synth.generate(n_kf=3, beams=32, az_steps=900, n_kp=500, seed=21)
    X = synth.candidates(x_gt, 3, 0.4)
    out = {ln.split()[0]: [] for ln in r.stdout.splitlines()}
    for ln in r.stdout.splitlines():
        out[ln.split()[0]].append(ln.split()[1:])
    with capi.Context() as ctx:
        ctx.upload(pack)
        loss = host.BALoss(ctx)
        bbo = [np.array(b) for b in loss.eval_block(X)]
        nx = out["NOMAD_X"][0]
        assert nx[:2] == ["1", "1"] and np.array_equal(np.array([float(v) for v in nx[2:6]]), bbo[1], equal_nan=True)
        for b, row in enumerate(out["NOMAD_B"]):
            assert row[:3] == [str(b), "1", "1"] and np.array_equal(np.array([float(v) for v in row[3:7]]), bbo[b], equal_nan=True)
        nb = ctx.associate(X[0])
        L = ctx.linearize(X[1])[0]
        ce = out["CERES"][0]
        assert int(ce[0]) == nb.sum()
        assert np.isclose(float(ce[1]), L[0], rtol=1e-12) and np.isclose(float(ce[2]), L[1], rtol=1e-9) and np.isclose(float(ce[3]), L[8], rtol=1e-9)
        Bk = ctx.eval_blocks(X[1], rmax=20)
        g2 = out["G2O"][0]
        assert int(g2[0]) == nb.sum() and np.isclose(float(g2[1]), (Bk["residuals"] ** 2).sum(), rtol=1e-12)
        assert np.isclose(float(g2[2]), Bk["jacobians"][:, 0, 0].sum(), rtol=1e-9)
