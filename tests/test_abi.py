"""The C-ABI library loads and exports every symbol include/stlcalib.h declares (no GPU needed)."""
import ctypes as C
import os
import re

import pytest

from conftest import ROOT, has_cuda


def _declared(header):
    txt = open(os.path.join(ROOT, "include", header)).read()
    txt = re.sub(r"/\*.*?\*/", "", txt, flags=re.S)
    return sorted(set(re.findall(r"\b(stl_[a-z0-9_]+)\s*\(", txt)))


def test_calib_exports_every_declared_symbol(pkg):
    names = _declared("stlcalib.h")
    assert len(names) >= 18
    lib = pkg._abi.load_calib()  # raises if libstlcalib.so is not built
    for n in names:
        assert hasattr(lib, n), f"{n} declared in stlcalib.h but not exported"
    assert set(names) == set(pkg._abi.CALIB_SYMBOLS), "python symbol table out of sync with the header"
    assert lib.stl_abi_version() == 2


def test_synth_exports_every_declared_symbol(pkg):
    lib = pkg._abi.load_synth()
    for n in _declared("stlsynth.h"):
        assert hasattr(lib, n)


def test_default_params_match_yaml_values(pkg):
    """stl_default_params == KITTI-00 YAML (config/calib/00/iba_calib_global.yml:21-48) and the
    ctypes layout matches the C struct (a field mismatch would scramble these)."""
    lib = pkg._abi.load_calib()
    p = pkg._abi.Params()
    lib.stl_default_params(C.byref(p))
    q = pkg.default_params()
    for name, _ in pkg._abi.Params._fields_:
        a, b = getattr(p, name), getattr(q, name)
        if name == "err_weight":
            assert list(a) == list(b) == [1.0, 1.0]
        else:
            assert a == b, name
    assert (p.max_pixel_dist, p.norm_max_pts, p.norm_min_pts, p.norm_radius) == (1.5, 30, 5, 0.6)
    assert (p.corr_3d_2d_threshold, p.corr_3d_3d_threshold, p.norm_reg_threshold, p.min_diff_dist) == (40.0, 10.0, 0.02, 0.2)
    assert (p.he_threshold, p.valid_rate, p.num_min_corr, p.use_plane) == (0.094, 0.95, 30, 1)


def test_finalize_and_bbo_epilogue(pkg):
    """BAError's epilogue (iba_global.cpp:330-343) and eval_x's BBO (iba_global.cpp:386-388), host-only."""
    lib = pkg._abi.load_calib()
    p = pkg.default_params()
    s = pkg._abi.EvalSums(100.0, 20.0, 0.5, 10.0, 60.0, 50.0, 44.0, 40.0, 30.0, 10.0, 11.0, 500.0)
    o = pkg._abi.BAErrorOut()
    lib.stl_finalize(C.byref(p), C.byref(s), C.byref(o))
    assert (o.f1, o.f2, o.C, o.valid_cnt_3d_2d, o.cnt_3d_2d) == (2.0, 0.5, 0.05, 50, 60)
    bbo = (C.c_double * 4)()
    lib.stl_bbo(C.byref(p), C.byref(o), bbo)
    assert bbo[0] == 2.5 and bbo[1] == 0.05 - 0.094 and bbo[2] == -0.05 - 0.094
    assert bbo[3] == 0.95 - 50.0 / 61.0  # the "+1" of iba_global.cpp:388
    # no valid edge -> DBL_MAX sentinel
    z = pkg._abi.EvalSums(*([0.0] * 12))
    lib.stl_finalize(C.byref(p), C.byref(z), C.byref(o))
    import sys
    assert o.f1 == sys.float_info.max and o.f2 == sys.float_info.max


@pytest.mark.skipif(has_cuda(), reason="only meaningful on a box without a GPU")
def test_create_fails_loudly_without_gpu(pkg):
    """No CPU fallback: without an sm_100 device the product refuses to start."""
    import importlib
    capi = importlib.import_module(pkg.__name__ + ".capi")
    with pytest.raises(pkg._abi.StlError) as ei:
        capi.Context()
    assert ei.value.code == 3  # STL_ERR_NO_DEVICE


def test_product_never_touches_the_oracle():
    """The product path must not import, link or call anything under oracle/."""
    pdir = os.path.join(ROOT, "spatial-temporal-lidar-camera-calibration_b200")
    for dp, _, files in os.walk(pdir):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h", ".hpp", ".cpp")) or f == "Makefile":
                txt = open(os.path.join(dp, f), errors="ignore").read()
                assert "oracle/" not in txt and "liboracle" not in txt and "import oracle" not in txt, f
    import subprocess
    out = subprocess.run(["ldd", os.path.join(pdir, "libstlcalib.so")], capture_output=True, text=True).stdout
    assert "oracle" not in out


def test_k1_tables_arrive_by_bulk_copy(pkg):
    """The association kernel's per-keyframe tables are brought into shared memory by ONE cp.async.bulk with mbarrier completion
    and its covisible pixels are asked into L2 by cp.async.bulk.prefetch.L2: the sm_100a SASS of the built library must hold the
    bulk-copy unit's instructions (UBLKCP / UBLKPF) and the transaction-count barrier (SYNCS).  No GPU needed: cuobjdump reads
    the embedded cubin."""
    import shutil
    import subprocess
    if shutil.which("cuobjdump") is None:
        pytest.skip("cuobjdump not on PATH")
    so = os.path.join(ROOT, "spatial-temporal-lidar-camera-calibration_b200", "libstlcalib.so")
    pkg._abi.load_calib()
    sass = subprocess.run(["cuobjdump", "-sass", "-arch", "sm_100a", so], capture_output=True, text=True).stdout
    k1 = sass[sass.find("k_assoc2d"):]
    k1 = k1[:k1.find("Function :", 10)] if k1.find("Function :", 10) > 0 else k1
    assert "UBLKCP.S.G" in k1, "K1 lost its bulk copy global -> shared"
    assert "UBLKPF.L2" in k1, "K1 lost its bulk L2 prefetch"
    assert "SYNCS.ARRIVE.TRANS64" in k1 and "SYNCS.PHASECHK.TRANS64.TRYWAIT" in k1, "K1 lost its mbarrier"
