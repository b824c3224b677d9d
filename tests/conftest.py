import importlib
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

PKG = "spatial-temporal-lidar-camera-calibration_b200"


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


@pytest.fixture(scope="session")
def pkg():
    return importlib.import_module(PKG)


@pytest.fixture(scope="session")
def synth():
    return importlib.import_module(PKG + ".synth")


@pytest.fixture(scope="session")
def oracle_mod():
    from oracle import oracle as O
    O.load("port")  # builds liboracle.so if missing
    return O


@pytest.fixture(scope="session")
def small_pack(synth):
    """6 keyframes, full-size scans (~117k points each), 2000 keypoints."""
    pack, x_gt, twl = synth.generate(n_kf=6, seed=1000)
    return pack, x_gt


@pytest.fixture(scope="session")
def small_candidates(synth, small_pack):
    _, x_gt = small_pack
    return synth.candidates(x_gt, 4, spread=0.5)


def has_cuda():
    try:
        import torch
        return torch.cuda.is_available()
    except Exception:
        return False


@pytest.fixture(scope="session", params=[1, 0], ids=["plane-index", "plane-fit"])
def gpu_ctx(request, pkg, small_pack):
    """One context per flavour of the plane fit: looked up in the index built at upload (the default), or fitted
    per query at evaluation time (params.plane_index = 0, the reference's order of work)."""
    if not has_cuda():
        pytest.skip("no CUDA device")
    capi = importlib.import_module(PKG + ".capi")
    p = pkg.default_params()
    p.plane_index = request.param
    ctx = capi.Context(params=p)
    ctx.upload(small_pack[0])
    yield ctx
    ctx.close()
