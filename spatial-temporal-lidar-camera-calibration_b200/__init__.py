"""B200-native calibration cost-evaluation path.

The directory name is not a Python identifier; import it with
``importlib.import_module("spatial-temporal-lidar-camera-calibration_b200")`` or
through the root-level alias module ``stlcalib_b200``.
"""
from . import _abi  # noqa: F401
from .pack import KeyFramePack, default_params  # noqa: F401
