python -m pytest tests/test_calib_init.py tests/test_cpp_shim.py tests/test_gpr_fit.py -m gpu -x -q 2>&1 | tail -15
