timeout 900 python -m pytest tests/test_stack_gpu.py tests/test_calibration_gpu.py -m gpu -x -q 2>&1 | tail -5
timeout 300 python tools/r2/deep_time.py 2>&1 | tail -12
