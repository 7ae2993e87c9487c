timeout 900 python -m pytest tests/test_imaging_gpu.py tests/test_pipelines_gpu.py tests/test_golden_gpu.py -m gpu -x -q 2>&1 | tail -3
timeout 300 python tools/stats_once.py 2>&1 | tail -4
python tools/r2/c4_trace.py 2>&1 | tail -2
