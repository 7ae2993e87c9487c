timeout 900 python -m pytest tests/test_stars_gpu.py tests/test_golden_gpu.py tests/test_preview_gpu.py tests/test_pipelines_gpu.py -m gpu -x -q 2>&1 | tail -5
AB_TRACE=1 timeout 600 python bench.py --steps 5 --warmup 3 --secondary c3 2>&1 | tail -30
