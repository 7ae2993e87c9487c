timeout 900 python -m pytest tests/test_fits_gpu.py tests/test_stack_gpu.py -m gpu -x -q 2>&1 | tail -4
timeout 600 python bench.py --steps 10 --warmup 3 --secondary none > gpurun_out/r2_fits_line.json 2> gpurun_out/r2_fits_err.txt; tail -3 gpurun_out/r2_fits_err.txt
python -c "
import json
d=json.loads([l for l in open('gpurun_out/r2_fits_line.json') if l.startswith('{')][-1])
print(d['e2e']); print(d['secondary'].get('e2e_bitpix16'))"
