timeout 300 python tools/tri_once.py
timeout 900 ncu --metrics gpu__time_duration.sum,smsp__inst_executed.sum --clock-control none -k regex:tri_ -c 14 --csv --log-file gpurun_out/r2_tri_launches.csv python tools/tri_once.py > gpurun_out/r2_tri.log 2>&1
