AB_TRACE=1 timeout 300 python tools/detect_once.py 2>&1 | tail -6
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -s 60 -c 80 --csv --log-file gpurun_out/r2_detect_launches.csv python tools/detect_once.py > gpurun_out/r2_detect.log 2>&1
