timeout 300 python tools/warp_once.py
timeout 900 ncu --set full --clock-control none --import-source on -k regex:gather_kernel -s 1 -c 1 -o gpurun_out/r2_warp -f python tools/warp_once.py > gpurun_out/r2_warp_ncu.log 2>&1
tail -2 gpurun_out/r2_warp_ncu.log
