timeout 300 python tools/tiles_once.py
timeout 900 ncu --set full --clock-control none --import-source on -k regex:bg_tiles_onchip -c 1 -o gpurun_out/r2_tiles -f python tools/tiles_once.py > gpurun_out/r2_tiles_ncu.log 2>&1
tail -3 gpurun_out/r2_tiles_ncu.log
