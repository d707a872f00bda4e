cd /root/repo
python tools/debug_standin.py > gpurun_out/debug_standin.log 2>&1
export GA_B200_LIB=/root/repo/gaussiananything_b200/libga_b200_a4tpi2.so
STEPS=3 timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches5.csv python tools/raster_variants.py l > /dev/null 2>&1
STEPS=2 timeout 600 ncu --set full --clock-control none --import-source on -k regex:render_bwd_[ab] -s 8 -c 2 -f -o gpurun_out/r02_bwd_split python tools/raster_variants.py ncu > gpurun_out/ncu5.log 2>&1
