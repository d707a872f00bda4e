cd /root/repo
python -m pytest tests/test_raster_gpu.py tests/test_standin_gpu.py -q -x 2>&1 | tail -15 > gpurun_out/t3.log
rm -f gpurun_out/variants3.jsonl
for g in 32 16 8; do GA_B200_FWD_GROUP=$g python tools/raster_variants.py fwd$g >> gpurun_out/variants3.jsonl 2>> gpurun_out/variants3.err; done
STEPS=2 timeout 600 ncu --set full --clock-control none --import-source on -k regex:render_ -s 8 -c 2 -f -o gpurun_out/r02_render python tools/raster_variants.py ncu > gpurun_out/ncu3.log 2>&1
python bench.py --steps 20 > gpurun_out/bench3.json 2> gpurun_out/bench3.err
