cd /root/repo
python -m pytest tests -m gpu -q 2>&1 | tail -15 > gpurun_out/t9.log
python bench.py --steps 50 > gpurun_out/bench9.json 2> gpurun_out/bench9.err
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench9_ref.json 2>> gpurun_out/bench9.err
STEPS=3 timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r02_launches_raster.csv python tools/raster_variants.py l > /dev/null 2>&1
STEPS=2 timeout 900 ncu --set full --clock-control none --import-source on -k regex:render_ -s 16 -c 5 -f -o gpurun_out/r02_render_final python tools/raster_variants.py ncu > gpurun_out/ncu9.log 2>&1
