cd /root/repo
timeout 1500 python -m pytest tests -m gpu -q 2>&1 | tail -4 > gpurun_out/r02_pytest_gpu.txt
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r02_smoke.txt 2>&1
timeout 900 python bench.py --steps 100 > gpurun_out/r02_bench.json 2> gpurun_out/r02_bench.err
timeout 300 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r02_bench_reference.json 2>/dev/null
STEPS=3 timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r02_launches_raster.csv python tools/raster_variants.py l > /dev/null 2>&1
STEPS=2 timeout 900 ncu --set full --clock-control none --import-source on -k regex:render_ -s 24 -c 6 -f -o gpurun_out/r02_render python tools/raster_variants.py ncu > /dev/null 2>&1
