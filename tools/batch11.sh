cd /root/repo
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke11.log 2>&1
python -m pytest tests/test_raster_gpu.py -q 2>&1 | tail -5 > gpurun_out/t11.log
rm -f gpurun_out/variants11.jsonl
python tools/raster_variants.py main >> gpurun_out/variants11.jsonl 2>> gpurun_out/variants11.err
python tools/raster_variants.py main_again >> gpurun_out/variants11.jsonl 2>> gpurun_out/variants11.err
