cd /root/repo
timeout 300 python -m pytest tests/test_raster_gpu.py -q 2>&1 | tail -5 > gpurun_out/t13.log
rm -f gpurun_out/variants13.jsonl
timeout 120 python tools/raster_variants.py tma >> gpurun_out/variants13.jsonl 2>> gpurun_out/variants13.err
GA_B200_LIB=/root/repo/gaussiananything_b200/libga_b200_notma.so timeout 120 python tools/raster_variants.py notma >> gpurun_out/variants13.jsonl 2>> gpurun_out/variants13.err
timeout 120 python tools/raster_variants.py tma2 >> gpurun_out/variants13.jsonl 2>> gpurun_out/variants13.err
