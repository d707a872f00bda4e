cd /root/repo
timeout 300 python -m pytest tests/test_raster_gpu.py -q 2>&1 | tail -4 > gpurun_out/t14.log
rm -f gpurun_out/variants14.jsonl
timeout 120 python tools/raster_variants.py b256 >> gpurun_out/variants14.jsonl 2>> gpurun_out/variants14.err
for v in b128 b64; do GA_B200_LIB=/root/repo/gaussiananything_b200/libga_b200_$v.so timeout 120 python tools/raster_variants.py $v >> gpurun_out/variants14.jsonl 2>> gpurun_out/variants14.err; done
