cd /root/repo
python -m pytest tests/test_raster_gpu.py tests/test_standin_gpu.py tests/test_vae_decoder_gpu.py -q 2>&1 | tail -25 > gpurun_out/t6.log
rm -f gpurun_out/variants6.jsonl gpurun_out/variants6.err
python tools/raster_variants.py lists_a3_tpi4 >> gpurun_out/variants6.jsonl 2>> gpurun_out/variants6.err
LISTK=0 python tools/raster_variants.py nolists >> gpurun_out/variants6.jsonl 2>> gpurun_out/variants6.err
for v in a4 tpi2 a4tpi2; do GA_B200_LIB=/root/repo/gaussiananything_b200/libga_b200_$v.so python tools/raster_variants.py lists_$v >> gpurun_out/variants6.jsonl 2>> gpurun_out/variants6.err; done
export GA_B200_LIB=/root/repo/gaussiananything_b200/libga_b200_a4tpi2.so
STEPS=3 timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches6.csv python tools/raster_variants.py l > /dev/null 2>&1
