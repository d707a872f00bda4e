cd /root/repo
python -m pytest tests/test_raster_gpu.py tests/test_standin_gpu.py tests/test_vae_decoder_gpu.py -q 2>&1 | tail -25 > gpurun_out/t4.log
rm -f gpurun_out/variants4.jsonl gpurun_out/variants4.err
python tools/raster_variants.py split_a3_tpi4 >> gpurun_out/variants4.jsonl 2>> gpurun_out/variants4.err
GA_B200_BWD_SPLIT=0 python tools/raster_variants.py fused >> gpurun_out/variants4.jsonl 2>> gpurun_out/variants4.err
for v in a4 tpi2 a4tpi2 a4tpi8; do GA_B200_LIB=/root/repo/gaussiananything_b200/libga_b200_$v.so python tools/raster_variants.py $v >> gpurun_out/variants4.jsonl 2>> gpurun_out/variants4.err; done
