cd /root/repo
python -m pytest tests/test_raster_gpu.py -q 2>&1 | tail -8 > gpurun_out/t8.log
rm -f gpurun_out/variants8.jsonl gpurun_out/variants8.err
python tools/raster_variants.py default_a3tpi4 >> gpurun_out/variants8.jsonl 2>> gpurun_out/variants8.err
for v in a4 a4tpi2; do GA_B200_LIB=/root/repo/gaussiananything_b200/libga_b200_$v.so python tools/raster_variants.py $v >> gpurun_out/variants8.jsonl 2>> gpurun_out/variants8.err; done
GA_B200_LIB=/root/repo/gaussiananything_b200/libga_b200_a4tpi2.so LISTK=48 python tools/raster_variants.py a4tpi2_k48 >> gpurun_out/variants8.jsonl 2>> gpurun_out/variants8.err
export GA_B200_LIB=/root/repo/gaussiananything_b200/libga_b200_a4tpi2.so
STEPS=3 timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches8.csv python tools/raster_variants.py l > /dev/null 2>&1
