cd /root/repo
rm -f gpurun_out/variants15.jsonl
timeout 120 python tools/raster_variants.py main128 >> gpurun_out/variants15.jsonl 2>> gpurun_out/variants15.err
for v in t4 u1 c5; do GA_B200_LIB=/root/repo/gaussiananything_b200/libga_b200_$v.so timeout 120 python tools/raster_variants.py $v >> gpurun_out/variants15.jsonl 2>> gpurun_out/variants15.err; done
