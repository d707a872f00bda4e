cd /root/repo
python tools/diag_c2_bwd.py > gpurun_out/diag_c2.log 2>&1
for cfg in "32:" "16:" "8:"; do g=${cfg%%:*}; GA_B200_FWD_GROUP=$g python tools/raster_variants.py fwd$g >> gpurun_out/variants.jsonl 2>> gpurun_out/variants.err; done
GA_B200_LIB=/root/repo/gaussiananything_b200/libga_b200_bwd3.so python tools/raster_variants.py bwd3 >> gpurun_out/variants.jsonl 2>> gpurun_out/variants.err
GA_B200_LIB=/root/repo/gaussiananything_b200/libga_b200_bwd3.so GA_B200_FWD_GROUP=16 python tools/raster_variants.py bwd3_fwd16 >> gpurun_out/variants.jsonl 2>> gpurun_out/variants.err
