cd /root/repo
rm -f gpurun_out/variants10.jsonl gpurun_out/variants10.err
python tools/raster_variants.py main >> gpurun_out/variants10.jsonl 2>> gpurun_out/variants10.err
for v in b2l b2l16 b2l4 b1l b2lc2; do GA_B200_LIB=/root/repo/gaussiananything_b200/libga_b200_$v.so python tools/raster_variants.py $v >> gpurun_out/variants10.jsonl 2>> gpurun_out/variants10.err; done
export GA_B200_LIB=/root/repo/gaussiananything_b200/libga_b200_b2l.so
STEPS=2 timeout 900 ncu --set full --clock-control none --import-source on -k regex:render_bwd_a -s 9 -c 1 -f -o gpurun_out/r02_bwd_a_lists python tools/raster_variants.py ncu > gpurun_out/ncu10.log 2>&1
python tools/sweep_gemm.py > gpurun_out/gemm_sweep10.txt 2>&1
