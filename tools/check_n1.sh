cd /root/repo
timeout 900 python -m pytest tests/test_vae_decoder_gpu.py tests/test_vae_ops_gpu.py -x -q > gpurun_out/n1_tests.txt 2>&1
echo "pytest rc=$?" >> gpurun_out/n1_tests.txt
: > gpurun_out/n1_leg.jsonl
for v in "" _rows1 _m16 _rows4; do
GA_B200_LIB=/root/repo/gaussiananything_b200/libga_b200$v.so timeout 300 python - >> gpurun_out/n1_leg.jsonl 2>> gpurun_out/n1_leg.err <<'PY'
import json, sys, os
sys.path.insert(0, ".")
import torch, bench
dev = torch.device("cuda:0")
d = bench.run_vae_decoder_leg(dev, reps=10)
d["lib"] = os.environ.get("GA_B200_LIB")
print(json.dumps(d))
PY
done
