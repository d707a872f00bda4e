cd /root/repo
timeout 900 python -m pytest tests/test_vae_decoder_gpu.py tests/test_vae_ops_gpu.py -x -q > gpurun_out/n1_tests.txt 2>&1
echo "pytest rc=$?" >> gpurun_out/n1_tests.txt
: > gpurun_out/n1_leg.jsonl
for v in "" _small0; do
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
cat > /tmp/n1_prof.py <<'PY'
import sys, torch
sys.path.insert(0, ".")
from gaussiananything_b200.vae_decoder import SurfelDecoder, random_state_dict
dev = torch.device("cuda:0")
dec = SurfelDecoder(random_state_dict(768, 12, 10, seed=0), 12, 12, device=dev)
lat = torch.randn(2, 768, 10, device=dev); xyz = (torch.rand(2, 768, 3, device=dev) - 0.5) * 0.8
for _ in range(2):
    dec.decode(lat, xyz)
torch.cuda.synchronize()
PY
GA_B200_VAE_GRAPH=0 timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 2000 --csv --log-file gpurun_out/r02_launches_n1.csv \
    python /tmp/n1_prof.py > /dev/null 2>&1
GA_B200_VAE_GRAPH=0 timeout 200 ncu --set full --clock-control none --import-source on -k regex:micro_attention -s 3 -c 1 -f -o gpurun_out/r02_micro_attn \
    python /tmp/n1_prof.py > /dev/null 2>&1
