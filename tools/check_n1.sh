cd /root/repo
timeout 900 python -m pytest tests/test_vae_decoder_gpu.py tests/test_vae_ops_gpu.py -x -q > gpurun_out/n1_tests.txt 2>&1
echo "pytest rc=$?" >> gpurun_out/n1_tests.txt
timeout 600 python - > gpurun_out/n1_leg.jsonl 2> gpurun_out/n1_leg.err <<'PY'
import json, sys, os
sys.path.insert(0, ".")
import torch, bench
dev = torch.device("cuda:0")
print(json.dumps(bench.run_vae_decoder_leg(dev)))
os.environ["GA_B200_VAE_GRAPH"] = "0"
print(json.dumps(bench.run_vae_decoder_leg(dev)))
PY
timeout 900 python bench.py --steps 20 --warmup 5 --no-dit > gpurun_out/bench_c5.json 2> gpurun_out/bench_c5.err
