cd /root/repo
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/r02_gpu.csv
timeout 1500 python -m pytest tests -m gpu -q 2>&1 | tail -4 > gpurun_out/r02_pytest_gpu.txt
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r02_smoke.txt 2>&1
timeout 900 python bench.py --steps 100 > gpurun_out/r02_bench.json 2> gpurun_out/r02_bench.err
timeout 300 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r02_bench_reference.json 2>/dev/null
# DiT launch list (GEMM tile widths changed) and the VAE decoder's list + micro-attention capture (eager launches: GA_B200_VAE_GRAPH=0)
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -s 2000 -c 600 --csv --log-file gpurun_out/r02_launches_dit.csv \
    python tools/dit_leg.py > /dev/null 2>&1
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
ls -la gpurun_out | grep r02_
