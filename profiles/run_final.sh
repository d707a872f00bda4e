#!/bin/bash
# Round-end validation + profiling batch (one B200, under gpurun).  Outputs land in gpurun_out/.
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/gpu.csv
timeout 600 python -m pytest tests -m gpu -q 2>&1 | tail -3 > gpurun_out/pytest_gpu.txt; cat gpurun_out/pytest_gpu.txt
timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -2
timeout 600 python bench.py > gpurun_out/bench_final.json 2> gpurun_out/bench_final.err; tail -c 600 gpurun_out/bench_final.json; tail -3 gpurun_out/bench_final.err
timeout 300 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_reference.json 2>/dev/null; cat gpurun_out/bench_reference.json
# launch lists (cold-cache, serialised: compare shares)
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 300 --csv --log-file gpurun_out/launches_raster.csv \
    python bench.py --steps 2 --warmup 3 --no-dit > /dev/null 2>&1
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -s 2000 -c 600 --csv --log-file gpurun_out/launches_dit.csv \
    python tools/dit_leg.py > /dev/null 2>&1
# full captures: the two render kernels, one GEMM of each hot epilogue, the attention kernel
timeout 300 ncu --set full --clock-control none --import-source on -k regex:render_ -s 6 -c 2 -o gpurun_out/final_render \
    python bench.py --steps 2 --warmup 3 --no-dit > /dev/null 2>&1
timeout 200 ncu --set full --clock-control none --import-source on -k regex:gemm -s 3 -c 1 -o gpurun_out/final_gemm_mlp1 \
    python tools/one_gemm.py 4096 3072 768 128 gelu > /dev/null 2>&1
bash profiles/run_profile_attn.sh > /dev/null 2>&1; mv gpurun_out/prof_attn.ncu-rep gpurun_out/final_attn.ncu-rep
# row N1 (VAE decoder): launch list of one decode at the deployed size + the micro-attention kernel in full
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
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 2000 --csv --log-file gpurun_out/launches_n1.csv \
    python /tmp/n1_prof.py > /dev/null 2>&1
timeout 200 ncu --set full --clock-control none --import-source on -k regex:micro_attention -s 3 -c 1 -o gpurun_out/final_micro_attn \
    python /tmp/n1_prof.py > /dev/null 2>&1
ls -la gpurun_out | tail -15
