#!/bin/bash
# ncu recipe for the DiT kernels (one B200).  Outputs in gpurun_out/.
mkdir -p gpurun_out
cat > /tmp/dit_prof.py <<'PY'
import sys, torch
sys.path.insert(0, ".")
from gaussiananything_b200 import dit
torch.manual_seed(0)
dev = torch.device("cuda:0")
m = dit.DiT_models["DiT-PixArt-PCD-CLAY-B"](input_size=32, num_classes=0, learn_sigma=False, in_channels=3,
                                            context_dim=1024, roll_out=True, pooling_ctx_dim=768)
m.randomize_zero_init_().to(dev)
B, N, M = 2, 2048, 1369
x = torch.randn(B, N, 3, device=dev); t = torch.rand(B, device=dev)
ctx = {"img_crossattn": torch.randn(B, M, 1024, device=dev), "img_vector": torch.randn(B, 1024, device=dev)}
m(x, t, ctx)            # builds the engine (graph capture happens here)
m._engine.use_graph = False
for _ in range(2):
    m.forward_with_cfg(x, t, ctx, 4.0)
torch.cuda.synchronize()
PY
ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/dit_launches.csv \
    python /tmp/dit_prof.py > gpurun_out/dit_ncu.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:"gemm_bf16|attn_fwd" -s 300 -c 8 \
    -o gpurun_out/prof_dit python /tmp/dit_prof.py > gpurun_out/dit_ncu2.log 2>&1
ls -la gpurun_out | tail -5
