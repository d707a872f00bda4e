#!/bin/bash
# ncu --set full of the attention kernel alone (self-attention shape of C3), one B200.
mkdir -p gpurun_out
cat > /tmp/attn_prof.py <<'PY'
import sys, ctypes as C, torch
sys.path.insert(0, ".")
from gaussiananything_b200 import dit
L = dit._bind(); dev = torch.device("cuda:0")
st = C.c_void_p(torch.cuda.current_stream(dev).cuda_stream)
B, H, N = 2, 12, 2048
q = torch.randn(B*H, N, 64, device=dev).bfloat16(); k = torch.randn(B*H, N, 64, device=dev).bfloat16()
vt = torch.randn(B*H, 64, N, device=dev).bfloat16(); o = torch.zeros(B, N, H*64, device=dev, dtype=torch.bfloat16)
for _ in range(4):
    L.ga_attention_bf16(dit._p(q), dit._p(k), dit._p(vt), dit._p(o), B, H, N, N, N, N, 0.125, 20.0, st)
torch.cuda.synchronize()
PY
ncu --set full --clock-control none --import-source on -k regex:attn_fwd -s 2 -c 1 -o gpurun_out/prof_attn \
    python /tmp/attn_prof.py > gpurun_out/attn_ncu.log 2>&1
ls -la gpurun_out/prof_attn.ncu-rep
