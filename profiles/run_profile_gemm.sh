#!/bin/bash
# ncu --set full of the tcgen05 GEMM (M=4096 N=3072 K=768, bias+bf16 epilogue): configs 128 and 9256
mkdir -p gpurun_out
cat > /tmp/gemm_prof.py <<'PY'
import sys, ctypes as C, math, torch
sys.path.insert(0, ".")
from gaussiananything_b200 import dit
L = dit._bind(); dev = torch.device("cuda:0")
st = C.c_void_p(torch.cuda.current_stream(dev).cuda_stream)
M, N, K = 4096, 3072, 768
A = torch.randn(M, K, device=dev).bfloat16(); W = (torch.randn(N, K, device=dev) / math.sqrt(K)).bfloat16()
bias = torch.randn(N, device=dev); out = torch.zeros(M, N, device=dev, dtype=torch.bfloat16)
e = dit.GaGemmEpilogue(mode=dit.EPI_BF16, bias=bias.data_ptr(), out=out.data_ptr(), ld_out=N)
for cfg in (128, 9256):
    for _ in range(3):
        L.ga_gemm_bf16_tn(dit._p(A), K, dit._p(W), K, M, N, K, C.byref(e), cfg, st)
torch.cuda.synchronize()
PY
ncu --set full --clock-control none --import-source on -k regex:gemm_bf16 -s 2 -c 1 -o gpurun_out/prof_gemm128 \
    python /tmp/gemm_prof.py > gpurun_out/gemm_ncu.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:pair_kernel -s 2 -c 1 -o gpurun_out/prof_gemm9256 \
    python /tmp/gemm_prof.py >> gpurun_out/gemm_ncu.log 2>&1
ls -la gpurun_out/prof_gemm*.ncu-rep
