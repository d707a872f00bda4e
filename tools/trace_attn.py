"""Debug timeline of the attention kernel (variant chosen by GA_B200_ATTN, see dit_attention.cu): builds a -DGA_B200_TRACE copy of the library into gpurun_out/ (never the
product .so), runs the C3 self-attention shape and prints where softmax warp 2 and the MMA thread spend their cycles.
    python tools/trace_attn.py build     (no GPU needed: cross-compiles)
    python tools/trace_attn.py run       (on the GPU box)
"""
import ctypes as C
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CSRC = os.path.join(ROOT, "gaussiananything_b200", "csrc")
OUT = os.path.join(ROOT, "gpurun_out", "libga_b200_trace.so")


def build():
    os.makedirs(os.path.dirname(OUT), exist_ok=True)
    srcs = [os.path.join(CSRC, f) for f in ("dit_attention.cu", "dit_gemm.cu")]
    cmd = ["nvcc", "-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo", "-std=c++17", "-Xcompiler", "-fPIC",
           "--expt-relaxed-constexpr", "-DGA_B200_TRACE", "-shared", "-cudart", "shared", "-o", OUT] + srcs + ["-lcuda"]
    subprocess.check_call(cmd)
    print(OUT)


def run():
    import numpy as np
    import torch
    L = C.CDLL(OUT)
    vp, i32, f32 = C.c_void_p, C.c_int, C.c_float
    L.ga_attention_bf16.argtypes = [vp, vp, vp, vp, i32, i32, i32, i32, i32, i32, f32, f32, vp]
    dev = torch.device("cuda:0")
    st = C.c_void_p(torch.cuda.current_stream(dev).cuda_stream)
    B, H, N = 2, 12, 2048
    q = torch.randn(B * H, N, 64, device=dev).bfloat16()
    k = torch.randn(B * H, N, 64, device=dev).bfloat16()
    vt = torch.randn(B * H, 64, N, device=dev).bfloat16()
    o = torch.zeros(B, N, H * 64, device=dev, dtype=torch.bfloat16)
    for _ in range(3):
        rc = L.ga_attention_bf16(q.data_ptr(), k.data_ptr(), vt.data_ptr(), o.data_ptr(), B, H, N, N, N, N, 0.125, 20.0, st)
        assert rc == 0, rc
        torch.cuda.synchronize()
    torch.cuda.synchronize()
    nc = 384
    n = nc * 160
    buf = (C.c_ulonglong * n)()
    assert L.ga_debug_attn_trace(buf, n) == 0
    t = np.frombuffer(buf, dtype=np.uint64).reshape(nc, 160).astype(np.int64)
    smid, start, end = t[:, 0], t[:, 1], t[:, 2]
    life = end - start
    print("CTA lifetime cycles: min %d median %d max %d" % (life.min(), np.median(life), life.max()))
    ev = t[:, 8:8 + 16 * 8].reshape(nc, 16, 8)          # first 16 key blocks of every CTA
    first = ev[:, 0, 0] - start
    print("start -> first S ready: median %d" % np.median(first))
    names = ["wait S", "tmem ld", "exp h0", "p_empty wait + rest -> p_full"]
    blk = ev[:, 1:, :]                     # steady state blocks
    prev_done = ev[:, :-1, 3]
    print("softmax warp 2, per key block (median cycles over CTAs and blocks):")
    print("  wait for S after previous p_full arrive: %d" % np.median(blk[:, :, 0] - prev_done))
    print("  S ready -> tmem loaded:                  %d" % np.median(blk[:, :, 1] - blk[:, :, 0]))
    print("  tmem loaded -> first 32 exps done:       %d" % np.median(blk[:, :, 2] - blk[:, :, 1]))
    print("  -> p_full arrive:                        %d" % np.median(blk[:, :, 3] - blk[:, :, 2]))
    print("  whole block:                             %d" % np.median(blk[:, :, 3] - prev_done))
    m = ev[:, :15, :]
    print("MMA thread: s_empty seen -> S issued %d ; S issued -> p_full seen %d ; p_full -> PV issued %d" % (
        np.median(m[:, :, 5] - m[:, :, 4]), np.median(m[:, :, 6] - m[:, :, 5]), np.median(m[:, :, 7] - m[:, :, 6])))
    print("  s_empty(j) seen relative to softmax tmem-loaded(j): %d" % np.median(m[:, :, 4] - ev[:, :15, 1]))
    print("  S(j+1) ready (softmax saw) - S(j+1) issued: %d" % np.median(ev[:, 1:, 0][:, :15] - m[:, :, 5]))
    # per-SM view: how many CTAs per SM, and total span
    span = end.max() - start.min()
    print("kernel span %d cycles; CTAs per SM max %d" % (span, np.bincount(smid).max()))
    med = np.median(ev[:, 1:, 3] - ev[:, :-1, 3], axis=1)
    print("per-CTA median block duration: percentiles 10/50/90: %s" % np.percentile(med, [10, 50, 90]).tolist())
    print("CTAs per SM histogram:", np.bincount(np.bincount(smid)).tolist())
    slow = med > 3500
    print("slow CTAs: %d ; their SMs also host: %s" % (slow.sum(), [int((smid == sm).sum()) for sm in smid[slow][:12]]))
    print("slow CTA ids:", np.nonzero(slow)[0][:40].tolist())
    print("slow CTA sms:", smid[slow][:40].tolist())
    c = int(np.nonzero(slow)[0][0]) if slow.any() else 0
    e = ev[c]
    print("CTA %d breakdown per block [waitS, tmem, exp32, rest]:" % c)
    for b in range(1, 8):
        print("   ", [int(e[b, 0] - e[b - 1, 3]), int(e[b, 1] - e[b, 0]), int(e[b, 2] - e[b, 1]), int(e[b, 3] - e[b, 2])],
              " mma: s_empty->Sissued %d, p_full seen at +%d after softmax arrive, PV issued +%d" % (
                  e[b, 5] - e[b, 4], e[b, 6] - e[b, 3], e[b, 7] - e[b, 6]))
    for cta in (0, 200, nc - 1):
        e = ev[cta]
        print("CTA %d sm %d life %d: block durations %s" % (cta, smid[cta], life[cta], (e[1:, 3] - e[:-1, 3]).tolist()))


if __name__ == "__main__":
    build() if sys.argv[1:2] == ["build"] else run()
