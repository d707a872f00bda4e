"""Row N1 (VAE decoder -> surfels) at the deployed size, on CPU: builds the reference decoder with random weights
(tests/golden/_ref_stubs.py), runs the reference and the oracle once, checks they agree, and prints the FLOP count per
sample and the CPU time -- the numbers the future CUDA path and its bench leg will be measured against.
Build container only (needs /root/reference).   python tools/n1_cost.py"""
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tests", "golden"))
sys.path.insert(0, ROOT)
import make_vae_golden as mk  # noqa: E402  (installs the stubs)
import torch  # noqa: E402
from oracle import vae_decoder_oracle as vo  # noqa: E402

mk.D, mk.DEPTH, mk.HEADS, mk.TOK, mk.ZC = 768, 12, 12, 16, 10
torch.set_num_threads(max(1, os.cpu_count() // 2))
m, g = mk.build(seed=3)
B, N, D = 1, mk.D, mk.D
lat = torch.randn(B, N, mk.ZC, generator=g)
xyz = (torch.rand(B, N, 3, generator=g) - 0.5) * 0.8
ret = {"latent_normalized": lat, "query_pcd_xyz": xyz}
with torch.no_grad():
    t0 = time.perf_counter()
    latent = m.vit_decode_backbone(ret, 64)
    out = m.vit_decode_postprocess(latent, dict(ret))
    t_ref = time.perf_counter() - t0
    sd = {k: v for k, v in m.state_dict().items()}
    t0 = time.perf_counter()
    mine = vo.decode(sd, lat, xyz, mk.HEADS, mk.DEPTH, 0.45, float(m.skip_weight))
    t_or = time.perf_counter() - t0
for k in ("gaussians_base", "gaussians_upsampled", "gaussians_upsampled_2", "gaussians_upsampled_3"):
    e = float((mine[k] - out[k]).norm() / out[k].norm())
    print("%-24s %-16s rel err oracle vs reference %.2e" % (k, tuple(out[k].shape), e))


def block_flops(L, D, mlp=4):               # one pre-norm transformer block over L tokens of width D (multiply-adds x 2)
    return 2 * L * D * (3 * D + D + 2 * mlp * D) + 4 * L * L * D


dit2 = 12 * (block_flops(N, D) + 2 * N * D * 6 * D)                      # + the per-token adaLN GEMM
u1 = N * 2 * block_flops(9, D)                                            # f = 8, depth 2
u2 = N * 8 * 1 * block_flops(5, D)                                        # f = 4, depth 1
u3 = N * 32 * 1 * block_flops(4, D)                                       # f = 3, depth 1
tot = dit2 + u1 + u2 + u3
print("FLOP per sample: DiT2 %.2f G, up-samplers %.2f + %.2f + %.2f G, total %.2f TFLOP" % (
    dit2 / 1e9, u1 / 1e9, u2 / 1e9, u3 / 1e9, tot / 1e12))
print("CPU (%d threads): reference %.1f s, oracle %.1f s per sample -> %.3f samples/s" % (
    torch.get_num_threads(), t_ref, t_or, 1.0 / t_or))
