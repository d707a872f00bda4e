"""Mirror of the deployed image-to-3D DiT denoisers of the reference:

  dit.dit_i23d.DiT_I23D_PCD_PixelArt_noclip               (/root/reference/dit/dit_i23d.py:437-567)
  dit.dit_i23d.DiT_I23D_PCD_PixelArt_noclip_clay_stage2   (/root/reference/dit/dit_i23d.py:664-750)
  block ImageCondDiTBlockPixelArtRMSNormClayLRM           (/root/reference/dit/dit_models_xformers.py:717-787)
  registry DiT_models                                     (/root/reference/dit/dit_i23d.py:1665-1697)

Same constructor arguments, attribute names and state_dict key layout (a
reference checkpoint loads with strict=True), same `forward(x, timesteps,
context)` / `forward_with_cfg(x, t, context, cfg_scale)` contracts.  The
forward pass does no arithmetic in torch: every op is a kernel of
libga_b200.so (tcgen05 GEMMs with fused epilogues, tcgen05 flash attention,
fused RMSNorm+modulate, ...) enqueued on the current CUDA stream, optionally
replayed from a CUDA graph.  There is no CPU / eager fallback.

Precision: GEMM / attention operands bf16 (the reference runs them under bf16
autocast), fp32 accumulation, fp32 residual stream, fp32 norms, fp32 output.
"""
import ctypes as C
import math

import torch
import torch.nn as nn

from . import _lib


def _p(t):
    return C.c_void_p(t.data_ptr() if t is not None else 0)


class GaGemmEpilogue(C.Structure):
    _fields_ = [("mode", C.c_int), ("bias", C.c_void_p), ("out", C.c_void_p), ("ld_out", C.c_int),
                ("gate", C.c_void_p), ("gate_ld", C.c_int), ("rows_per_batch", C.c_int),
                ("q", C.c_void_p), ("k", C.c_void_p), ("vt", C.c_void_p),
                ("qn_w", C.c_void_p), ("kn_w", C.c_void_p), ("heads", C.c_int), ("first_part", C.c_int),
                ("tok_pitch", C.c_int), ("eps", C.c_float)]


EPI_BF16, EPI_GELU_BF16, EPI_F32, EPI_RESID_GATE_F32, EPI_HEADS = 0, 1, 2, 3, 4
_bound = False


def _bind():
    global _bound
    L = _lib.lib()
    if _bound:
        return L
    vp, i32, i64, f32 = C.c_void_p, C.c_int, C.c_int64, C.c_float
    L.ga_gemm_bf16_tn.argtypes = [vp, i32, vp, i32, i32, i32, i32, C.POINTER(GaGemmEpilogue), i32, vp]
    L.ga_attention_bf16.argtypes = [vp, vp, vp, vp, i32, i32, i32, i32, i32, i32, f32, f32, vp]
    L.ga_rmsnorm_modulate.argtypes = [vp, vp, vp, vp, i32, i32, vp, i32, i32, f32, vp]
    L.ga_linear_small.argtypes = [vp, vp, vp, vp, i32, i32, i32, i32, i32, i32, vp]
    L.ga_timestep_sinusoid.argtypes = [vp, vp, i32, i32, vp]
    L.ga_layernorm_rows.argtypes = [vp, vp, vp, vp, i32, i32, f32, vp]
    L.ga_add_tables.argtypes = [vp, vp, vp, i32, i32, i32, i32, vp]
    L.ga_embed_fc1.argtypes = [vp, i32, vp, i32, vp, vp, vp, i32, i32, vp]
    L.ga_xyz_posenc.argtypes = [vp, vp, i32, vp]
    L.ga_final_layer.argtypes = [vp, vp, vp, vp, vp, i32, i32, i32, i32, f32, vp]
    L.ga_cfg_combine.argtypes = [vp, vp, i64, f32, vp]
    L.ga_axpy.argtypes = [vp, vp, f32, i64, vp]
    L.ga_f32_to_bf16.argtypes = [vp, vp, i64, vp]
    for n in ("ga_gemm_bf16_tn", "ga_attention_bf16", "ga_rmsnorm_modulate", "ga_linear_small",
              "ga_timestep_sinusoid", "ga_layernorm_rows", "ga_add_tables", "ga_embed_fc1", "ga_xyz_posenc",
              "ga_final_layer", "ga_cfg_combine", "ga_axpy", "ga_f32_to_bf16"):
        getattr(L, n).restype = i32
    _bound = True
    return L


def _ck(rc, what):
    if rc != 0:
        raise RuntimeError("libga_b200: %s failed with code %d" % (what, rc))


_GEMM_CFG_ENV = None


def _gemm_config(M, N, mode=None):
    """Tile width (+ 1000 * cluster size, ga_b200.h) from a two-term cost model fitted to the round-2 sweep
    (tools/sweep_gemm.py, profiles/r02_gemm_sweep.txt, cuBLAS column as yardstick):

        cost(BN) = ceil(tiles(BN) / 148) * (BN + 64)

    -- waves of the persistent grid times the per-tile work (the main loop scales with BN, the +64 is the fixed
    TMA-fill / epilogue-drain share that makes wide tiles more efficient per byte).  It reproduces the measured winner on
    all nine DiT shapes: 192 for 4096x768 (K = 768: 8.1 vs 9.5 us; K = 3072: 16.6 vs 21.8), 4096x3072 and 1536x4096,
    256 for 4096x2304 and 1536x3072, 128 for the under-filled 1536x1024 GEMMs of the deployed size.
    The HEADS epilogue (whole 64-wide heads per half tile: 128 or 256 only) stays at 128.  GA_B200_GEMM_CFG="big,small" overrides."""
    global _GEMM_CFG_ENV
    if _GEMM_CFG_ENV is None:
        import os
        _GEMM_CFG_ENV = os.environ.get("GA_B200_GEMM_CFG", "")
    if _GEMM_CFG_ENV:
        big, small = (int(v) for v in _GEMM_CFG_ENV.split(","))
        return big if (M >= 1024 and N >= 512) else small
    rows = -(-M // 128)
    best, best_cost = 128, None
    if mode == EPI_HEADS:
        return 128                      # 256-wide HEADS tiles measured slower at M = 1536 (deployed qkv: +1.7 % per NFE)
    for bn in (128, 192, 256):
        if bn > 128 and N < bn:
            continue
        tiles = rows * -(-N // bn)
        cost = -(-tiles // 148) * (bn + 64)
        if best_cost is None or cost < best_cost:
            best, best_cost = bn, cost
    return best


def _round_up(x, m):
    return (x + m - 1) // m * m


# --------------------------------------------------------------------------
# parameter containers with the reference's names (no forward of their own)
# --------------------------------------------------------------------------
class _Weight(nn.Module):
    def __init__(self, n, init=1.0):
        super().__init__()
        self.weight = nn.Parameter(torch.full((n,), float(init)))


class _Bias(nn.Module):
    def __init__(self, n):
        super().__init__()
        self.bias = nn.Parameter(torch.zeros(n))


class _Mlp(nn.Module):          # timm Mlp layout: fc1, fc2
    def __init__(self, cin, hidden, cout):
        super().__init__()
        self.fc1 = nn.Linear(cin, hidden)
        self.fc2 = nn.Linear(hidden, cout)


class _SelfAttn(nn.Module):
    def __init__(self, dim, heads):
        super().__init__()
        self.qkv = nn.Linear(dim, 3 * dim, bias=True)
        self.proj = nn.Linear(dim, dim)
        self.q_norm = _Weight(dim // heads)
        self.k_norm = _Weight(dim // heads)


class _CrossAttn(nn.Module):
    def __init__(self, dim, ctx_dim, heads):
        super().__init__()
        self.to_q = nn.Linear(dim, dim, bias=False)
        self.to_k = nn.Linear(ctx_dim, dim, bias=False)
        self.q_norm = _Weight(dim // heads)
        self.k_norm = _Weight(dim // heads)
        self.to_v = nn.Linear(ctx_dim, dim, bias=False)
        self.to_out = nn.Sequential(nn.Linear(dim, dim), nn.Dropout(0.0))


class _FusedMLP(nn.Module):     # xformers FusedMLP key layout: mlp.{0.weight, 1.bias, 2.weight, 3.bias}
    def __init__(self, dim, mult):
        super().__init__()
        self.mlp = nn.Sequential(nn.Linear(dim, mult * dim, bias=False), _Bias(mult * dim),
                                 nn.Linear(mult * dim, dim, bias=False), _Bias(dim))


class ImageCondDiTBlockPixelArtRMSNormClayLRM(nn.Module):
    """Parameter layout of /root/reference/dit/dit_models_xformers.py:717-763 (CA -> gated SA -> gated FFN)."""

    def __init__(self, hidden_size, num_heads, context_dim, mlp_ratio=4, **kw):
        super().__init__()
        self.scale_shift_table = nn.Parameter(torch.randn(6, hidden_size) / hidden_size ** 0.5)
        self.norm1 = _Weight(hidden_size)
        self.norm2 = _Weight(hidden_size)
        self.attn = _SelfAttn(hidden_size, num_heads)
        self.mlp = _FusedMLP(hidden_size, int(mlp_ratio))
        self.attention_y_norm = _Weight(1024)                    # present, unused (reference :456-458)
        self.cross_attn_dino = _CrossAttn(hidden_size, context_dim, num_heads)
        self.prenorm_ca_dino = _Weight(hidden_size)
        self.adaLN_modulation = None


class _TimestepEmbedder(nn.Module):
    def __init__(self, hidden, freq=256):
        super().__init__()
        self.mlp = nn.Sequential(nn.Linear(freq, hidden), nn.SiLU(), nn.Linear(hidden, hidden))
        self.frequency_embedding_size = freq


class _FinalLayer(nn.Module):   # T2IFinalLayer (/root/reference/dit/dit_models_xformers.py:62-85)
    def __init__(self, hidden, out_channels):
        super().__init__()
        self.linear = nn.Linear(hidden, out_channels)
        self.scale_shift_table = nn.Parameter(torch.randn(2, hidden) / hidden ** 0.5)
        self.adaLN_modulation = None
        self.out_channels = out_channels


class _CaptionEmbedder(nn.Module):
    def __init__(self, cin, hidden):
        super().__init__()
        self.y_proj = _Mlp(cin, hidden, hidden)


class _XYZPosEmbed(nn.Module):
    def __init__(self, dim):
        super().__init__()
        self.xyz_projection = nn.Linear(63, dim)


class DiT_I23D_PCD_PixelArt_noclip(nn.Module):
    def __init__(self, input_size=32, patch_size=2, in_channels=4, hidden_size=1152, depth=28, num_heads=16,
                 mlp_ratio=4, class_dropout_prob=0.1, num_classes=1000, learn_sigma=True, mixing_logit_init=-3,
                 mixed_prediction=True, context_dim=False, pooling_ctx_dim=768, roll_out=False, vit_blk=None,
                 final_layer_blk=None, create_cap_embedder=True, use_clay_ca=False, has_caption=False,
                 rope_scaling_factor=1.0, ntk_factor=1.0, enable_rope=False, **kw):
        super().__init__()
        if enable_rope:
            raise NotImplementedError("enable_rope=True is dead code in the reference (SURVEY.md F6)")
        if has_caption:
            raise NotImplementedError("caption conditioning is not on the deployed i23d path")
        if hidden_size % num_heads or hidden_size // num_heads != 64:
            raise ValueError("the B200 attention kernel is specialised for head_dim 64 (all reference archs)")
        assert roll_out
        self.depth, self.mlp_ratio, self.learn_sigma = depth, mlp_ratio, learn_sigma
        self.in_channels = in_channels
        self.out_channels = in_channels * 2 if learn_sigma else in_channels
        self.patch_size, self.num_heads, self.embed_dim = patch_size, num_heads, hidden_size
        self.roll_out, self.plane_n, self.context_dim = roll_out, 3, context_dim
        self.enable_rope, self.freqs_cis, self.use_clay_ca, self.has_caption = False, None, use_clay_ca, False
        self.x_embed_in = in_channels
        self.x_embedder = _Mlp(in_channels, hidden_size, hidden_size)
        self.t_embedder = _TimestepEmbedder(hidden_size)
        self.y_embedder = None
        self.blocks = nn.ModuleList([
            ImageCondDiTBlockPixelArtRMSNormClayLRM(hidden_size, num_heads, context_dim, mlp_ratio)
            for _ in range(depth)])
        self.final_layer = _FinalLayer(hidden_size, self.out_channels)
        self.clip_spatial_proj = _CaptionEmbedder(1024, hidden_size)          # present, unused
        self.adaLN_modulation = nn.Sequential(nn.SiLU(), nn.Linear(hidden_size, 6 * hidden_size))
        self.cap_embedder = nn.Sequential(nn.LayerNorm(pooling_ctx_dim), nn.Linear(pooling_ctx_dim, hidden_size)) \
            if create_cap_embedder else nn.Identity()
        self.attention_y_norm = _Weight(1024)                                  # present, unused
        self.pooled_vec_embedder = nn.Sequential(nn.LayerNorm(context_dim), nn.Linear(context_dim, hidden_size))
        self.initialize_weights()
        self._engine = None
        self.cfg_dedup = False             # opt-in: skip the duplicate CFG half when uc == c bit for bit (forward_with_cfg)

    # reference init (dit_models_xformers.py:1117-1159, dit_i23d.py:213-214,508-509)
    def initialize_weights(self):
        for m in self.modules():
            if isinstance(m, nn.Linear):
                nn.init.xavier_uniform_(m.weight)
                if m.bias is not None:
                    nn.init.constant_(m.bias, 0)
        nn.init.normal_(self.t_embedder.mlp[0].weight, std=0.02)
        nn.init.normal_(self.t_embedder.mlp[2].weight, std=0.02)
        for lin in (self.final_layer.linear, self.adaLN_modulation[1], self.pooled_vec_embedder[1]):
            nn.init.constant_(lin.weight, 0)
            nn.init.constant_(lin.bias, 0)
        if isinstance(self.cap_embedder, nn.Sequential):
            nn.init.constant_(self.cap_embedder[1].weight, 0)
            nn.init.constant_(self.cap_embedder[1].bias, 0)

    def randomize_zero_init_(self, std=0.02, seed=0):
        """the zero-initialised tensors would make the output identically 0."""
        g = torch.Generator().manual_seed(seed)
        for lin in (self.final_layer.linear, self.adaLN_modulation[1], self.pooled_vec_embedder[1]):
            lin.weight.data.copy_(torch.randn(lin.weight.shape, generator=g) * std)
            lin.bias.data.copy_(torch.randn(lin.bias.shape, generator=g) * std)
        self.invalidate()
        return self

    def invalidate(self):
        """Call after changing parameters in place (the bf16 weight pack is cached)."""
        self._engine = None

    def load_state_dict(self, *a, **k):
        r = super().load_state_dict(*a, **k)
        self.invalidate()
        return r

    def _apply(self, fn, *a, **k):
        r = super()._apply(fn, *a, **k)
        self._engine = None
        return r

    # ------------------------------------------------------------------ forward
    def _extra_embed(self, context):
        return None, None          # stage 1: no xyz conditioning

    def _engine_for(self, x, context):
        assert isinstance(context, dict)
        if not x.is_cuda:
            raise RuntimeError("gaussiananything_b200 DiT needs CUDA tensors: there is no CPU fallback")
        if torch.is_grad_enabled():
            # inference-only mirror (the denoiser FORWARD under the ODE sampler): nothing here records
            # an autograd graph, so a caller expecting gradients must be told instead of silently getting none
            wants = x.requires_grad or any(isinstance(v, torch.Tensor) and v.requires_grad for v in context.values())
            if wants or (self.training and any(p.requires_grad for p in self.parameters())):
                raise RuntimeError(
                    "gaussiananything_b200 DiT is inference-only (no autograd through the CUDA kernels): call it under "
                    "torch.no_grad() / torch.inference_mode(), or after .eval() with inputs that do not require grad")
        if self._engine is None or self._engine.device != x.device:
            self._engine = _DiTEngine(self, x.device)
        elif self._engine.param_versions != _param_versions(self):
            self._engine = _DiTEngine(self, x.device)          # parameters changed in place: stale bf16 pack / graph
        return self._engine

    def forward(self, x, timesteps=None, context=None, y=None, get_attr='', **kwargs):
        return self._engine_for(x, context).run(x, timesteps, context, cfg_scale=None)

    def forward_with_cfg(self, x, t, context, cfg_scale):
        """/root/reference/dit/dit_i23d.py:159-172: one 2B forward, u + s (c - u), duplicated.

        `self.cfg_dedup = True` (opt-in): when the conditional and unconditional halves of x, t and
        every context tensor are bit-identical -- the reference's stage-2 call, where `uc == c`
        (nsr/lsgm/flow_matching_trainer.py:2004, sgm/modules/encoders/modules.py:166-168) -- u + s (c - u) == c
        exactly, so only B rows are evaluated and duplicated.  Same bits as the 2B call (tests/test_dit_gpu.py)."""
        eng = self._engine_for(x, context)
        if getattr(self, "cfg_dedup", False) and x.shape[0] % 2 == 0 and eng.halves_identical(x, t, context):
            h = x.shape[0] // 2
            ctx_h = eng.half_context(context)
            y = eng.run(x[:h], t[:h] if t.numel() > 1 else t, ctx_h, cfg_scale=None)
            return torch.cat([y, y], 0)
        return eng.run(x, t, context, cfg_scale=float(cfg_scale))


def _param_versions(model):
    """Sum of the parameters' in-place version counters (an optimizer step / .data.copy_ / load bumps it)."""
    tot = 0
    for p in model.parameters():
        try:
            tot += p._version
        except RuntimeError:              # inference tensors do not track versions
            pass
    return tot


class DiT_I23D_PCD_PixelArt_noclip_clay_stage2(DiT_I23D_PCD_PixelArt_noclip):
    def __init__(self, *a, use_pe_cond=False, **kw):
        super().__init__(*a, **kw)
        self.has_caption = False
        self.use_pe_cond = use_pe_cond
        extra = 0 if use_pe_cond else 3
        self.x_embed_in = self.in_channels + extra
        self.x_embedder = _Mlp(self.in_channels + extra, self.embed_dim, self.embed_dim)
        nn.init.xavier_uniform_(self.x_embedder.fc1.weight); nn.init.constant_(self.x_embedder.fc1.bias, 0)
        nn.init.xavier_uniform_(self.x_embedder.fc2.weight); nn.init.constant_(self.x_embedder.fc2.bias, 0)
        if use_pe_cond:
            self.xyz_pos_embed = _XYZPosEmbed(self.embed_dim)
            nn.init.xavier_uniform_(self.xyz_pos_embed.xyz_projection.weight)
            nn.init.constant_(self.xyz_pos_embed.xyz_projection.bias, 0)


# --------------------------------------------------------------------------
# engine: bf16 weight pack + workspaces + launch sequence (+ CUDA graph)
# --------------------------------------------------------------------------
class _DiTEngine:
    def __init__(self, model, device):
        self.L = _bind()
        self.device = device
        self.m = model
        self.D, self.H, self.depth = model.embed_dim, model.num_heads, model.depth
        self.Cin, self.Cout = model.in_channels, model.out_channels
        self.Dc = model.context_dim
        self.stage2 = isinstance(model, DiT_I23D_PCD_PixelArt_noclip_clay_stage2)
        self.use_pe = self.stage2 and model.use_pe_cond
        self.use_graph = True
        self.tap_blocks = False            # tests: keep the residual stream after every block (s["taps"][l])
        self.param_versions = _param_versions(model)
        with torch.inference_mode(False), torch.cuda.device(device):
            self._pack()
        self._shape = None
        self._ctx_ref = None               # STRONG reference to the context tensor whose K/V are cached
        self._ctx_ver = None
        self._half_ctx = None
        self._graph = None

    # ---- weights
    def _pack(self):
        m, dev = self.m, self.device
        f32 = lambda t: t.detach().to(device=dev, dtype=torch.float32).contiguous()
        b16 = lambda t: t.detach().to(device=dev, dtype=torch.bfloat16).contiguous()
        w = {}
        w["t0_w"], w["t0_b"] = f32(m.t_embedder.mlp[0].weight), f32(m.t_embedder.mlp[0].bias)
        w["t2_w"], w["t2_b"] = f32(m.t_embedder.mlp[2].weight), f32(m.t_embedder.mlp[2].bias)
        w["pv_ln_w"], w["pv_ln_b"] = f32(m.pooled_vec_embedder[0].weight), f32(m.pooled_vec_embedder[0].bias)
        w["pv_w"], w["pv_b"] = f32(m.pooled_vec_embedder[1].weight), f32(m.pooled_vec_embedder[1].bias)
        w["ada_w"], w["ada_b"] = f32(m.adaLN_modulation[1].weight), f32(m.adaLN_modulation[1].bias)
        w["fc1_w"], w["fc1_b"] = f32(m.x_embedder.fc1.weight), f32(m.x_embedder.fc1.bias)
        w["fc2_w"], w["fc2_b"] = b16(m.x_embedder.fc2.weight), f32(m.x_embedder.fc2.bias)
        if self.use_pe:
            wx = torch.zeros(self.D, 64, device=dev, dtype=torch.float32)
            wx[:, :63] = f32(m.xyz_pos_embed.xyz_projection.weight)
            w["xyz_w"], w["xyz_b"] = wx.to(torch.bfloat16).contiguous(), f32(m.xyz_pos_embed.xyz_projection.bias)
        w["tables"] = torch.stack([f32(b.scale_shift_table) for b in m.blocks]).contiguous()      # [L,6,D]
        w["table_f"] = f32(m.final_layer.scale_shift_table)
        w["fin_w"], w["fin_b"] = f32(m.final_layer.linear.weight), f32(m.final_layer.linear.bias)
        blocks = []
        for b in m.blocks:
            ca, sa, mlp = b.cross_attn_dino, b.attn, b.mlp.mlp
            blocks.append(dict(
                pre_w=f32(b.prenorm_ca_dino.weight), n1_w=f32(b.norm1.weight), n2_w=f32(b.norm2.weight),
                caq_w=b16(ca.to_q.weight), cakv_w=b16(torch.cat([ca.to_k.weight, ca.to_v.weight], 0)),
                caq_n=f32(ca.q_norm.weight), cak_n=f32(ca.k_norm.weight),
                cao_w=b16(ca.to_out[0].weight), cao_b=f32(ca.to_out[0].bias),
                qkv_w=b16(sa.qkv.weight), qkv_b=f32(sa.qkv.bias), q_n=f32(sa.q_norm.weight), k_n=f32(sa.k_norm.weight),
                proj_w=b16(sa.proj.weight), proj_b=f32(sa.proj.bias),
                w1=b16(mlp[0].weight), b1=f32(mlp[1].bias), w2=b16(mlp[2].weight), b2=f32(mlp[3].bias),
                # |q.k|/8 <= (sqrt(64) max|wq|)(sqrt(64) max|wk|)/8 for RMS-normalised q, k (+2% for bf16 rounding):
                # lets the attention kernel skip the running maximum (ga_b200.h: score_bound)
                ca_bound=8.16 * float(ca.q_norm.weight.detach().abs().max()) * float(ca.k_norm.weight.detach().abs().max()),
                sa_bound=8.16 * float(sa.q_norm.weight.detach().abs().max()) * float(sa.k_norm.weight.detach().abs().max())))
        self.w, self.wb = w, blocks

    # ---- workspaces for a (B, N, M) problem
    def _alloc(self, B, N, M):
        dev, D, H = self.device, self.D, self.H
        R = B * N
        z = lambda *s, dt=torch.float32: torch.zeros(*s, device=dev, dtype=dt)
        bf = torch.bfloat16
        self.Np, self.Mp = _round_up(N, 128), _round_up(M, 128)
        s = dict(
            x_in=z(B, N, self.Cin), t_in=z(B), vec_in=z(B, self.Dc), y=z(B, N, self.Cout), y_cfg=z(B, N, self.Cout),
            sinus=z(B, 256), h1=z(B, D), temb=z(B, D), vln=z(B, self.Dc), t0=z(B, 6 * D),
            mod=z(self.depth, B, 6 * D), modf=z(B, 2 * D),
            e1=z(R, D, dt=bf), xres=z(R, D), h=z(R, D, dt=bf), ao=z(R, D, dt=bf), hid=z(R, 4 * D, dt=bf),
            q=z(B * H, self.Np, 64, dt=bf), k=z(B * H, self.Np, 64, dt=bf), vt=z(B * H, 64, self.Np, dt=bf),
            ctx=z(B * M, self.Dc, dt=bf),
            kc=z(self.depth, B * H, self.Mp, 64, dt=bf), vtc=z(self.depth, B * H, 64, self.Mp, dt=bf))
        if self.stage2:
            s["xyz_in"] = z(B, N, 3)
            if self.use_pe:
                s["pe"] = z(R, 64, dt=bf)
        self.s = s
        self._shape = (B, N, M)
        self._graph = None
        self._ctx_ref = None

    # ---- launches
    def _gemm(self, A, W, M, N, K, epi, st, bn=None):
        if bn is None:
            bn = _gemm_config(M, N, epi.mode)
        _ck(self.L.ga_gemm_bf16_tn(_p(A), K, _p(W), K, M, N, K, C.byref(epi), bn, st), "ga_gemm_bf16_tn")

    def _epi(self, mode, **kw):
        e = GaGemmEpilogue()
        e.mode = mode
        e.eps = 1e-5
        for k, v in kw.items():
            setattr(e, k, v.data_ptr() if isinstance(v, torch.Tensor) else v)
        return e

    def _context_kv(self, st):
        """Timestep-independent cross-attention K/V of every block: once per context."""
        B, N, M = self._shape
        s, D, H = self.s, self.D, self.H
        for l, wb in enumerate(self.wb):
            e = self._epi(EPI_HEADS, k=s["kc"][l], vt=s["vtc"][l], kn_w=wb["cak_n"], heads=H, first_part=1,
                          tok_pitch=self.Mp, rows_per_batch=M)
            self._gemm(s["ctx"], wb["cakv_w"], B * M, 2 * D, self.Dc, e, st)

    def _forward_launches(self, st, cfg_scale):
        L, s, w = self.L, self.s, self.w
        B, N, M = self._shape
        D, H, R = self.D, self.H, B * N
        # ---- prologue: timestep + pooled vector -> adaLN tables
        _ck(L.ga_timestep_sinusoid(_p(s["t_in"]), _p(s["sinus"]), B, 256, st), "sinusoid")
        _ck(L.ga_linear_small(_p(s["sinus"]), _p(w["t0_w"]), _p(w["t0_b"]), _p(s["h1"]), B, D, 256, 0, 1, 0, st), "t0")
        _ck(L.ga_linear_small(_p(s["h1"]), _p(w["t2_w"]), _p(w["t2_b"]), _p(s["temb"]), B, D, D, 0, 0, 0, st), "t2")
        _ck(L.ga_layernorm_rows(_p(s["vec_in"]), _p(w["pv_ln_w"]), _p(w["pv_ln_b"]), _p(s["vln"]), B, self.Dc, 1e-5, st), "ln")
        _ck(L.ga_linear_small(_p(s["vln"]), _p(w["pv_w"]), _p(w["pv_b"]), _p(s["temb"]), B, D, self.Dc, 0, 0, 1, st), "pv")
        _ck(L.ga_linear_small(_p(s["temb"]), _p(w["ada_w"]), _p(w["ada_b"]), _p(s["t0"]), B, 6 * D, D, 1, 0, 0, st), "ada")
        _ck(L.ga_add_tables(_p(w["tables"]), _p(s["t0"]), _p(s["mod"]), self.depth, B, 6 * D, 6 * D, st), "tables")
        _ck(L.ga_add_tables(_p(w["table_f"]), _p(s["temb"]), _p(s["modf"]), 1, B, 2 * D, D, st), "table_f")
        # ---- token embedder
        concat = self.stage2 and not self.use_pe
        _ck(L.ga_embed_fc1(_p(s["x_in"]), self.Cin, _p(s["xyz_in"]) if concat else None, 3 if concat else 0,
                           _p(w["fc1_w"]), _p(w["fc1_b"]), _p(s["e1"]), R, D, st), "embed_fc1")
        self._gemm(s["e1"], w["fc2_w"], R, D, D, self._epi(EPI_F32, bias=w["fc2_b"], out=s["xres"], ld_out=D), st)
        if self.use_pe:
            _ck(L.ga_xyz_posenc(_p(s["xyz_in"]), _p(s["pe"]), R, st), "xyz_pe")
            self._gemm(s["pe"], w["xyz_w"], R, D, 64,
                       self._epi(EPI_RESID_GATE_F32, bias=w["xyz_b"], out=s["xres"], ld_out=D, rows_per_batch=N), st)
        # ---- blocks
        scale = 1.0 / math.sqrt(64.0)
        for l, wb in enumerate(self.wb):
            mod = s["mod"][l]                       # [B, 6D]
            ch = lambda j: mod[:, j * D:(j + 1) * D]
            # cross attention (pre-norm, residual)
            _ck(L.ga_rmsnorm_modulate(_p(s["xres"]), _p(wb["pre_w"]), None, None, 0, N, _p(s["h"]), R, D, 1e-5, st), "prenorm")
            self._gemm(s["h"], wb["caq_w"], R, D, D,
                       self._epi(EPI_HEADS, q=s["q"], qn_w=wb["caq_n"], heads=H, first_part=0, tok_pitch=self.Np,
                                 rows_per_batch=N), st)
            _ck(L.ga_attention_bf16(_p(s["q"]), _p(s["kc"][l]), _p(s["vtc"][l]), _p(s["ao"]), B, H, N, M, self.Np,
                                    self.Mp, scale, wb["ca_bound"], st), "cross attention")
            self._gemm(s["ao"], wb["cao_w"], R, D, D,
                       self._epi(EPI_RESID_GATE_F32, bias=wb["cao_b"], out=s["xres"], ld_out=D, rows_per_batch=N), st)
            # gated self attention
            _ck(L.ga_rmsnorm_modulate(_p(s["xres"]), _p(wb["n1_w"]), _p(ch(0)), _p(ch(1)), 6 * D, N, _p(s["h"]), R, D,
                                      1e-5, st), "norm1")
            self._gemm(s["h"], wb["qkv_w"], R, 3 * D, D,
                       self._epi(EPI_HEADS, bias=wb["qkv_b"], q=s["q"], k=s["k"], vt=s["vt"], qn_w=wb["q_n"],
                                 kn_w=wb["k_n"], heads=H, first_part=0, tok_pitch=self.Np, rows_per_batch=N), st)
            _ck(L.ga_attention_bf16(_p(s["q"]), _p(s["k"]), _p(s["vt"]), _p(s["ao"]), B, H, N, N, self.Np, self.Np,
                                    scale, wb["sa_bound"], st), "self attention")
            self._gemm(s["ao"], wb["proj_w"], R, D, D,
                       self._epi(EPI_RESID_GATE_F32, bias=wb["proj_b"], out=s["xres"], ld_out=D, gate=ch(2),
                                 gate_ld=6 * D, rows_per_batch=N), st)
            # gated FFN
            _ck(L.ga_rmsnorm_modulate(_p(s["xres"]), _p(wb["n2_w"]), _p(ch(3)), _p(ch(4)), 6 * D, N, _p(s["h"]), R, D,
                                      1e-5, st), "norm2")
            self._gemm(s["h"], wb["w1"], R, 4 * D, D,
                       self._epi(EPI_GELU_BF16, bias=wb["b1"], out=s["hid"], ld_out=4 * D), st)
            self._gemm(s["hid"], wb["w2"], R, D, 4 * D,
                       self._epi(EPI_RESID_GATE_F32, bias=wb["b2"], out=s["xres"], ld_out=D, gate=ch(5),
                                 gate_ld=6 * D, rows_per_batch=N), st)
            if self.tap_blocks:
                s["taps"][l].copy_(s["xres"])
        # ---- final layer (+ CFG combine)
        _ck(L.ga_final_layer(_p(s["xres"]), _p(s["modf"]), _p(w["fin_w"]), _p(w["fin_b"]), _p(s["y"]), R, D,
                             self.Cout, N, 1e-6, st), "final layer")
        if cfg_scale is not None:
            _ck(L.ga_cfg_combine(_p(s["y"]), _p(s["y_cfg"]), (B // 2) * N * self.Cout, cfg_scale, st), "cfg")

    @staticmethod
    def _version(t):
        try:
            return t._version
        except RuntimeError:               # inference tensors (torch.inference_mode) do not track versions
            return None

    def invalidate_context(self):
        """Forget the cached cross-attention K/V (needed only after an in-place change of a context tensor that was
        created under torch.inference_mode, where no version counter exists)."""
        self._ctx_ref = None

    def _context_is_cached(self, ctx_tok):
        # The cache holds a STRONG reference to the tensor it was computed from: while it is alive its address
        # cannot be handed to another tensor, so object identity (+ the in-place version counter) is a sound key.
        # Round 1 keyed on data_ptr(): a freed context's block is routinely recycled for the next sample's context.
        return (self._ctx_ref is not None and ctx_tok is self._ctx_ref
                and self._version(ctx_tok) == self._ctx_ver)

    def halves_identical(self, x, t, context):
        """cond | uncond halves bit-identical?  Context tensors: decided once per context object; x, t: per call."""
        key = tuple((k, id(v), self._version(v)) for k, v in sorted(context.items()) if isinstance(v, torch.Tensor))
        if self._half_ctx is None or self._half_ctx[0] != key:
            same = True
            for v in context.values():
                if isinstance(v, torch.Tensor):
                    h = v.shape[0] // 2
                    same = same and v.shape[0] % 2 == 0 and bool(torch.equal(v[:h], v[h:]))
            half = {k: (v[:v.shape[0] // 2].contiguous() if isinstance(v, torch.Tensor) else v)
                    for k, v in context.items()} if same else None
            self._half_ctx = (key, same, half, list(context.values()))      # the list pins the ids
        if not self._half_ctx[1]:
            return False
        h = x.shape[0] // 2
        if t.numel() > 1 and not bool(torch.equal(t.reshape(-1)[:h], t.reshape(-1)[h:])):
            return False
        return bool(torch.equal(x[:h], x[h:]))

    def half_context(self, context):
        return self._half_ctx[2]

    def run(self, x, t, context, cfg_scale):
        with torch.cuda.device(self.device):
            return self._run(x, t, context, cfg_scale)

    def _run(self, x, t, context, cfg_scale):
        ctx_tok, vec = context["img_crossattn"], context["img_vector"]
        B, N, _ = x.shape
        M = ctx_tok.shape[1]
        if cfg_scale is not None and B % 2:
            raise ValueError("forward_with_cfg needs an even batch (cond | uncond)")
        if self._shape != (B, N, M):
            with torch.inference_mode(False):          # workspaces outlive this call: never inference tensors
                self._alloc(B, N, M)
        s = self.s
        if self.tap_blocks and "taps" not in s:
            with torch.inference_mode(False):
                s["taps"] = torch.zeros(self.depth, B * N, self.D, device=self.device)
        dev = self.device
        stream = torch.cuda.current_stream(dev)
        st = C.c_void_p(stream.cuda_stream)
        s["x_in"].copy_(x.reshape(B, N, self.Cin).to(torch.float32))
        s["t_in"].copy_(t.reshape(-1).to(torch.float32).expand(B) if t.numel() == 1 else t.reshape(B).to(torch.float32))
        s["vec_in"].copy_(vec.reshape(B, self.Dc).to(torch.float32))
        if self.stage2:
            s["xyz_in"].copy_(context["fps-xyz"].reshape(B, N, 3).to(torch.float32))
        if not self._context_is_cached(ctx_tok):
            c32 = ctx_tok.reshape(B * M, self.Dc).to(torch.float32).contiguous()
            _ck(self.L.ga_f32_to_bf16(_p(c32), _p(s["ctx"]), c32.numel(), st), "ctx->bf16")
            self._context_kv(st)
            self._ctx_ref, self._ctx_ver = ctx_tok, self._version(ctx_tok)
        gkey = (cfg_scale, self.tap_blocks)
        if self.use_graph:
            if self._graph is None or self._graph[0] != gkey:
                # warm-up run (sets kernel attributes), then capture the same launch sequence
                self._forward_launches(st, cfg_scale)
                torch.cuda.synchronize(dev)
                g = torch.cuda.CUDAGraph()
                with torch.cuda.graph(g):
                    cst = C.c_void_p(torch.cuda.current_stream(dev).cuda_stream)
                    self._forward_launches(cst, cfg_scale)
                self._graph = (gkey, g)
            self._graph[1].replay()
        else:
            self._forward_launches(st, cfg_scale)
        out = s["y_cfg"] if cfg_scale is not None else s["y"]
        return out.clone()

    @property
    def launches_per_forward(self):
        per_block = 11
        return 8 + 2 + (2 if self.use_pe else 0) + per_block * self.depth + 1


# --------------------------------------------------------------------------
# registry (names of /root/reference/dit/dit_i23d.py:1665-1697 that map to this block type)
# --------------------------------------------------------------------------
def DiT_L_Pixelart_clay_pcd(**kw):
    return DiT_I23D_PCD_PixelArt_noclip(depth=24, use_clay_ca=True, hidden_size=1024, patch_size=1, num_heads=16,
                                        enable_rope=False, **kw)


def DiT_B_Pixelart_clay_pcd(**kw):
    return DiT_I23D_PCD_PixelArt_noclip(depth=12, use_clay_ca=True, hidden_size=768, patch_size=1, num_heads=12, **kw)


def DiT_L_Pixelart_clay_pcd_stage2(**kw):
    return DiT_I23D_PCD_PixelArt_noclip_clay_stage2(depth=24, use_clay_ca=True, hidden_size=1024, patch_size=1,
                                                    num_heads=16, use_pe_cond=True, **kw)


def DiT_B_Pixelart_clay_pcd_stage2(**kw):
    # As in the reference (dit_i23d.py:1554-1559) this entry leaves num_heads at its default of 16, i.e.
    # head_dim 48, and conditions by concatenation (use_pe_cond=False).  head_dim 48 is not supported by the
    # B200 attention kernel, so constructing it raises ValueError; the deployed stage-2 model is the -L entry.
    return DiT_I23D_PCD_PixelArt_noclip_clay_stage2(depth=12, use_clay_ca=True, hidden_size=768, **kw)


DiT_models = {
    'DiT-PixArt-PCD-CLAY-L': DiT_L_Pixelart_clay_pcd,
    'DiT-PixArt-PCD-CLAY-B': DiT_B_Pixelart_clay_pcd,
    'DiT-PixArt-PCD-CLAY-stage2-L': DiT_L_Pixelart_clay_pcd_stage2,
    'DiT-PixArt-PCD-CLAY-stage2-B': DiT_B_Pixelart_clay_pcd_stage2,
}
