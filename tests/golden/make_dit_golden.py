"""Generates tests/golden/dit_*.npz by running the REFERENCE's own DiT and
transport code (/root/reference/dit/dit_i23d.py, dit/dit_models_xformers.py,
vit/vision_transformer.py, ldm/modules/attention.py, dit/norm.py, transport/*)
on CPU in fp32.  Third-party packages that are absent from this image are
replaced by stubs that restate their published semantics:
  * xformers.ops.memory_efficient_attention == softmax(q k^T / sqrt(d)) v  (-> torch SDPA)
  * xformers FusedMLP == Linear(no bias) -> +bias -> exact GELU -> Linear(no bias) -> +bias,
    state-dict keys mlp.mlp.{0.weight,1.bias,2.weight,3.bias}
  * timm Mlp / PatchEmbed; torchdiffeq.odeint fixed-grid euler/midpoint/heun/rk4
  * everything else missing (kiui, pytorch3d, ...) is a dummy: never on this path.
Only runs inside the build container (needs /root/reference); the .npz files it
writes are committed and are what tests/ read.  Weights are stored as bf16 bit
patterns (every value is exactly bf16-representable).
    python tests/golden/make_dit_golden.py
"""
import sys, types, importlib, torch, torch.nn as nn, torch.nn.functional as F
sys.path.insert(0, '/root/reference')

def mod(name, **attrs):
    import importlib.machinery as _im; m = types.ModuleType(name); m.__dict__.update(attrs); m.__spec__ = _im.ModuleSpec(name, None); m.__path__ = []; sys.modules[name] = m
    parent, _, child = name.rpartition('.')
    if parent and parent in sys.modules: setattr(sys.modules[parent], child, m)
    return m

# ---- xformers stub (published semantics of xformers 0.0.22)
def memory_efficient_attention(q, k, v, attn_bias=None, op=None, p=0.0, scale=None):
    # q,k,v: [B, M, H, K] or [B, M, K]
    if q.dim() == 4:
        o = F.scaled_dot_product_attention(q.transpose(1,2), k.transpose(1,2), v.transpose(1,2), scale=scale)
        return o.transpose(1,2)
    return F.scaled_dot_product_attention(q, k, v, scale=scale)
xf = mod('xformers', __version__='0.0.22.post7')
mod('xformers.ops', memory_efficient_attention=memory_efficient_attention, unbind=torch.unbind, fmha=types.SimpleNamespace(),
    MemoryEfficientAttentionFlashAttentionOp=None, MemoryEfficientAttentionCutlassOp=None)
mod('xformers.components')
import enum
class Activation(str, enum.Enum):
    GeLU='gelu'; ReLU='relu'
mod('xformers.components.activations', Activation=Activation, build_activation=lambda a: nn.GELU())
class FusedMLP(nn.Module):
    def __init__(self, dim_model, dropout, activation, hidden_layer_multiplier, bias=True, *a, **k):
        super().__init__()
        dim_mlp = hidden_layer_multiplier*dim_model
        class FusedDropoutBias(nn.Module):
            def __init__(s, n, act):
                super().__init__(); s.bias = nn.Parameter(torch.zeros(n)); s.act = act
            def forward(s, x):
                x = x + s.bias
                return F.gelu(x) if s.act else x
        self.mlp = nn.Sequential(nn.Linear(dim_model, dim_mlp, bias=False), FusedDropoutBias(dim_mlp, True),
                                 nn.Linear(dim_mlp, dim_model, bias=False), FusedDropoutBias(dim_model, False))
    def forward(self, x): return self.mlp(x)
mod('xformers.components.feedforward')
fm = mod('xformers.components.feedforward.fused_mlp', FusedMLP=FusedMLP)
sys.modules['xformers.components.feedforward'].fused_mlp = fm
sys.modules['xformers.components'].Activation = Activation

# ---- timm stub
class Mlp(nn.Module):
    def __init__(self, in_features, hidden_features=None, out_features=None, act_layer=nn.GELU, bias=True, drop=0., **kw):
        super().__init__()
        out_features = out_features or in_features; hidden_features = hidden_features or in_features
        self.fc1 = nn.Linear(in_features, hidden_features, bias=bias); self.act = act_layer()
        self.drop1 = nn.Dropout(drop); self.norm = nn.Identity()
        self.fc2 = nn.Linear(hidden_features, out_features, bias=bias); self.drop2 = nn.Dropout(drop)
    def forward(self, x): return self.drop2(self.fc2(self.norm(self.drop1(self.act(self.fc1(x))))))
class PatchEmbed(nn.Module):
    def __init__(self, img_size=224, patch_size=16, in_chans=3, embed_dim=768, bias=True, **kw):
        super().__init__()
        self.num_patches = (img_size//patch_size)**2
        self.proj = nn.Conv2d(in_chans, embed_dim, patch_size, patch_size, bias=bias)
    def forward(self, x): return self.proj(x).flatten(2).transpose(1,2)
mod('timm'); mod('timm.models'); mod('timm.models.vision_transformer', Mlp=Mlp, PatchEmbed=PatchEmbed, Attention=nn.Identity)
torch.cuda.is_available = lambda: True
# ---- generic stub for absent third-party modules (never on the DiT/transport math path)
import importlib.abc, importlib.machinery, os
class _Dummy:
    def __init__(self,*a,**k): pass
    def __call__(self,*a,**k): return _Dummy()
    def __getattr__(self,n): return _Dummy()
    def __mro_entries__(self, bases): return (object,)
class _StubModule(types.ModuleType):
    __path__ = []
    def __getattr__(self, n):
        if n.startswith('__'): raise AttributeError(n)
        return _Dummy()
ABSENT = {'lz4','point_cloud_utils','kiui','pytorch3d','open3d','trimesh','kornia','open_clip','clip','lpips','omegaconf','imageio','matplotlib','plyfile','pytorch_lightning','blobfile','mpi4py','webdataset','lmdb','cv2','skimage','mcubes','xatlas','nvdiffrast','tensorboardX','torch_scatter','skvideo','easydict','pymeshlab','rembg','gradio','diff_surfel_rasterization','diff_gaussian_rasterization','simple_knn','pytorch_fid','torchmetrics','taming','click','dnnlib_missing','vision_aided_loss','ftfy','piq','sklearn_missing', 'ipdb', 'apex_missing', 'fvcore', 'iopath','huggingface_hub_missing','safetensors_missing','timm_missing','torchvision_missing','wandb','deepspeed','accelerate_missing','diffusers','tyro','roma','pymcubes','Imath','OpenEXR','torch_efficient_distloss','flash_attn_missing','pointnet2_ops', 'threestudio', 'jaxtyping', 'typeguard', 'nerfacc', 'tinycudann', 'pysdf', 'igl', 'fpsample', 'torch_cluster', 'gsplat', 'spconv', 'ninja_missing'}
class _Finder(importlib.abc.MetaPathFinder, importlib.abc.Loader):
    def find_spec(self, name, path, target=None):
        top = name.split('.')[0]
        if top in ABSENT: return importlib.machinery.ModuleSpec(name, self, is_package=True)
        return None
    def create_module(self, spec): return _StubModule(spec.name)
    def exec_module(self, m): pass
sys.meta_path.append(_Finder())
# ---- torchdiffeq stub: fixed-grid solvers exactly on the given time grid (torchdiffeq FixedGridODESolver semantics)
def odeint(func, y0, t, *, method='dopri5', atol=None, rtol=None, **kw):
    assert method in ('euler','midpoint','rk4','heun2','heun'), method
    ys=[y0]; y=y0
    for i in range(len(t)-1):
        t0,t1=t[i],t[i+1]; dt=t1-t0
        if method=='euler': dy = dt*func(t0,y)
        elif method=='midpoint':
            h=0.5*dt; dy = dt*func(t0+h, y+h*func(t0,y))
        elif method in ('heun','heun2'):
            k1=func(t0,y); k2=func(t1,y+dt*k1); dy=0.5*dt*(k1+k2)
        else:
            k1=func(t0,y); k2=func(t0+dt/3,y+dt*k1/3); k3=func(t0+dt*2/3,y+dt*(k2-k1/3)); k4=func(t1,y+dt*(k1-k2+k3))
            dy=(k1+3*(k2+k3)+k4)*dt*0.125
        y=y+dy; ys.append(y)
    return torch.stack(ys,0)
mod('torchdiffeq', odeint=odeint)

import io, contextlib, os
import numpy as np
from dit import dit_i23d, dit_models_xformers as dmx
import transport as ref_transport

OUT = os.path.dirname(os.path.abspath(__file__))
UNUSED = ("clip_spatial_proj.", "cap_embedder.", "attention_y_norm.")


def bf16_round_(t):
    t.data.copy_(t.data.to(torch.bfloat16).to(torch.float32))


def build(cls, seed, **kw):
    torch.manual_seed(seed)
    with contextlib.redirect_stdout(io.StringIO()):
        m = cls(patch_size=1, vit_blk=dmx.ImageCondDiTBlockPixelArtRMSNormClayLRM, use_clay_ca=True,
                input_size=32, num_classes=0, learn_sigma=False, roll_out=True, pooling_ctx_dim=768, **kw)
    g = torch.Generator().manual_seed(seed + 1)
    for n, p in m.named_parameters():
        # the reference zero-initialises these (dit_i23d.py:213-214,508-509; dit_models_xformers.py:1158-1159):
        # re-randomise so the output is not identically zero; also make norm gains non-trivial
        if n.startswith(("final_layer.linear", "adaLN_modulation", "pooled_vec_embedder.1")) or n.endswith(".bias"):
            p.data.copy_(torch.randn(p.shape, generator=g) * 0.02)
        if "norm" in n and n.endswith(".weight"):
            p.data.copy_(1.0 + 0.1 * torch.randn(p.shape, generator=g))
        bf16_round_(p)
    return m.eval()


def run(name, cls, seed, B, N, M, Cin, ctx_dim, stage2, **kw):
    m = build(cls, seed, in_channels=Cin, context_dim=ctx_dim, **kw)
    g = torch.Generator().manual_seed(100 + seed)
    x = torch.randn(2 * B, N, Cin, generator=g)
    t = torch.rand(2 * B, generator=g)
    ctx = {"img_crossattn": torch.randn(2 * B, M, ctx_dim, generator=g),
           "img_vector": torch.randn(2 * B, ctx_dim, generator=g)}
    if stage2:
        ctx["fps-xyz"] = torch.rand(2 * B, N, 3, generator=g) * 2 - 1
    acts = {}
    hooks = [blk.register_forward_hook(lambda mod, a, out, i=i: acts.__setitem__("block%d" % i, out.detach().clone()))
             for i, blk in enumerate(m.blocks)]
    with torch.no_grad():
        y = m(x, t, ctx)
        y_cfg = m.forward_with_cfg(x, t, ctx, 4.0)
        for h in hooks:
            h.remove()
        tr = ref_transport.create_transport("GVP", "velocity", None, None, None, "lognorm")
        sampler = ref_transport.Sampler(tr)
        traj_euler = sampler.sample_ode(sampling_method="euler", num_steps=5)(x, m.forward_with_cfg, context=ctx, cfg_scale=4.0)
        traj_heun = sampler.sample_ode(sampling_method="heun2", num_steps=4)(x, m.forward_with_cfg, context=ctx, cfg_scale=4.0)
    sd = {k: v for k, v in m.state_dict().items() if not any(u in k for u in UNUSED)}
    save = {"sd__" + k: v.to(torch.bfloat16).view(torch.int16).numpy() for k, v in sd.items()}
    for k, v in sd.items():
        assert torch.equal(v.to(torch.bfloat16).to(torch.float32), v), k
    save.update(x=x.numpy(), t=t.numpy(), y=y.numpy(), y_cfg=y_cfg.numpy(),
                traj_euler=traj_euler.numpy(), traj_heun=traj_heun.numpy(),
                **{"ctx__" + k: v.numpy() for k, v in ctx.items()},
                **{"act__" + k: v.numpy() for k, v in acts.items()})
    save["meta"] = np.array([kw["depth"], kw["hidden_size"], kw["num_heads"], Cin, ctx_dim, int(stage2),
                             int(kw.get("use_pe_cond", False))])
    np.savez_compressed(os.path.join(OUT, name + ".npz"), **save)
    print(name, "params", sum(v.numel() for v in sd.values()), "|y|", float(y.abs().mean()))


run("dit_stage1_small", dit_i23d.DiT_I23D_PCD_PixelArt_noclip, 1, B=1, N=24, M=10, Cin=3, ctx_dim=64, stage2=False,
    depth=2, hidden_size=128, num_heads=2)
run("dit_stage2_small", dit_i23d.DiT_I23D_PCD_PixelArt_noclip_clay_stage2, 2, B=1, N=16, M=7, Cin=10, ctx_dim=32,
    stage2=True, depth=1, hidden_size=64, num_heads=1, use_pe_cond=True)
# stage 2 in concatenation mode (x_embedder sees cat([fps_xyz, x]); the stage2-B registry entry uses it)
run("dit_stage2_concat_small", dit_i23d.DiT_I23D_PCD_PixelArt_noclip_clay_stage2, 3, B=2, N=20, M=9, Cin=10, ctx_dim=32,
    stage2=True, depth=1, hidden_size=64, num_heads=1, use_pe_cond=False)
