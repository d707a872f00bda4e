"""Stubs of the third-party packages the reference imports but this image lacks, restating their published
semantics (see make_dit_golden.py for the list).  Importing this module puts /root/reference on sys.path and installs
the stubs; used only by the golden generators (build container only)."""
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
    if isinstance(y0, tuple):
        # torchdiffeq flattens a tuple state into one tensor, solves, and returns a tuple of per-component stacks
        shapes = [c.shape for c in y0]; sizes = [c[0].numel() if c.dim() > 1 else 1 for c in y0]; B = y0[0].shape[0]
        flat = lambda cs: torch.cat([c.reshape(B, -1) for c in cs], 1)
        unflat = lambda y: tuple(p.reshape(sh) for p, sh in zip(torch.split(y, sizes, 1), shapes))
        out = odeint(lambda tt, y: flat(func(tt, unflat(y))), flat(y0), t, method=method)
        return tuple(torch.stack([unflat(o)[i] for o in out], 0) for i in range(len(y0)))
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

