"""ctypes driver of baseline/libga_standin.so: the reference-algorithm GPU baseline of the rasteriser, driven the way
/root/reference/nsr/gs_surfel.py:65-114 drives the upstream package -- one launch set and one device->host read of
num_rendered PER VIEW.  Used only by bench.py (`gpu_standin`) and tests/test_standin_gpu.py."""
import ctypes as C
import os

import torch

from . import build as _build

_L = None


def lib():
    global _L
    if _L is None:
        L = C.CDLL(_build.build())
        vp, i32, f32 = C.c_void_p, C.c_int, C.c_float
        L.st_create.restype = vp
        L.st_create.argtypes = [i32, i32, i32]
        L.st_destroy.argtypes = [vp]
        L.st_forward.argtypes = [vp, vp, vp, vp, vp, f32, vp, vp, vp, vp]
        L.st_forward.restype = i32
        L.st_backward.argtypes = [vp, vp, vp, vp, vp, f32, vp, vp, vp, vp]
        L.st_backward.restype = i32
        _L = L
    return _L


class StandinRasterizer:
    def __init__(self, P, H, W, views, device="cuda:0"):
        self.L = lib()
        self.P, self.H, self.W, self.V = P, H, W, views
        self.dev = torch.device(device)
        self.ctx = [self.L.st_create(P, H, W) for _ in range(views)]
        assert all(self.ctx), "st_create failed"

    def __del__(self):
        for c in getattr(self, "ctx", []):
            self.L.st_destroy(C.c_void_p(c))

    def forward(self, g13, vms, pms, bg, scale_modifier=1.0):
        """g13 [P,13], vms/pms [V,4,4] -> color [V,3,H,W], allmap [V,7,H,W], radii [V,P], num_rendered list."""
        p = lambda t: C.c_void_p(t.data_ptr())
        st = C.c_void_p(torch.cuda.current_stream(self.dev).cuda_stream)
        V, H, W, P = self.V, self.H, self.W, self.P
        g13, vms, pms, bg = g13.contiguous().float(), vms.contiguous().float(), pms.contiguous().float(), bg.contiguous().float()
        color = torch.empty(V, 3, H, W, device=self.dev)
        allmap = torch.empty(V, 7, H, W, device=self.dev)
        radii = torch.empty(V, P, device=self.dev, dtype=torch.int32)
        self._saved = (g13, vms, pms, bg, scale_modifier)
        nr = []
        for v in range(V):                       # the reference's per-view loop
            n = self.L.st_forward(C.c_void_p(self.ctx[v]), p(g13), p(vms[v]), p(pms[v]), p(bg), scale_modifier, p(color[v]),
                                  p(allmap[v]), p(radii[v]), st)
            if n < 0:
                raise RuntimeError("st_forward failed: %d" % n)
            nr.append(n)
        return color, allmap, radii, nr

    def backward(self, d_color, d_allmap):
        p = lambda t: C.c_void_p(t.data_ptr())
        st = C.c_void_p(torch.cuda.current_stream(self.dev).cuda_stream)
        g13, vms, pms, bg, mod = self._saved
        d_color, d_allmap = d_color.contiguous().float(), d_allmap.contiguous().float()
        grad = torch.zeros(self.P, 13, device=self.dev)
        for v in range(self.V):
            rc = self.L.st_backward(C.c_void_p(self.ctx[v]), p(g13), p(vms[v]), p(pms[v]), p(bg), mod, p(d_color[v]), p(d_allmap[v]),
                                    p(grad), st)
            if rc != 0:
                raise RuntimeError("st_backward failed: %d" % rc)
        return grad
