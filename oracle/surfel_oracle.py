"""ctypes front-end of the C surfel-rasteriser oracle (oracle/surfel_oracle.c).

TEST INFRASTRUCTURE ONLY -- see the header of surfel_oracle.c.  PARITY UNPINNED
(the reference vendors neither the rasteriser nor any golden vectors).

`rasterize` mirrors what one call of the reference's
``GaussianRasterizer(raster_settings)(means3D, means2D, opacities, colors_precomp,
scales, rotations)`` computes (/root/reference/nsr/gs_surfel.py:85-114) and
additionally exposes every integer intermediate (radii, tiles_touched, rects,
sorted keys/ids, tile ranges, n_contrib) so the CUDA path can be compared
bit-for-bit.
"""
import ctypes as C
import os

import numpy as np

from . import build as _build

_LIB = None


def lib():
    global _LIB
    if _LIB is None:
        _LIB = C.CDLL(_build.build())
        _LIB.so_bin.restype = C.c_int64
    return _LIB


def set_num_threads(n):
    """Number of OpenMP threads the C oracle uses from now on (returns what is in effect)."""
    L = lib()
    L.so_set_num_threads.restype = C.c_int
    return int(L.so_set_num_threads(C.c_int(int(n))))


def set_variant(radius_formula=0, quat_norm_grad=0):
    """Selects the two unpinned judgement calls (see surfel_oracle.c); the defaults are (0, 0)."""
    lib().so_set_variant(C.c_int(int(radius_formula)), C.c_int(int(quat_norm_grad)))


def get_variant():
    a, b = C.c_int(0), C.c_int(0)
    lib().so_get_variant(C.byref(a), C.byref(b))
    return a.value, b.value


def _p(a, t):
    return a.ctypes.data_as(C.POINTER(t))


def _f32(a):
    return np.ascontiguousarray(np.asarray(a, dtype=np.float32))


def rasterize(means3D, opacities, scales, rotations, colors, viewmatrix, projmatrix,
              bg, H, W, scale_modifier=1.0):
    """Forward pass.  Returns a dict with outputs, state and integer intermediates."""
    L = lib()
    means3D, opacities, scales = _f32(means3D), _f32(opacities).reshape(-1), _f32(scales)
    rotations, colors = _f32(rotations), _f32(colors)
    vm, pm, bg = _f32(viewmatrix).reshape(16), _f32(projmatrix).reshape(16), _f32(bg)
    P = means3D.shape[0]
    transmat = np.zeros((P, 9), np.float32)
    normal_opacity = np.zeros((P, 4), np.float32)
    xy = np.zeros((P, 2), np.float32)
    depth = np.zeros(P, np.float32)
    radii = np.zeros(P, np.int32)
    tiles_touched = np.zeros(P, np.int32)
    rect = np.zeros((P, 4), np.int32)
    L.so_preprocess(C.c_int(P), _p(means3D, C.c_float), _p(opacities, C.c_float),
                    _p(scales, C.c_float), _p(rotations, C.c_float),
                    _p(vm, C.c_float), _p(pm, C.c_float), C.c_int(H), C.c_int(W),
                    C.c_float(scale_modifier),
                    _p(transmat, C.c_float), _p(normal_opacity, C.c_float), _p(xy, C.c_float),
                    _p(depth, C.c_float), _p(radii, C.c_int), _p(tiles_touched, C.c_int),
                    _p(rect, C.c_int))
    D = int(tiles_touched.astype(np.int64).sum())
    gx, gy = (W + 15) // 16, (H + 15) // 16
    keys = np.zeros(max(D, 1), np.uint64)
    ids = np.zeros(max(D, 1), np.uint32)
    ranges = np.zeros((gx * gy, 2), np.int32)
    got = L.so_bin(C.c_int(P), _p(depth, C.c_float), _p(radii, C.c_int), _p(rect, C.c_int),
                   C.c_int(H), C.c_int(W), C.c_int64(D), _p(keys, C.c_uint64),
                   _p(ids, C.c_uint32), _p(ranges, C.c_int32))
    assert got == D, (got, D)
    out_color = np.zeros((3, H, W), np.float32)
    out_allmap = np.zeros((7, H, W), np.float32)
    final_T = np.zeros((3, H, W), np.float32)
    n_contrib = np.zeros((2, H, W), np.int32)
    L.so_render_forward(C.c_int(H), C.c_int(W), _p(ranges, C.c_int32), _p(ids, C.c_uint32),
                        _p(xy, C.c_float), _p(transmat, C.c_float),
                        _p(normal_opacity, C.c_float), _p(colors, C.c_float), _p(bg, C.c_float),
                        _p(out_color, C.c_float), _p(out_allmap, C.c_float),
                        _p(final_T, C.c_float), _p(n_contrib, C.c_int32))
    return dict(color=out_color, allmap=out_allmap, radii=radii,
                transmat=transmat, normal_opacity=normal_opacity, xy=xy, depth=depth,
                tiles_touched=tiles_touched, rect=rect, num_rendered=D,
                keys=keys[:D], ids=ids[:D], ranges=ranges, final_T=final_T,
                n_contrib=n_contrib,
                _inputs=dict(means3D=means3D, opacities=opacities, scales=scales,
                             rotations=rotations, colors=colors, vm=vm, pm=pm, bg=bg,
                             H=H, W=W, scale_modifier=float(scale_modifier)))


def rasterize_backward(fwd, grad_color, grad_allmap):
    """Backward pass of `rasterize`; returns grads for the five differentiable inputs."""
    L = lib()
    i = fwd["_inputs"]
    P, H, W = i["means3D"].shape[0], i["H"], i["W"]
    grad_color, grad_allmap = _f32(grad_color), _f32(grad_allmap)
    g_T = np.zeros((P, 9), np.float64)
    g_m2 = np.zeros((P, 2), np.float64)
    g_n = np.zeros((P, 3), np.float64)
    g_op = np.zeros(P, np.float64)
    g_col = np.zeros((P, 3), np.float64)
    ids = fwd["ids"] if fwd["ids"].size else np.zeros(1, np.uint32)
    L.so_render_backward(C.c_int(H), C.c_int(W), _p(fwd["ranges"], C.c_int32), _p(ids, C.c_uint32),
                         _p(fwd["xy"], C.c_float), _p(fwd["transmat"], C.c_float),
                         _p(fwd["normal_opacity"], C.c_float), _p(i["colors"], C.c_float),
                         _p(i["bg"], C.c_float), _p(fwd["final_T"], C.c_float),
                         _p(fwd["n_contrib"], C.c_int32),
                         _p(grad_color, C.c_float), _p(grad_allmap, C.c_float),
                         _p(g_T, C.c_double), _p(g_m2, C.c_double), _p(g_n, C.c_double),
                         _p(g_op, C.c_double), _p(g_col, C.c_double))
    g_means = np.zeros((P, 3), np.float64)
    g_scales = np.zeros((P, 2), np.float64)
    g_rots = np.zeros((P, 4), np.float64)
    L.so_preprocess_backward(C.c_int(P), _p(i["means3D"], C.c_float), _p(i["scales"], C.c_float),
                             _p(i["rotations"], C.c_float), _p(i["vm"], C.c_float),
                             _p(i["pm"], C.c_float), C.c_int(H), C.c_int(W),
                             C.c_float(i["scale_modifier"]), _p(fwd["radii"], C.c_int),
                             _p(fwd["transmat"], C.c_float),
                             _p(g_T, C.c_double), _p(g_m2, C.c_double), _p(g_n, C.c_double),
                             _p(g_means, C.c_double), _p(g_scales, C.c_double), _p(g_rots, C.c_double))
    return dict(means3D=g_means, opacities=g_op.reshape(P, 1), scales=g_scales,
                rotations=g_rots, colors=g_col,
                dL_dtransmat=g_T, dL_dmean2D=g_m2, dL_dnormal=g_n)


# camera + synthetic-scene builders live in tools/synth.py (numpy only) so that bench.py's GPU arm can build its
# inputs without mapping this library; re-exported here for the tests that reach them through the oracle module
from tools.synth import camera_from_pose25, orbit_pose25, synthetic_surfels  # noqa: E402,F401
