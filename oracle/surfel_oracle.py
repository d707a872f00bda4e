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


# --------------------------------------------------------------------------
# camera + synthetic-scene helpers (SURVEY.md section 8d input spec)
# --------------------------------------------------------------------------

def camera_from_pose25(pose, znear=0.01, zfar=100.0):
    """Restates FlowMatchingEngine.c_to_3dgs_format
    (/root/reference/nsr/lsgm/flow_matching_trainer.py:2174-2228) with
    getWorld2View2 / getProjectionMatrix
    (/root/reference/utils/gs_utils/graphics_utils.py:38-85).
    pose: 25 floats = c2w(16, row major) + K(9, normalised).  Returns
    (cam_view[4,4], cam_view_proj[4,4], cam_pos[3], tanfov) in the reference's
    row-vector (transposed) layout, float32."""
    pose = np.asarray(pose, dtype=np.float32)
    c2w = pose[:16].reshape(4, 4)
    w2c = np.linalg.inv(c2w)
    R = np.transpose(w2c[:3, :3])
    T = w2c[:3, 3]
    fx = float(pose[16])
    fov = 2.0 * np.arctan(1.0 / (2.0 * fx))            # focal2fov(fx, 1)
    tanfov = float(np.tan(fov * 0.5))
    Rt = np.zeros((4, 4))
    Rt[:3, :3] = R.transpose()
    Rt[:3, 3] = T
    Rt[3, 3] = 1.0
    Rt = np.linalg.inv(np.linalg.inv(Rt))               # getWorld2View2 with trans=0, scale=1
    world_view = np.float32(Rt).T
    th = np.tan(fov / 2.0)
    top = th * znear
    right = th * znear
    Pm = np.zeros((4, 4), np.float32)
    Pm[0, 0] = 2.0 * znear / (2 * right)
    Pm[1, 1] = 2.0 * znear / (2 * top)
    Pm[3, 2] = 1.0
    Pm[2, 2] = zfar / (zfar - znear)
    Pm[2, 3] = -(zfar * znear) / (zfar - znear)
    full = (world_view.astype(np.float32) @ Pm.T).astype(np.float32)
    cam_pos = np.linalg.inv(world_view)[3, :3].astype(np.float32)
    return world_view.astype(np.float32), full, cam_pos, tanfov


def orbit_pose25(azim_deg, elev_deg, radius=1.8, fx=1.3889):
    """A look-at-origin camera in the same 25-float layout as
    /root/reference/assets/objv_eval_pose.pt (c2w row-major + normalised K);
    used when that fixture is not on disk (GPU box)."""
    az, el = np.deg2rad(azim_deg), np.deg2rad(elev_deg)
    eye = radius * np.array([np.cos(el) * np.cos(az), np.cos(el) * np.sin(az), np.sin(el)])
    fwd = -eye / np.linalg.norm(eye)
    up = np.array([0.0, 0.0, 1.0])
    right = np.cross(fwd, up)
    right /= np.linalg.norm(right)
    down = np.cross(fwd, right)
    c2w = np.eye(4)
    c2w[:3, 0], c2w[:3, 1], c2w[:3, 2], c2w[:3, 3] = right, down, fwd, eye
    K = np.array([fx, 0, 0.5, 0, fx, 0.5, 0, 0, 1.0])
    return np.concatenate([c2w.reshape(-1), K]).astype(np.float32)


def synthetic_surfels(P, seed=0, scale_boost=1.0):
    """[P,13] surfels per SURVEY.md 8(d): xyz U(-0.45,0.45)^3, opacity sigmoid(N),
    scales softplus(N(-2.5,1))*0.0045/ln2 clamped to [1e-4,0.05], unit quats,
    rgb 0.5*tanh(N)+0.5 (activations of /root/reference/vit/vit_triplane.py:1289-1313)."""
    rng = np.random.default_rng(seed)
    xyz = rng.uniform(-0.45, 0.45, (P, 3))
    op = 1.0 / (1.0 + np.exp(-rng.standard_normal((P, 1))))
    sc = np.log1p(np.exp(rng.standard_normal((P, 2)) - 2.5)) * (0.0045 / np.log(2.0)) * scale_boost
    sc = np.clip(sc, 1e-4, 0.05)
    q = rng.standard_normal((P, 4))
    q /= np.linalg.norm(q, axis=1, keepdims=True)
    rgb = 0.5 * np.tanh(rng.standard_normal((P, 3))) + 0.5
    return np.concatenate([xyz, op, sc, q, rgb], 1).astype(np.float32)
