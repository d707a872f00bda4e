"""CPU tests of the oracle itself: the C restatement against (a) fp64 autograd
of the independent torch restatement, (b) the committed golden vectors, and
(c) structural invariants of the binning.  (PARITY UNPINNED: the reference has
no golden vectors for this path; the goldens are regression pins produced by
tests/golden/make_golden.py from this same oracle.)"""
import os

import numpy as np
import pytest
import torch

from oracle import surfel_oracle as so
from oracle import surfel_torch as st
from tests.helpers import cameras, oracle_view, rel_l2, scene

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


@pytest.mark.parametrize("P,H,W,boost,seed", [(150, 40, 48, 40.0, 3), (60, 32, 32, 80.0, 4), (300, 48, 64, 15.0, 5)])
def test_c_oracle_matches_torch_autograd(P, H, W, boost, seed):
    g = scene(P, seed, boost, 0.004, 0.09)
    vs, ps, _, _ = cameras(1, start=seed)
    bg = [1.0, 0.5, 0.2]
    o = oracle_view(g, vs[0], ps[0], bg, H, W)
    T = lambda a: torch.tensor(a, dtype=torch.float64, requires_grad=True)
    m, op, sc, ro, co = T(g[:, 0:3]), T(g[:, 3:4]), T(g[:, 4:6]), T(g[:, 6:10]), T(g[:, 10:13])
    color, radii, allmap = st.rasterize(m, op, sc, ro, co, torch.tensor(vs[0], dtype=torch.float64),
                                        torch.tensor(ps[0], dtype=torch.float64), torch.tensor(bg), H, W)
    assert np.array_equal(radii.numpy(), o["radii"])
    assert rel_l2(o["color"], color.detach().numpy()) < 1e-5
    for c in range(7):
        assert rel_l2(o["allmap"][c], allmap[c].detach().numpy()) < (2e-3 if c == 6 else 1e-4), c
    rng = np.random.default_rng(seed)
    gc, ga = rng.standard_normal((3, H, W)), rng.standard_normal((7, H, W))
    ((color * torch.tensor(gc)).sum() + (allmap * torch.tensor(ga)).sum()).backward()
    b = so.rasterize_backward(o, gc, ga)
    for k, t in [("means3D", m), ("opacities", op), ("scales", sc), ("rotations", ro), ("colors", co)]:
        assert rel_l2(b[k], t.grad.numpy()) < 2e-4, k


def test_binning_invariants():
    g = scene(5000, 9, 5.0)
    vs, ps, _, _ = cameras(1)
    o = oracle_view(g, vs[0], ps[0], [1, 1, 1], 200, 168)
    D = o["num_rendered"]
    assert D == int(o["tiles_touched"].sum()) and D > 0
    assert np.all(o["keys"][1:] >= o["keys"][:-1])
    r = o["ranges"]
    assert int((r[:, 1] - r[:, 0]).sum()) == D
    tiles = (o["keys"] >> np.uint64(32)).astype(np.int64)
    for t in np.unique(tiles):
        assert np.all(tiles[r[t, 0]:r[t, 1]] == t)
    vis = o["radii"] > 0
    assert np.array_equal(vis, o["tiles_touched"] > 0)


def test_empty_and_degenerate_inputs():
    vs, ps, _, _ = cameras(1)
    g = scene(8, 1)
    g[:, 0:3] += 50.0
    o = oracle_view(g, vs[0], ps[0], [0.3, 0.6, 0.9], 32, 40)
    assert o["num_rendered"] == 0 and np.all(o["radii"] == 0)
    assert np.allclose(o["color"][0], 0.3) and np.allclose(o["allmap"], 0)
    b = so.rasterize_backward(o, np.ones((3, 32, 40)), np.ones((7, 32, 40)))
    assert all(np.all(b[k] == 0) for k in ("means3D", "opacities", "scales", "rotations", "colors"))


def test_golden_vectors():
    z = np.load(os.path.join(GOLD, "surfel_small.npz"))
    o = so.rasterize(z["g"][:, 0:3], z["g"][:, 3:4], z["g"][:, 4:6], z["g"][:, 6:10], z["g"][:, 10:13],
                     z["view"], z["proj"], z["bg"], int(z["H"]), int(z["W"]))
    assert np.array_equal(o["radii"], z["radii"])
    assert np.array_equal(o["ids"], z["ids"]) and np.array_equal(o["ranges"], z["ranges"])
    assert rel_l2(o["color"], z["color"]) < 1e-6 and rel_l2(o["allmap"], z["allmap"]) < 1e-5
    b = so.rasterize_backward(o, z["gc"], z["ga"])
    for k in ("means3D", "opacities", "scales", "rotations", "colors"):
        assert rel_l2(b[k], z["grad_" + k]) < 1e-5, k


def test_camera_helper_matches_reference_fixture_layout():
    # objv_eval_pose.pt row 0 of the reference (copied as literals: the file is not on the GPU box)
    pose = np.array([-2.3784e-08, 2.3736e-01, -9.7142e-01, 1.7213e+00, 1.0, 0.0, -2.3784e-08, 0.0,
                     0.0, -9.7142e-01, -2.3736e-01, 4.2059e-01, 0, 0, 0, 1.0,
                     1.3889, 0, 0.5, 0, 1.3889, 0.5, 0, 0, 3.9062e-03], np.float32)
    view, proj, pos, tf = so.camera_from_pose25(pose)
    assert abs(tf - 0.36) < 1e-4                      # tanfov = 0.5/1.3889
    assert np.allclose(pos, pose[[3, 7, 11]], atol=1e-3)
    # row-vector convention: [p,1] @ view has z = distance along the optical axis
    z = (np.array([0, 0, 0, 1.0]) @ view)[2]
    assert abs(z - np.linalg.norm(pose[[3, 7, 11]])) < 1e-3


# ---- known-answer tests: the oracle against closed-form geometry written down independently of it -------------------
def _ka_camera(H, W, F=1.2):
    from tools import synth
    pose = np.concatenate([np.eye(4).reshape(-1), np.array([F, 0, .5, 0, F, .5, 0, 0, 1.])]).astype(np.float32)
    view, proj, _, tanfov = synth.camera_from_pose25(pose)           # camera at the origin looking down +z
    py, px = np.mgrid[0:H, 0:W].astype(np.float64)
    return view, proj, tanfov, px, py, ((2 * px + 1) / W - 1) * tanfov, ((2 * py + 1) / H - 1) * tanfov


def _ka_rot(q):
    w, x, y, z = np.asarray(q, float) / np.linalg.norm(q)
    return np.array([[1 - 2 * (y * y + z * z), 2 * (x * y - w * z), 2 * (x * z + w * y)],
                     [2 * (x * y + w * z), 1 - 2 * (x * x + z * z), 2 * (y * z - w * x)],
                     [2 * (x * z - w * y), 2 * (y * z + w * x), 1 - 2 * (x * x + y * y)]])


def _ka_surfel(P0, s, q, o, cam, H, W):
    return _ka_surfel_frame(P0, s, _ka_rot(q), o, cam, H, W)


def _ka_surfel_frame(P0, s, R, o, cam, H, W):
    """Ray / plane intersection of every pixel's ray with the surfel's plane, local (u, v) in units of the scales, the
    2DGS low-pass (rho2d = 2 |pixel - box centre|^2), alpha = min(0.99, o exp(-min(rho3d, rho2d) / 2)) dropped below
    1/255, depth = hit depth (centre depth where the low-pass wins), normal turned towards the camera."""
    _, _, tanfov, px, py, dx, dy = cam
    tu, tv, n = R[:, 0], R[:, 1], R[:, 2]
    d = np.stack([dx, dy, np.ones_like(dx)], -1)
    t = (P0 @ n) / (d @ n)
    h = d * t[..., None]
    u, v = ((h - P0) @ tu) / s[0], ((h - P0) @ tv) / s[1]
    rho3 = u * u + v * v
    # upstream centres its screen-space low-pass on the centre of the bounding box of the projected 3-sigma ellipse
    # (compute_aabb's `center`, not the projected splat centre): here the extremes of the projected rim, located on a
    # 2 000-point sweep and polished by bounded scalar minimisation (no use of the oracle's quadric formula)
    from scipy.optimize import minimize_scalar

    def rim_pix(phi, axis):
        r = P0 + 3.0 * (np.cos(phi) * s[0] * tu + np.sin(phi) * s[1] * tv)
        return ((r[axis] / r[2] / tanfov + 1) * (W if axis == 0 else H) - 1) * 0.5

    def extreme(axis, sign):
        grid = np.linspace(0.0, 2 * np.pi, 2001)[:-1]
        k = int(np.argmin([sign * rim_pix(p, axis) for p in grid]))
        step = grid[1] - grid[0]
        r = minimize_scalar(lambda p: sign * rim_pix(p, axis), bounds=(grid[k] - step, grid[k] + step), method="bounded",
                            options={"xatol": 1e-13})
        return rim_pix(r.x, axis)
    cx = 0.5 * (extreme(0, 1.0) + extreme(0, -1.0))
    cy = 0.5 * (extreme(1, 1.0) + extreme(1, -1.0))
    rho2 = 2.0 * ((px - cx) ** 2 + (py - cy) ** 2)
    depth = np.where(rho3 <= rho2, h[..., 2], P0[2])
    a = np.minimum(0.99, o * np.exp(-0.5 * np.minimum(rho3, rho2)))
    a = np.where((a < 1.0 / 255.0) | (depth < 0.2), 0.0, a)
    return a, depth, (n if (P0 @ n) < 0 else -n)


def _ka_run(surfels, cam, H, W, bg):
    P0 = np.array([s[0] for s in surfels], np.float32)
    sc = np.array([s[1] for s in surfels], np.float32)
    q = np.array([s[2] for s in surfels], np.float32)
    o = np.array([s[3] for s in surfels], np.float32)
    col = np.array([s[4] for s in surfels], np.float32)
    return so.rasterize(P0, o, sc, q, col, cam[0], cam[1], bg, H, W)


def test_known_answer_single_tilted_surfel():
    """Alpha, colour, expected depth, median depth and the camera-facing normal of one surfel tilted 35 degrees: the
    homography-based intersection of the oracle against plain ray / plane geometry."""
    H = W = 64
    cam = _ka_camera(H, W)
    bg = np.array([0.1, 0.2, 0.3], np.float32)
    th = np.deg2rad(35.0)
    for q in ([1, 0, 0, 0], [np.cos(th / 2), 0, np.sin(th / 2), 0], [np.cos(th / 2), np.sin(th / 2) * 0.6, 0, np.sin(th / 2) * 0.8]):
        P0, s, o, col = np.array([0.05, -0.03, 2.0]), (0.15, 0.25), 0.7, np.array([0.2, 0.8, 0.5])
        out = _ka_run([(P0, s, q, o, col)], cam, H, W, bg)
        a, depth, n = _ka_surfel(P0, s, q, o, cam, H, W)
        am = out["allmap"]
        assert (a > 0).mean() > 0.05
        assert np.abs(am[1] - a).max() < 2e-6
        assert np.abs(out["color"] - (col[:, None, None] * a + bg[:, None, None] * (1 - a))).max() < 2e-6
        assert np.abs(am[0] - a * depth).max() < 5e-6
        assert np.abs(am[2:5] - a[None] * n[:, None, None]).max() < 2e-6
        assert np.abs(am[5] - np.where(a > 0, depth, 0.0)).max() < 5e-6          # median depth: T = 1 > 0.5 at the only hit
        assert np.abs(am[6]).max() == 0.0                                        # one layer has no distortion
        # radius = ceil(3 sigma of the larger projected axis) for the fronto-parallel case
        if q == [1, 0, 0, 0]:
            f_px = W / (2 * cam[2])
            assert int(out["radii"][0]) == int(np.ceil(3 * max(s) * f_px / P0[2]))


def test_known_answer_two_layers_compositing_and_distortion():
    """Two fronto-parallel surfels at different depths: front-to-back compositing of colour / depth / alpha, the median
    depth rule (last hit whose incoming transmittance is above 0.5) and the distortion term
    a1 a2 (1 - a1) (m2 - m1)^2 with m = far / (far - near) (1 - near / z), near 0.2, far 100."""
    H = W = 48
    cam = _ka_camera(H, W)
    bg = np.array([0.05, 0.1, 0.0], np.float32)
    A = (np.array([0.02, 0.01, 1.5]), (0.12, 0.12), [1, 0, 0, 0], 0.6, np.array([0.9, 0.1, 0.2]))
    B = (np.array([-0.04, 0.03, 2.5]), (0.3, 0.2), [1, 0, 0, 0], 0.9, np.array([0.1, 0.7, 0.9]))
    out = _ka_run([B, A], cam, H, W, bg)                     # given back to front: the oracle has to sort them
    a1, z1, _ = _ka_surfel(A[0], A[1], A[2], A[3], cam, H, W)
    a2, z2, _ = _ka_surfel(B[0], B[1], B[2], B[3], cam, H, W)
    T1 = 1 - a1
    am = out["allmap"]
    col = A[4][:, None, None] * a1 + B[4][:, None, None] * a2 * T1 + bg[:, None, None] * T1 * (1 - a2)
    assert np.abs(out["color"] - col).max() < 3e-6
    assert np.abs(am[1] - (a1 + a2 * T1)).max() < 3e-6
    assert np.abs(am[0] - (a1 * z1 + a2 * T1 * z2)).max() < 1e-5
    median = np.where((a2 > 0) & (T1 > 0.5), z2, np.where(a1 > 0, z1, 0.0))
    assert np.abs(am[5] - median).max() < 1e-5
    m = lambda z: 100.0 / (100.0 - 0.2) * (1 - 0.2 / z)
    dist = a1 * a2 * T1 * (m(z2) - m(z1)) ** 2
    assert np.abs(am[6] - dist).max() < 3e-6 and dist.max() > 1e-4


def test_known_answer_orbit_camera_conventions():
    """The same closed form under a look-at camera in the reference's 25-float pose layout (c2w + normalised K): the
    row-vector view / projection matrices, the view-space normal and depth are what a plain w2c transform gives."""
    from tools import synth
    H = W = 64
    pose = synth.orbit_pose25(40.0, 25.0, radius=2.2, fx=1.2)
    view, proj, _, tanfov = synth.camera_from_pose25(pose)
    w2c = np.linalg.inv(pose[:16].reshape(4, 4).astype(np.float64))
    py, px = np.mgrid[0:H, 0:W].astype(np.float64)
    cam = (view, proj, tanfov, px, py, ((2 * px + 1) / W - 1) * tanfov, ((2 * py + 1) / H - 1) * tanfov)
    th = np.deg2rad(50.0)
    q = np.array([np.cos(th / 2), 0.3 * np.sin(th / 2), np.sqrt(1 - 0.09 - 0.16) * np.sin(th / 2), 0.4 * np.sin(th / 2)])
    Pw, s, o, col = np.array([0.08, -0.05, 0.1]), (0.2, 0.12), 0.85, np.array([0.6, 0.4, 0.9])
    bg = np.zeros(3, np.float32)
    out = so.rasterize(Pw[None].astype(np.float32), np.array([o], np.float32), np.array([s], np.float32),
                       q[None].astype(np.float32), col[None].astype(np.float32), view, proj, bg, H, W)
    # camera-space surfel: centre and frame through w2c, then the camera-at-origin closed form
    Pc = w2c[:3, :3] @ Pw + w2c[:3, 3]
    Rc = w2c[:3, :3] @ _ka_rot(q)
    a, depth, n = _ka_surfel_frame(Pc, s, Rc, o, cam, H, W)
    am = out["allmap"]
    assert (a > 0).mean() > 0.03 and Pc[2] > 1.5
    assert np.abs(am[1] - a).max() < 3e-6
    assert np.abs(am[0] - a * depth).max() < 1e-5
    assert np.abs(am[2:5] - a[None] * n[:, None, None]).max() < 3e-6
    assert np.abs(out["color"] - col[:, None, None] * a).max() < 3e-6


def _ka_render(params, cam, H, W, bg):
    """Closed-form image of K surfels (params [K, 13] = xyz, opacity, scales, quaternion, rgb; float64): every surfel's
    alpha / depth / normal field from _ka_surfel_frame, composited front to back in the order of the centres' depths,
    with the seven auxiliary channels as the reference consumes them (nsr/gs_surfel.py:121-163: [0] sum w z, [1] sum w,
    [2:5] sum w n, [5] median depth, [6] distortion)."""
    order = np.argsort(params[:, 2], kind="stable")
    T = np.ones((H, W))
    color = np.zeros((3, H, W))
    am = np.zeros((7, H, W))
    A = np.zeros((H, W)); M1 = np.zeros((H, W)); M2 = np.zeros((H, W))
    for k in order:
        p = params[k]
        a, z, n = _ka_surfel_frame(p[0:3], p[4:6], _ka_rot(p[6:10]), p[3], cam, H, W)
        w = a * T
        m = 100.0 / (100.0 - 0.2) * (1 - 0.2 / np.where(z > 0, z, 1.0))
        am[6] += (m * m * A + M2 - 2 * m * M1) * w
        A += w; M1 += m * w; M2 += m * m * w
        color += p[10:13][:, None, None] * w
        am[0] += w * z
        am[1] += w
        am[2:5] += n[:, None, None] * w
        am[5] = np.where((a > 0) & (T > 0.5), z, am[5])
        T = T * (1 - a)
    return color + bg[:, None, None] * T, am


def test_known_answer_gradients_by_finite_differences():
    """Analytic backward of the oracle against central differences (float64) of the independent closed form above:
    three surfels -- a tilted one, a larger one behind it, and one that is smaller than a pixel so that the
    screen-space low-pass and its bounding-box centre carry the gradient -- all 13 parameters of each.  The quaternion
    gradient is compared in both variants (1: chained through the normalisation = the derivative of the closed form;
    0: upstream's vjp at q/|q|, equal to it after projecting out the radial component)."""
    H = W = 40
    cam = _ka_camera(H, W)
    bg = np.array([0.2, 0.1, 0.3])
    th1, th2, th3 = np.deg2rad(30.0), np.deg2rad(-20.0), np.deg2rad(40.0)
    params = np.array([
        [0.03, -0.02, 1.6, 0.75, 0.16, 0.10, np.cos(th1 / 2), 0.0, np.sin(th1 / 2), 0.0, 0.9, 0.2, 0.1],
        [-0.05, 0.04, 2.4, 0.90, 0.30, 0.22, np.cos(th2 / 2), np.sin(th2 / 2) * 0.6, np.sin(th2 / 2) * 0.8, 0.0, 0.1, 0.8, 0.6],
        [0.10, 0.12, 2.0, 0.95, 0.008, 0.011, np.cos(th3 / 2), np.sin(th3 / 2) * 0.8, np.sin(th3 / 2) * 0.6, 0.0, 0.5, 0.5, 1.0]], np.float64)
    rng = np.random.default_rng(5)
    wc = rng.standard_normal((3, H, W))
    wa = rng.standard_normal((7, H, W))
    wa[5] = 0.0                                          # the median depth is piecewise constant in every parameter

    def loss(p):
        c, am = _ka_render(p, cam, H, W, bg)
        return float((wc * c).sum() + (wa * am).sum())

    fd = np.zeros_like(params)
    for i in range(params.shape[0]):
        for j in range(13):
            e = np.zeros_like(params)
            e[i, j] = 1e-6
            fd[i, j] = (loss(params + e) - loss(params - e)) / 2e-6

    def oracle_grads(variant):
        rf, qn = so.get_variant()
        so.set_variant(rf, variant)
        try:
            f32 = params.astype(np.float32)
            fwd = so.rasterize(f32[:, 0:3], f32[:, 3], f32[:, 4:6], f32[:, 6:10], f32[:, 10:13], cam[0], cam[1],
                               bg.astype(np.float32), H, W)
            c, am = _ka_render(params, cam, H, W, bg)
            assert np.abs(fwd["color"] - c).max() < 5e-6 and np.abs(fwd["allmap"][[0, 1, 2, 3, 4, 6]] - am[[0, 1, 2, 3, 4, 6]]).max() < 1e-5
            return so.rasterize_backward(fwd, wc.astype(np.float32), wa.astype(np.float32))
        finally:
            so.set_variant(rf, qn)

    TOL = 1e-4                                           # measured: 1e-7 .. 3e-6

    def relerr(a, b):
        return float(np.linalg.norm(a - b) / max(np.linalg.norm(b), 1e-30))

    g1 = oracle_grads(1)
    for name, sl in (("means3D", slice(0, 3)), ("opacities", slice(3, 4)), ("scales", slice(4, 6)), ("rotations", slice(6, 10)),
                     ("colors", slice(10, 13))):
        for i in range(params.shape[0]):
            assert relerr(g1[name][i], fd[i, sl]) < TOL, (name, i, g1[name][i], fd[i, sl])
    g0 = oracle_grads(0)
    for i in range(params.shape[0]):
        q = params[i, 6:10]                              # unit quaternions: the two variants differ by the radial part only
        assert relerr(g0["rotations"][i] - q * (q @ g0["rotations"][i]), fd[i, 6:10]) < TOL
        assert relerr(g0["means3D"][i], fd[i, 0:3]) < TOL
    # the sub-pixel surfel really is in the low-pass regime: its position gradient is not negligible
    assert np.linalg.norm(fd[2, 0:3]) > 1e-2
