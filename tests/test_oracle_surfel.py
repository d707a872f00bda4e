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
