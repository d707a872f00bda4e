"""The GPU comparison baseline of bench.py (baseline/raster_standin.cu: the upstream flow restated literally --
per-view launches, global radix sort, per-pair atomics) must compute the same thing as the oracle, or the
`gpu_standin` ratio in the bench line would compare different work."""
import numpy as np
import pytest
import torch

from tests.helpers import cameras, oracle_view, rel_l2, scene

pytestmark = pytest.mark.gpu


def test_standin_matches_oracle_forward_and_backward():
    from baseline.raster_standin import StandinRasterizer
    from oracle import surfel_oracle as so
    P, H, W, V = 4000, 160, 144, 2
    g = scene(P, 90, 6.0)
    vs, ps, _, _ = cameras(V, start=1)
    bg = [1.0, 0.5, 0.2]
    dev = torch.device("cuda:0")
    r = StandinRasterizer(P, H, W, V)
    g13 = torch.tensor(g, device=dev)
    color, allmap, radii, nr = r.forward(g13, torch.tensor(vs, device=dev), torch.tensor(ps, device=dev), torch.tensor(bg, device=dev))
    rng = np.random.default_rng(0)
    gc = rng.standard_normal((V, 3, H, W)).astype(np.float32)
    ga = rng.standard_normal((V, 7, H, W)).astype(np.float32)
    grad = r.backward(torch.tensor(gc, device=dev), torch.tensor(ga, device=dev)).cpu().numpy()
    want = np.zeros((P, 13))
    for v in range(V):
        o = oracle_view(g, vs[v], ps[v], bg, H, W)
        assert nr[v] == o["num_rendered"]
        assert np.array_equal(radii[v].cpu().numpy(), o["radii"])
        assert rel_l2(color[v].cpu().numpy(), o["color"]) <= 1e-3
        for ch in (0, 1, 2, 3, 4):
            assert rel_l2(allmap[v, ch].cpu().numpy(), o["allmap"][ch]) <= 1e-3, ch
        b = so.rasterize_backward(o, gc[v], ga[v])
        want[:, 0:3] += b["means3D"]; want[:, 3:4] += b["opacities"]; want[:, 4:6] += b["scales"]
        want[:, 6:10] += b["rotations"]; want[:, 10:13] += b["colors"]
    for sl in (slice(0, 3), slice(3, 4), slice(4, 6), slice(6, 10), slice(10, 13)):
        assert rel_l2(grad[:, sl], want[:, sl]) <= 2e-3, sl
