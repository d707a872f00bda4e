"""Generates tests/golden/surfel_small.npz from the CPU oracle (regression pin).

The reference holds no golden vectors for the rasteriser and its
native dependency is not installable here, so these vectors come from
oracle/surfel_oracle.c itself; they freeze its behaviour and give the GPU tests
a fixture that needs no oracle build.  Run from the repo root:
    python tests/golden/make_golden.py
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import surfel_oracle as so  # noqa: E402
from tests.helpers import cameras, scene  # noqa: E402

P, H, W = 1500, 80, 96
g = scene(P, 77, 12.0)
vs, ps, _, _ = cameras(1, start=2)
bg = np.array([1.0, 0.5, 0.2], np.float32)
o = so.rasterize(g[:, 0:3], g[:, 3:4], g[:, 4:6], g[:, 6:10], g[:, 10:13], vs[0], ps[0], bg, H, W)
rng = np.random.default_rng(7)
gc = rng.standard_normal((3, H, W)).astype(np.float32)
ga = rng.standard_normal((7, H, W)).astype(np.float32)
b = so.rasterize_backward(o, gc, ga)
np.savez_compressed(os.path.join(ROOT, "tests", "golden", "surfel_small.npz"),
                    g=g, view=vs[0], proj=ps[0], bg=bg, H=H, W=W, gc=gc, ga=ga,
                    color=o["color"], allmap=o["allmap"], radii=o["radii"], ids=o["ids"], ranges=o["ranges"],
                    n_contrib=o["n_contrib"],
                    **{"grad_" + k: b[k].astype(np.float32) for k in ("means3D", "opacities", "scales", "rotations", "colors")})
print("wrote surfel_small.npz, D =", o["num_rendered"])
