"""CPU estimate (from the oracle's tile lists of the C2 scene) of how many evaluation rounds a warp needs per chunk
under different lane mappings of the composite kernels -- the numbers quoted in csrc/raster_render.cu:
  forward : one surfel per warp round (GS=32) | two 4x4 half-warp groups (GS=16) | four 4x2 groups (GS=8) | one list per lane
  backward: per-lane lists synchronised every 32 surfels (round 1) | per-lane lists over a whole group of 128
    python tools/raster_rounds.py [tiles]"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from tests.helpers import cameras, oracle_view, scene  # noqa: E402


def main():
    ntiles = int(sys.argv[1]) if len(sys.argv) > 1 else 200
    P, H, W = 100000, 512, 512
    g = scene(P, 40)
    vs, ps, _, _ = cameras(6)
    o = oracle_view(g, vs[0], ps[0], [1, 1, 1], H, W)
    xy, opa = o["xy"], g[:, 3]
    tau = 2 * np.log(np.maximum(255 * opa, 1.0)) * 1.001 + 1e-3
    r2 = np.sqrt(0.5 * tau) + 0.51                      # the low-pass disc of K1's cull box (the 3-D part is tiny here)
    bx0, bx1, by0, by1 = xy[:, 0] - r2, xy[:, 0] + r2, xy[:, 1] - r2, xy[:, 1] + r2
    ids, rng = o["ids"], o["ranges"]
    tot = dict(gs32=0, gs16=0, gs8=0, lane=0, lane_sub32=0, lane_grp128=0, useful=0)
    tiles = np.random.default_rng(0).choice(len(rng), ntiles, replace=False)
    nw = 0
    for t in tiles:
        a, b = rng[t]
        if b <= a:
            continue
        sid = ids[a:b].astype(int)
        ty, tx = divmod(int(t), 32)
        ox, oy = tx * 16, ty * 16
        X0, X1, Y0, Y1 = bx0[sid], bx1[sid], by0[sid], by1[sid]
        valid = opa[sid] * 255 >= 1

        def hits(x0, y0, w, h):
            return valid & ~((X1 < ox + x0) | (X0 > ox + x0 + w - 1) | (Y1 < oy + y0) | (Y0 > oy + y0 + h - 1))
        for wp in range(8):
            nw += 1
            lx0, ly0 = (wp & 1) * 8, (wp >> 1) * 4
            tot["gs32"] += hits(lx0, ly0, 8, 4).sum()
            tot["gs16"] += max(hits(lx0, ly0, 4, 4).sum(), hits(lx0 + 4, ly0, 4, 4).sum())
            tot["gs8"] += max(hits(lx0 + 4 * (k & 1), ly0 + 2 * (k >> 1), 4, 2).sum() for k in range(4))
            pl = np.stack([hits(lx0 + (l & 7), ly0 + (l >> 3), 1, 1) for l in range(32)])          # [32 lanes, n]
            tot["lane"] += pl.sum(1).max()
            tot["useful"] += pl.sum()
            n = pl.shape[1]
            tot["lane_sub32"] += sum(pl[:, s:s + 32].sum(1).max() for s in range(0, n, 32))
            tot["lane_grp128"] += sum(pl[:, s:s + 128].sum(1).max() for s in range(0, n, 128))
    print("rounds per warp and chunk (C2 scene, %d tiles): " % ntiles + ", ".join("%s %.1f" % (k, v / nw) for k, v in tot.items()))


if __name__ == "__main__":
    main()
