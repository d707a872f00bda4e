"""Shared helpers for the parity tests (scene builders + metrics)."""
import numpy as np

from tools import synth as so          # scene / camera builders: numpy only (the oracle is imported lazily below)


def rel_l2(a, b):
    a = np.asarray(a, dtype=np.float64)
    b = np.asarray(b, dtype=np.float64)
    return float(np.linalg.norm(a - b) / max(np.linalg.norm(b), 1e-30))


def cameras(n, start=0):
    """n look-at cameras in the reference's layout -> (view[n,4,4], proj[n,4,4], pos[n,3], tanfov)."""
    vs, ps, cs = [], [], []
    tf = None
    for k in range(n):
        pose = so.orbit_pose25(30.0 + 47.0 * (k + start), 20.0 + 11.0 * ((k + start) % 4) - 15.0)
        v, p, c, tf = so.camera_from_pose25(pose)
        vs.append(v); ps.append(p); cs.append(c)
    # camera_from_pose25 returns transposed VIEWS (Fortran order); hand out C-contiguous arrays
    return (np.ascontiguousarray(np.stack(vs)), np.ascontiguousarray(np.stack(ps)), np.ascontiguousarray(np.stack(cs)), tf)


def scene(P, seed, scale_boost=1.0, smin=None, smax=None):
    g = so.synthetic_surfels(P, seed, scale_boost=scale_boost)
    if smin is not None:
        g[:, 4:6] = np.clip(g[:, 4:6], smin, smax)
    return g


def oracle_view(g, view, proj, bg, H, W, scale_modifier=1.0):
    from oracle import surfel_oracle as so
    return so.rasterize(g[:, 0:3], g[:, 3:4], g[:, 4:6], g[:, 6:10], g[:, 10:13], view, proj, bg, H, W,
                        scale_modifier)
