"""Independent, differentiable (autograd) restatement of the surfel rasteriser.

TEST INFRASTRUCTURE ONLY.  PARITY UNPINNED (see oracle/surfel_oracle.c).

Purpose: cross-check the analytic backward of the C oracle (and therefore of
the CUDA kernels) against fp64 autograd of a second, vectorised statement of
the same forward.  All pixels x all surfels are evaluated densely in global
depth order; tile-list membership (the surfel's tile rectangle covers the
pixel's tile) is applied as a mask, which is equivalent to walking the pixel's
per-tile sorted list.  Small scenes only (memory is O(H*W*P)).

Upstream quirks reproduced with detach tricks so autograd == upstream's
analytic backward (github.com/hbb1/diff-surfel-rasterization backward.cu):
  * alpha = min(0.99, opacity*G): gradient passes straight through the clamp;
  * the dual-visible normal flip sign and the quaternion normalisation factor
    are constants;
  * all skip / stop decisions are non-differentiable masks.
Follows the call contract of /root/reference/nsr/gs_surfel.py:85-142.
"""
import torch

NEAR_N, FAR_N = 0.2, 100.0
FILTER_SIZE, FILTER_INV_SQUARE = 0.707106, 2.0


RADIUS_FORMULA = 0        # see oracle/surfel_oracle.c: 0 = ceil(max(ex, ey, 3 f)), 1 = ceil(3 max(ex, ey, f))
QUAT_NORM_GRAD = 0        # 0 = normalisation factor is a constant (detached), 1 = autograd chains through it


def _quat_to_R(q):
    s = 1.0 / q.norm(dim=-1, keepdim=True)
    if not QUAT_NORM_GRAD:
        s = s.detach()
    w, x, y, z = (q * s).unbind(-1)
    R = torch.stack([
        1 - 2 * (y * y + z * z), 2 * (x * y - w * z), 2 * (x * z + w * y),
        2 * (x * y + w * z), 1 - 2 * (x * x + z * z), 2 * (y * z - w * x),
        2 * (x * z - w * y), 2 * (y * z + w * x), 1 - 2 * (x * x + y * y)], -1)
    return R.reshape(-1, 3, 3)              # R[i, row, col]


def rasterize(means3D, opacities, scales, rotations, colors, viewmatrix, projmatrix,
              bg, H, W, scale_modifier=1.0):
    """All tensor args are torch tensors (any float dtype; use float64)."""
    dt = means3D.dtype
    P = means3D.shape[0]
    vm = viewmatrix.reshape(4, 4).to(dt)      # row-vector convention
    pm = projmatrix.reshape(4, 4).to(dt)
    opac = opacities.reshape(P)
    ones = torch.ones(P, 1, dtype=dt)
    p_view = torch.cat([means3D, ones], 1) @ vm            # [P,4]
    vz = p_view[:, 2]
    R = _quat_to_R(rotations)
    L0 = R[:, :, 0] * (scale_modifier * scales[:, 0:1])
    L1 = R[:, :, 1] * (scale_modifier * scales[:, 1:2])
    L2 = R[:, :, 2]
    # M^T rows: (L0,0), (L1,0), (p,1);  B = M^T A with A = pm (row-major)
    Mt = torch.stack([torch.cat([L0, 0 * ones], 1), torch.cat([L1, 0 * ones], 1),
                      torch.cat([means3D, ones], 1)], 1)     # [P,3,4]
    B = Mt @ pm                                            # [P,3,4]
    Tu = B[:, :, 0] * (0.5 * W) + B[:, :, 3] * (0.5 * (W - 1))
    Tv = B[:, :, 1] * (0.5 * H) + B[:, :, 3] * (0.5 * (H - 1))
    Tw = B[:, :, 3]
    normal = L2 @ vm[:3, :3]
    cosv = -(p_view[:, :3] * normal).sum(-1)
    mult = torch.where(cosv > 0, 1.0, -1.0).to(dt).detach()
    normal = normal * mult[:, None]

    t = torch.tensor([9.0, 9.0, -1.0], dtype=dt)
    d = (t * Tw * Tw).sum(-1)
    ok = (vz > NEAR_N) & (cosv != 0) & (d != 0)
    dsafe = torch.where(d == 0, torch.ones_like(d), d)
    f = t[None] / dsafe[:, None]
    cx = (f * Tu * Tw).sum(-1)
    cy = (f * Tv * Tw).sum(-1)
    hx0 = cx * cx - (f * Tu * Tu).sum(-1)
    hy0 = cy * cy - (f * Tv * Tv).sum(-1)
    ex = torch.sqrt(torch.clamp(hx0, min=1e-4))
    ey = torch.sqrt(torch.clamp(hy0, min=1e-4))
    if RADIUS_FORMULA == 0:
        radius = torch.ceil(torch.maximum(torch.maximum(ex, ey),
                                          torch.tensor(3.0 * FILTER_SIZE, dtype=dt))).detach()
    else:
        radius = torch.ceil(3.0 * torch.maximum(torch.maximum(ex, ey),
                                                torch.tensor(FILTER_SIZE, dtype=dt))).detach()
    gx, gy = (W + 15) // 16, (H + 15) // 16
    ri = radius.to(torch.int64).to(dt)

    def _clampi(v, hi):
        return torch.clamp(torch.trunc(v).to(torch.int64), 0, hi)
    x0 = _clampi((cx.detach() - ri) / 16.0, gx)
    y0 = _clampi((cy.detach() - ri) / 16.0, gy)
    x1 = _clampi((cx.detach() + ri + 15.0) / 16.0, gx)
    y1 = _clampi((cy.detach() + ri + 15.0) / 16.0, gy)
    ok = ok & ((x1 - x0) * (y1 - y0) > 0)
    radii = torch.where(ok, radius.to(torch.int64), torch.zeros_like(x0))

    # global depth order (depth then index == stable sort of the duplication order)
    order = torch.argsort(torch.where(ok, vz.detach(), torch.full_like(vz, float("inf"))), stable=True)
    order = order[ok[order]]

    def g(a):
        return a[order]
    Tu_, Tv_, Tw_, cx_, cy_ = g(Tu), g(Tv), g(Tw), g(cx), g(cy)
    nrm_, op_, col_ = g(normal), g(opac), g(colors)
    x0_, y0_, x1_, y1_ = g(x0), g(y0), g(x1), g(y1)

    ys, xs = torch.meshgrid(torch.arange(H), torch.arange(W), indexing="ij")
    pxf = xs.reshape(-1, 1).to(dt)                         # [HW,1]
    pyf = ys.reshape(-1, 1).to(dt)
    tx = (xs.reshape(-1, 1) // 16)
    ty = (ys.reshape(-1, 1) // 16)
    member = (tx >= x0_[None]) & (tx < x1_[None]) & (ty >= y0_[None]) & (ty < y1_[None])

    k = pxf[..., None] * Tw_[None] - Tu_[None]              # [HW,Q,3]
    l = pyf[..., None] * Tw_[None] - Tv_[None]
    p = torch.cross(k, l, dim=-1)
    pz = p[..., 2]
    valid = member & (pz != 0)
    pzs = torch.where(pz == 0, torch.ones_like(pz), pz)
    s0, s1 = p[..., 0] / pzs, p[..., 1] / pzs
    rho3d = s0 * s0 + s1 * s1
    dx, dy = cx_[None] - pxf, cy_[None] - pyf
    rho2d = FILTER_INV_SQUARE * (dx * dx + dy * dy)
    use3d = rho3d <= rho2d
    rho = torch.where(use3d, rho3d, rho2d)
    depth = torch.where(use3d, s0 * Tw_[None, :, 0] + s1 * Tw_[None, :, 1] + Tw_[None, :, 2],
                        Tw_[None, :, 2].expand_as(s0))
    valid = valid & (depth >= NEAR_N)
    power = -0.5 * rho
    valid = valid & ~(power > 0)
    Gs = torch.exp(power)
    a_raw = op_[None] * Gs
    alpha = a_raw + (torch.clamp(a_raw, max=0.99) - a_raw).detach()   # straight-through clamp
    valid = valid & (alpha.detach() >= 1.0 / 255.0)
    a_eff = torch.where(valid, alpha, torch.zeros_like(alpha))
    one_m = 1 - a_eff
    Tincl = torch.cumprod(one_m, dim=1)                    # T after element j (if all accepted)
    Texcl = torch.cat([torch.ones_like(Tincl[:, :1]), Tincl[:, :-1]], 1)
    stop = valid & (Tincl.detach() < 1e-4)
    alive = (torch.cumsum(stop.to(torch.int64), 1) == 0)
    contrib = valid & alive
    cf = contrib.to(dt)
    # transmittance restricted to contributing elements
    a_c = a_eff * cf
    Tin = torch.cumprod(1 - a_c, dim=1)
    Tex = torch.cat([torch.ones_like(Tin[:, :1]), Tin[:, :-1]], 1)
    w = a_c * Tex
    T_final = Tin[:, -1] if Tin.shape[1] > 0 else torch.ones(H * W, dtype=dt)
    dsafe2 = torch.where(contrib, depth, torch.ones_like(depth))
    m = FAR_N / (FAR_N - NEAR_N) * (1 - NEAR_N / dsafe2)
    m = torch.where(contrib, m, torch.zeros_like(m))
    A_before = 1 - Tex
    M1_before = torch.cumsum(m * w, 1) - m * w
    M2_before = torch.cumsum(m * m * w, 1) - m * m * w
    dist = ((m * m * A_before + M2_before - 2 * m * M1_before) * w).sum(1)
    Dacc = (torch.where(contrib, depth, torch.zeros_like(depth)) * w).sum(1)
    Nacc = (nrm_[None] * w[..., None]).sum(1)
    Cacc = (col_[None] * w[..., None]).sum(1)
    # median depth: last contributing element whose incoming T > 0.5
    med_mask = contrib & (Tex.detach() > 0.5)
    Q = med_mask.shape[1]
    if Q > 0:
        idx = torch.arange(Q)[None].expand_as(med_mask)
        med_idx = torch.where(med_mask, idx, torch.full_like(idx, -1)).max(1).values
        has = med_idx >= 0
        med_depth = torch.where(has, depth.gather(1, med_idx.clamp(min=0)[:, None])[:, 0],
                                torch.zeros(H * W, dtype=dt))
    else:
        med_depth = torch.zeros(H * W, dtype=dt)
    bgv = bg.to(dt)
    color = (Cacc + T_final[:, None] * bgv[None]).T.reshape(3, H, W)
    allmap = torch.stack([Dacc, 1 - T_final, Nacc[:, 0], Nacc[:, 1], Nacc[:, 2], med_depth, dist], 0)
    allmap = allmap.reshape(7, H, W)
    return color, radii.to(torch.int32), allmap
