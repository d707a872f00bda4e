"""Drop-in for the `diff_surfel_rasterization` module the reference imports at
/root/reference/nsr/gs_surfel.py:15 and drives at :85-114.

Same names, argument meaning and error behaviour as the upstream Python binding
(github.com/hbb1/diff-surfel-rasterization, diff_surfel_rasterization/__init__.py):
`GaussianRasterizationSettings` (NamedTuple), `GaussianRasterizer` (nn.Module)
whose forward returns `(color[3,H,W], radii[P] int32, allmap[7,H,W])`.  The
arithmetic runs in libga_b200.so; this file is argument marshalling only.
Not supported (the reference never uses them): spherical harmonics (`shs`) and
`cov3D_precomp` -- both raise.
"""
from typing import NamedTuple

import torch
import torch.nn as nn

from . import raster as _raster


class GaussianRasterizationSettings(NamedTuple):
    image_height: int
    image_width: int
    tanfovx: float
    tanfovy: float
    bg: torch.Tensor
    scale_modifier: float
    viewmatrix: torch.Tensor
    projmatrix: torch.Tensor
    sh_degree: int
    campos: torch.Tensor
    prefiltered: bool
    debug: bool


def rasterize_gaussians(means3D, means2D, sh, colors_precomp, opacities, scales, rotations,
                        cov3Ds_precomp, raster_settings):
    rs = raster_settings
    P = means3D.shape[0]
    gauss13 = torch.cat([means3D.reshape(P, 3), opacities.reshape(P, 1), scales.reshape(P, 2),
                         rotations.reshape(P, 4), colors_precomp.reshape(P, 3)], dim=1).float()[None]
    color, allmap, radii = _raster.rasterize_surfels_batched(
        gauss13, rs.viewmatrix.reshape(1, 1, 4, 4), rs.projmatrix.reshape(1, 1, 4, 4), rs.bg,
        int(rs.image_height), int(rs.image_width), float(rs.scale_modifier))
    return color[0, 0], radii[0, 0], allmap[0, 0]


class GaussianRasterizer(nn.Module):
    def __init__(self, raster_settings):
        super().__init__()
        self.raster_settings = raster_settings

    def markVisible(self, positions):
        """Frustum test of upstream `_C.mark_visible` (view-space z > 0.2)."""
        with torch.no_grad():
            vm = self.raster_settings.viewmatrix.reshape(4, 4).to(positions)
            z = positions @ vm[:3, 2] + vm[3, 2]
            return z > 0.2

    def forward(self, means3D, means2D, opacities, shs=None, colors_precomp=None,
                scales=None, rotations=None, cov3D_precomp=None):
        if (shs is None and colors_precomp is None) or (shs is not None and colors_precomp is not None):
            raise Exception('Please provide excatly one of either SHs or precomputed colors!')
        if ((scales is None or rotations is None) and cov3D_precomp is None) or \
                ((scales is not None or rotations is not None) and cov3D_precomp is not None):
            raise Exception('Please provide exactly one of either scale/rotation pair or precomputed 3D covariance!')
        if shs is not None:
            raise NotImplementedError("gaussiananything_b200: SH colours are not on the reference's path "
                                      "(sh_degree=0, colors_precomp; /root/reference/nsr/gs_surfel.py:94,108)")
        if cov3D_precomp is not None:
            raise NotImplementedError("gaussiananything_b200: cov3D_precomp is not on the reference's path")
        return rasterize_gaussians(means3D, means2D, shs, colors_precomp, opacities, scales, rotations,
                                   cov3D_precomp, self.raster_settings)
