"""API-name adapter for /root/reference/nsr/gaussian_renderer/__init__.py:18-100
(`render(viewpoint_camera, pc, pipe, bg_color, scaling_modifier, override_color)`).

That reference function is vestigial (it imports the 3-DoF Inria rasteriser and
has no callers); BASELINE.json's north_star names it, so the same
call shape is offered on top of the surfel kernels: `pc` must expose 2-DoF
scales (`get_scaling` [P,2]) and either `override_color` or `get_features`
already reduced to RGB [P,3].
"""
import torch

from .diff_surfel_rasterization import GaussianRasterizationSettings, GaussianRasterizer


def render(viewpoint_camera, pc, pipe, bg_color, scaling_modifier=1.0, override_color=None):
    import math
    screenspace_points = torch.zeros_like(pc.get_xyz, dtype=pc.get_xyz.dtype, requires_grad=True,
                                          device=pc.get_xyz.device) + 0
    tanfovx = math.tan(viewpoint_camera.FoVx * 0.5)
    tanfovy = math.tan(viewpoint_camera.FoVy * 0.5)
    raster_settings = GaussianRasterizationSettings(
        image_height=int(viewpoint_camera.image_height),
        image_width=int(viewpoint_camera.image_width),
        tanfovx=tanfovx, tanfovy=tanfovy, bg=bg_color, scale_modifier=scaling_modifier,
        viewmatrix=viewpoint_camera.world_view_transform,
        projmatrix=viewpoint_camera.full_proj_transform,
        sh_degree=getattr(pc, "active_sh_degree", 0), campos=viewpoint_camera.camera_center,
        prefiltered=False, debug=getattr(pipe, "debug", False))
    rasterizer = GaussianRasterizer(raster_settings=raster_settings)
    colors = override_color if override_color is not None else pc.get_features
    if colors.dim() != 2 or colors.shape[1] != 3:
        raise NotImplementedError("gaussiananything_b200.gaussian_renderer.render needs RGB colours [P,3]")
    rendered_image, radii, allmap = rasterizer(
        means3D=pc.get_xyz, means2D=screenspace_points, shs=None, colors_precomp=colors,
        opacities=pc.get_opacity, scales=pc.get_scaling, rotations=pc.get_rotation, cov3D_precomp=None)
    return {"render": rendered_image, "viewspace_points": screenspace_points,
            "visibility_filter": radii > 0, "radii": radii, "allmap": allmap}
