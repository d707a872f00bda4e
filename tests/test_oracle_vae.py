"""The VAE-decoder oracle (latent tokens -> surfels) against goldens produced by the
reference's own code (tests/golden/make_vae_golden.py).  CPU only; no CUDA path exists for this row yet."""
import os

import torch

from oracle import vae_decoder_oracle as vo

GOLD = os.path.join(os.path.dirname(__file__), "golden", "vae_decoder_small.npz")


def rel(a, b):
    return float((a.double() - b.double()).norm() / b.double().norm().clamp_min(1e-30))


def test_vae_decoder_oracle_matches_reference_goldens_stage_by_stage():
    g = vo.load_golden(GOLD)
    with torch.no_grad():
        out = vo.decode(g["sd"], g["latent"], g["xyz"], g["heads"], g["depth"], g["scene_max"], g["skip_weight"])
    for k, ref in g["out"].items():
        assert out[k].shape == ref.shape, (k, out[k].shape, ref.shape)
        assert rel(out[k], ref) < 2e-6, (k, rel(out[k], ref))


def test_cascade_shapes_and_surfel_invariants():
    """8 * 4 * 3 children per token; the packed 13 channels are valid rasteriser input: unit quaternions, opacity and
    colour in (0,1), positive scales; every child stays within one offset radius of its parent."""
    g = vo.load_golden(GOLD)
    with torch.no_grad():
        out = vo.decode(g["sd"], g["latent"], g["xyz"], g["heads"], g["depth"], g["scene_max"], g["skip_weight"])
    B, N = g["latent"].shape[:2]
    assert out["gaussians_base"].shape == (B, N, 13)
    assert out["gaussians_upsampled"].shape == (B, N * 8, 13)
    assert out["gaussians_upsampled_2"].shape == (B, N * 32, 13)
    assert out["gaussians_upsampled_3"].shape == (B, N * 96, 13)
    for k in ("gaussians_base", "gaussians_upsampled", "gaussians_upsampled_2", "gaussians_upsampled_3"):
        s = out[k]
        assert torch.allclose(s[..., 6:10].norm(dim=-1), torch.ones(s.shape[:2]), atol=1e-5)
        assert (s[..., 3] > 0).all() and (s[..., 3] < 1).all()
        assert (s[..., 4:6] > 0).all()
        assert (s[..., 10:] >= 0).all() and (s[..., 10:] <= 1).all()
    parent = out["gaussians_upsampled"][..., :3].reshape(B, N, 8, 3)
    child = out["gaussians_upsampled_2"][..., :3].reshape(B, N, 8, 4, 3)
    assert float((child - parent[:, :, :, None]).abs().max()) <= 0.5 * g["scene_max"] + 1e-6


def test_decoded_surfels_render_through_the_surfel_oracle():
    """Row N1's output is row a1's input: the [B, N*f, 13] buffer the decoder produces goes straight into the
    rasteriser (here the CPU oracle, one 64 x 64 view) -- finite image, part of it covered."""
    import numpy as np
    from oracle import surfel_oracle as so
    g = vo.load_golden(GOLD)
    with torch.no_grad():
        out = vo.decode(g["sd"], g["latent"], g["xyz"], g["heads"], g["depth"], g["scene_max"], g["skip_weight"])
    s = out["gaussians_upsampled_3"][0].numpy().astype(np.float32)          # [6144, 13]
    view, proj, _pos, _tan = so.camera_from_pose25(so.orbit_pose25(30.0, 20.0))
    r = so.rasterize(s[:, 0:3], s[:, 3:4], s[:, 4:6], s[:, 6:10], s[:, 10:13], view, proj,
                     np.zeros(3, np.float32), 64, 64)
    assert np.isfinite(r["color"]).all() and np.isfinite(r["allmap"]).all()
    alpha = r["allmap"][1]
    assert float(alpha.max()) > 0.05 and int((r["radii"] > 0).sum()) > 100
