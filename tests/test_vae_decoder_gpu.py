"""Row N1 end to end on the GPU: gaussiananything_b200.vae_decoder.SurfelDecoder against the goldens the reference's
own decoder produced (tests/golden/make_vae_golden.py) -- every stage the reference exposes."""
import os

import pytest
import torch

pytestmark = pytest.mark.gpu
GOLD = os.path.join(os.path.dirname(__file__), "golden", "vae_decoder_small.npz")


def rel(a, b):
    return float((a.double().cpu() - b.double().cpu()).norm() / b.double().cpu().norm().clamp_min(1e-30))


def test_surfel_decoder_matches_reference_goldens():
    from oracle import vae_decoder_oracle as vo
    from gaussiananything_b200.vae_decoder import SurfelDecoder
    g = vo.load_golden(GOLD)
    dev = torch.device("cuda:0")
    dec = SurfelDecoder(g["sd"], g["heads"], g["depth"], g["scene_max"], g["skip_weight"], device=dev)
    out = dec.decode(g["latent"].to(dev), g["xyz"].to(dev))
    torch.cuda.synchronize()
    # bf16 tensor-core operands against the reference's fp32 run: the bar is the bf16 bar of the DiT tests (2e-2
    # against a pure-fp32 golden); the packed surfels are bounded activations of those features
    pairs = [("latent_from_vit", "latent_from_vit", 2e-2), ("gaussian_base_pre_activate", "base_pre_activate", 3e-2),
             ("gaussians_base", "gaussians_base", 1e-2), ("gaussians_upsampled", "gaussians_upsampled", 1e-2),
             ("gaussians_upsampled_2", "gaussians_upsampled_2", 1e-2), ("gaussians_upsampled_3", "gaussians_upsampled_3", 1e-2),
             ("gaussians", "gaussians", 1e-2)]
    errs = {}
    for mine, ref, tol in pairs:
        assert out[mine].shape == g["out"][ref].shape, (mine, out[mine].shape, g["out"][ref].shape)
        assert torch.isfinite(out[mine]).all(), mine
        errs[mine] = (rel(out[mine], g["out"][ref]), tol)
    assert all(e < tol for e, tol in errs.values()), errs
    s = out["gaussians_upsampled_3"]
    assert torch.allclose(s[..., 6:10].norm(dim=-1), torch.ones(s.shape[:2], device=dev), atol=1e-5)
    assert (s[..., 3] > 0).all() and (s[..., 3] < 1).all() and (s[..., 4:6] > 0).all()
    # deterministic: same bits on a second run
    out2 = dec.decode(g["latent"].to(dev), g["xyz"].to(dev))
    assert torch.equal(out2["gaussians_upsampled_3"], s)


def test_decoded_surfels_feed_the_rasteriser():
    """N1's output buffer is the rasteriser's input: decode on the GPU, render one view with the CUDA rasteriser."""
    from oracle import surfel_oracle as so
    from oracle import vae_decoder_oracle as vo
    from gaussiananything_b200.gs_surfel import GaussianRenderer2DGS
    from gaussiananything_b200.vae_decoder import SurfelDecoder
    g = vo.load_golden(GOLD)
    dev = torch.device("cuda:0")
    dec = SurfelDecoder(g["sd"], g["heads"], g["depth"], g["scene_max"], g["skip_weight"], device=dev)
    surf = dec.decode(g["latent"].to(dev), g["xyz"].to(dev))["gaussians_upsampled_3"]          # [B, 6144, 13]
    view, proj, pos, tanfov = so.camera_from_pose25(so.orbit_pose25(30.0, 20.0))
    B = surf.shape[0]
    cv = torch.tensor(view, device=dev)[None, None].expand(B, 1, 4, 4).contiguous()
    cvp = torch.tensor(proj, device=dev)[None, None].expand(B, 1, 4, 4).contiguous()
    cp = torch.tensor(pos, device=dev)[None, None].expand(B, 1, 3).contiguous()
    img = GaussianRenderer2DGS(64, 3, {}).render(surf, cv, cvp, cp, tanfov)
    assert torch.isfinite(img["image"]).all() and float(img["alpha"].max()) > 0.05


def test_surfel_decoder_deployed_size_properties():
    """Deployed size (768 tokens, width 768, 12 blocks, cascade 8*4*3 -> 73 728 surfels per sample), random weights:
    the oracle cannot run this in seconds, so size-independent properties: shapes, finite, valid surfels, same bits on
    a second run, batch items independent."""
    from gaussiananything_b200.vae_decoder import SurfelDecoder, random_state_dict
    dev = torch.device("cuda:0")
    dec = SurfelDecoder(random_state_dict(768, 12, 10, seed=1), 12, 12, device=dev)
    torch.manual_seed(0)
    lat = torch.randn(2, 768, 10, device=dev)
    xyz = (torch.rand(2, 768, 3, device=dev) - 0.5) * 0.8
    out = dec.decode(lat, xyz)
    s = out["gaussians_upsampled_3"]
    assert s.shape == (2, 73728, 13) and out["gaussians"].shape == (2, 6144, 13)
    assert torch.isfinite(s).all()
    assert torch.allclose(s[..., 6:10].norm(dim=-1), torch.ones(s.shape[:2], device=dev), atol=1e-5)
    assert (s[..., 3] >= 0).all() and (s[..., 3] <= 1).all() and (s[..., 4:6] > 0).all()
    out2 = dec.decode(lat, xyz)
    assert torch.equal(out2["gaussians_upsampled_3"], s)
    out3 = dec.decode(lat.flip(0).contiguous(), xyz.flip(0).contiguous())
    assert rel(out3["gaussians_upsampled_3"].flip(0), s) < 1e-5



def test_surfel_ae_dropin_behaviours_and_multi_lod_render():
    """SurfelAE mirrors nsr.script_util.AE.forward on the decode / render behaviours; triplane_decode renders the
    four levels of detail (one batched launch set each, on four streams) == rendering each level on its own."""
    from gaussiananything_b200.gs_surfel import GaussianRenderer2DGS
    from gaussiananything_b200.vae_decoder import SurfelAE, SurfelDecoder, random_state_dict
    from tests.helpers import cameras
    dev = torch.device("cuda:0")
    D, depth = 128, 2
    dec = SurfelDecoder(random_state_dict(D, depth, 10, seed=3), D // 64, depth, device=dev)
    ae = SurfelAE(dec)
    torch.manual_seed(0)
    B, V = 2, 3
    lat = {"latent_normalized": torch.randn(B, D, 10, device=dev), "query_pcd_xyz": (torch.rand(B, D, 3, device=dev) - 0.5) * 0.6}
    ret = ae(latent=lat, behaviour="decode_gs_after_vae_no_render")
    for k, f in (("gaussians_base", 1), ("gaussians_upsampled", 8), ("gaussians_upsampled_2", 32), ("gaussians_upsampled_3", 96)):
        assert ret[k].shape == (B, D * f, 13)
    assert ret["gaussians"] is ret["gaussians_upsampled"] and "query_pcd_xyz" in ret
    vs, ps, cs, tf = cameras(B * V)
    c = {"cam_view": torch.tensor(vs, device=dev).reshape(B, V, 4, 4), "cam_view_proj": torch.tensor(ps, device=dev).reshape(B, V, 4, 4),
         "cam_pos": torch.tensor(cs, device=dev).reshape(B, V, 3), "tanfov": torch.tensor(tf)}
    bg = torch.tensor([0.0, 0.0, 0.0], device=dev)
    out = ae(img=None, c=c, latent=ret, behaviour="triplane_dec", bg_color=bg, render_all_scale=True)
    assert list(out) == ["gaussians_base", "gaussians_upsampled", "gaussians_upsampled_2", "gaussians_upsampled_3"]
    rnd = GaussianRenderer2DGS(512, 3, {})
    for key, size in SurfelAE.OUTPUT_SIZE.items():
        r = out[key]
        assert r["image"].shape == (B, V, 3, size, size) and r["image_depth"].shape == (B, V, 1, size, size)
        ref = rnd.render(ret[key], c["cam_view"], c["cam_view_proj"], c["cam_pos"], tf, bg_color=bg, output_size=size)
        assert torch.equal(r["image"], ref["image"]) and torch.equal(r["image_mask"], ref["alpha"])
        assert torch.equal(r["image_raw"], ref["image"] * 2 - 1)
    two = ae(latent=ret, c=c, behaviour="triplane_dec")             # rand_base_render: one random coarse level + the finest
    assert len(two) == 2 and list(two)[-1] == "gaussians_upsampled_3"
    full = ae(latent=lat, c=c, behaviour="decode_after_vae")
    assert "gaussians_upsampled_3" in full
    with pytest.raises(NotImplementedError):
        ae(img=torch.zeros(1), behaviour="enc")


def test_graph_replay_matches_eager_launches_bit_for_bit():
    """decode() replays a CUDA graph captured per batch size; the eager launch sequence (use_graph=False) must give the
    same bits, new inputs must reach the static buffers, returned tensors must survive later calls, and a call from
    torch.inference_mode (the reference's eval path) must work before and after the capture."""
    from oracle import vae_decoder_oracle as vo
    from gaussiananything_b200.vae_decoder import SurfelDecoder
    g = vo.load_golden(GOLD)
    dev = torch.device("cuda:0")
    dec = SurfelDecoder(g["sd"], g["heads"], g["depth"], g["scene_max"], g["skip_weight"], device=dev)
    lat, xyz = g["latent"].to(dev), g["xyz"].to(dev)
    torch.manual_seed(3)
    lat2 = lat + 0.3 * torch.randn_like(lat)
    dec.use_graph = False
    eager1, eager2, eager_one = dec.decode(lat, xyz), dec.decode(lat2, xyz), dec.decode(lat[:1], xyz[:1])
    assert not dec._graphs
    dec.use_graph = True
    with torch.inference_mode():
        first = dec.decode(lat, xyz)                    # captures
    assert lat.shape[0] in dec._graphs
    second = dec.decode(lat2, xyz)                      # replays with new inputs
    with torch.inference_mode():
        third = dec.decode(lat, xyz)
    torch.cuda.synchronize()
    for k in eager1:
        assert torch.equal(first[k], eager1[k]), k
        assert torch.equal(second[k], eager2[k]), k
        assert torch.equal(third[k], eager1[k]), k
    assert not torch.equal(first["gaussians"], second["gaussians"])
    # a different batch size gets its own graph
    one = dec.decode(lat[:1], xyz[:1])
    assert torch.equal(one["gaussians_upsampled_3"], eager_one["gaussians_upsampled_3"])
    assert set(dec._graphs) == {lat.shape[0], 1}
