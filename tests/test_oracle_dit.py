"""CPU: the DiT oracle and the transport mirror against golden vectors produced by the reference's OWN code
(tests/golden/make_dit_golden.py), plus the DiT module's reference-compatible parameter layout."""
import os

import pytest
import torch

from oracle import dit_oracle as do

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
NAMES = ["dit_stage1_small", "dit_stage2_small", "dit_stage2_concat_small"]


def rel(a, b):
    return float((a.double() - b.double()).norm() / b.double().norm())


@pytest.mark.parametrize("name", NAMES)
def test_oracle_matches_reference_forward(name):
    g = do.load_golden(os.path.join(GOLD, name + ".npz"))
    c = g["cfg"]
    y, acts = do.forward(g["sd"], g["x"], g["t"], g["ctx"], c["heads"], c["depth"], return_acts=True)
    assert rel(y, g["y"]) < 5e-6
    for k, v in g["acts"].items():
        assert rel(acts[k], v) < 5e-6, k
    yc = do.forward_with_cfg(g["sd"], g["x"], g["t"], g["ctx"], 4.0, c["heads"], c["depth"])
    assert rel(yc, g["y_cfg"]) < 5e-6
    half = yc.shape[0] // 2
    assert torch.equal(yc[:half], yc[half:])            # CFG output is duplicated (dit_i23d.py:171)


@pytest.mark.parametrize("name", NAMES)
def test_bf16_emulation_stays_close(name):
    g = do.load_golden(os.path.join(GOLD, name + ".npz"))
    c = g["cfg"]
    ye = do.forward(g["sd"], g["x"], g["t"], g["ctx"], c["heads"], c["depth"], emulate_bf16=True)
    assert 1e-5 < rel(ye, g["y"]) < 2e-2


@pytest.mark.parametrize("name", NAMES)
def test_transport_mirror_matches_reference_trajectories(name):
    from gaussiananything_b200 import transport as tr
    g = do.load_golden(os.path.join(GOLD, name + ".npz"))
    c = g["cfg"]
    model = lambda x, t, context, cfg_scale: do.forward_with_cfg(g["sd"], x, t, context, cfg_scale, c["heads"], c["depth"])
    s = tr.Sampler(tr.create_transport("GVP", "velocity", None, None, None, "lognorm"))
    te = s.sample_ode(sampling_method="euler", num_steps=5)(g["x"], model, context=g["ctx"], cfg_scale=4.0)
    th_ = s.sample_ode(sampling_method="heun2", num_steps=4)(g["x"], model, context=g["ctx"], cfg_scale=4.0)
    assert te.shape == g["traj_euler"].shape and rel(te, g["traj_euler"]) < 5e-6
    assert th_.shape == g["traj_heun"].shape and rel(th_, g["traj_heun"]) < 5e-6


def test_transport_api_and_solvers():
    import math
    from gaussiananything_b200 import transport as tr
    from gaussiananything_b200.transport import path
    t = tr.create_transport()                                   # Linear + velocity: eps = 0, t in [0, 1]
    assert isinstance(t.path_sampler, path.ICPlan) and (t.train_eps, t.sample_eps) == (0, 0)
    assert t.check_interval(0, 0, eval=True) == (0, 1)
    tv = tr.create_transport("VP", "noise")
    assert isinstance(tv.path_sampler, path.VPCPlan) and tv.sample_eps == 1e-3
    with pytest.raises(ValueError):
        tr.create_transport(snr_type="bogus")
    # dx/dt = -x  ->  x(1) = x0 / e ; orders of accuracy of the in-tree fixed-grid solvers + adaptive dopri5
    s = tr.Sampler(t)
    x0 = torch.ones(2, 3)
    model = lambda x, tt: -x
    exact = math.exp(-1.0)
    errs = {}
    for m, n in [("euler", 101), ("midpoint", 21), ("heun2", 21), ("rk4", 11), ("dopri5", 3)]:
        traj = s.sample_ode(sampling_method=m, num_steps=n, atol=1e-8, rtol=1e-8)(x0, model)
        assert traj.shape == (n, 2, 3)
        errs[m] = abs(float(traj[-1][0, 0]) - exact)
    assert errs["euler"] < 3e-3 and errs["midpoint"] < 2e-4 and errs["heun2"] < 2e-4
    assert errs["rk4"] < 1e-5 and errs["dopri5"] < 1e-6
    # GVP coefficients: alpha^2 + sigma^2 = 1 and plan() returns d/dt of the interpolant
    p = path.GVPCPlan()
    tt = torch.rand(5)
    a, da = p.compute_alpha_t(tt)
    sg, ds = p.compute_sigma_t(tt)
    assert torch.allclose(a * a + sg * sg, torch.ones(5), atol=1e-6)
    x1, xn = torch.randn(5, 4), torch.randn(5, 4)
    _, xt, ut = p.plan(tt, xn, x1)
    assert torch.allclose(xt, a[:, None] * x1 + sg[:, None] * xn) and torch.allclose(ut, da[:, None] * x1 + ds[:, None] * xn)
    # training loss runs with any callable model
    loss = t.training_losses(lambda x, tt: torch.zeros_like(x), torch.randn(4, 6, 3))["loss"]
    assert loss.shape == (4,)


def test_dit_module_has_reference_state_dict_layout():
    from gaussiananything_b200 import dit
    g = do.load_golden(os.path.join(GOLD, "dit_stage1_small.npz"))
    c = g["cfg"]
    m = dit.DiT_I23D_PCD_PixelArt_noclip(input_size=32, num_classes=0, learn_sigma=False, in_channels=c["cin"],
                                         context_dim=c["ctx_dim"], roll_out=True, pooling_ctx_dim=768, patch_size=1,
                                         depth=c["depth"], hidden_size=c["hidden"], num_heads=c["heads"], use_clay_ca=True)
    keys = set(m.state_dict().keys())
    missing = [k for k in g["sd"] if k not in keys]
    assert not missing, missing
    for k, v in g["sd"].items():
        assert tuple(m.state_dict()[k].shape) == tuple(v.shape), k
    # keys the reference carries but never uses are present too (strict checkpoint loading)
    for k in ("clip_spatial_proj.y_proj.fc1.weight", "cap_embedder.1.weight", "attention_y_norm.weight",
              "blocks.0.attention_y_norm.weight", "adaLN_modulation.1.weight", "pooled_vec_embedder.0.weight"):
        assert k in keys, k
    assert m.in_channels == c["cin"] and m.roll_out and len(m.blocks) == c["depth"]
    with pytest.raises(RuntimeError):                     # no CPU fallback
        m(g["x"], g["t"], g["ctx"])
    m2 = dit.DiT_models["DiT-PixArt-PCD-CLAY-stage2-L"]
    assert callable(m2) and set(dit.DiT_models) >= {"DiT-PixArt-PCD-CLAY-L", "DiT-PixArt-PCD-CLAY-B"}


@pytest.mark.parametrize("name,cin", [("DiT-PixArt-PCD-CLAY-B", 3), ("DiT-PixArt-PCD-CLAY-L", 3),
                                      ("DiT-PixArt-PCD-CLAY-stage2-L", 10)])
def test_registry_entries_load_reference_checkpoints_strictly(name, cin):
    """The mirror modules have EXACTLY the reference's state_dict keys and shapes (fixture written by
    tests/golden/make_dit_keys.py from the reference's own registry entries): load_state_dict(strict=True) works."""
    import json
    import torch
    from gaussiananything_b200 import dit
    want = json.load(open(os.path.join(GOLD, "dit_state_dict_keys.json")))[name]
    with torch.device("meta"):                      # shapes only: DiT-L is 1.6 GB of fp32 parameters
        m = dit.DiT_models[name](input_size=32, num_classes=0, learn_sigma=False, in_channels=cin, context_dim=1024,
                                 roll_out=True, pooling_ctx_dim=768)
    got = {k: list(v.shape) for k, v in m.state_dict().items()}
    assert sorted(got) == sorted(want), (sorted(set(want) - set(got))[:5], sorted(set(got) - set(want))[:5])
    assert got == want
