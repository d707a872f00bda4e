"""gaussiananything_b200.transport (host logic) against numbers the REFERENCE's own
transport package produced (tests/golden/make_transport_golden.py -> transport_small.npz): training losses, SDE sampling
with every sampler / diffusion form / last-step rule, the Hutchinson likelihood ODE, and the settings for which the
reference itself raises (same exception type here: drop-in error behaviour).  Random draws are made in the reference's
order, so the comparison is on values, not distributions."""
import json
import os

import numpy as np
import pytest
import torch

GOLD = os.path.join(os.path.dirname(__file__), "golden", "transport_small.npz")


def _load():
    z = np.load(GOLD)
    meta = json.loads(bytes(z["meta"]).decode())
    A, x0 = torch.tensor(z["A"]), torch.tensor(z["x0"])
    model = lambda x, t, **kw: torch.tanh(x @ A) * (1.0 + t[:, None]) - 0.3 * x
    return z, meta, x0, model


def _close(a, b, tol=2e-5):
    """Same non-finite pattern (where the reference itself overflows, so must the mirror) and close finite values."""
    a, b = a.double(), torch.as_tensor(b).double()
    fin = torch.isfinite(b)
    if not torch.equal(torch.isfinite(a), fin) or not torch.equal(torch.isnan(a), torch.isnan(b)):
        return False
    if not torch.equal(a[~fin & ~torch.isnan(b)], b[~fin & ~torch.isnan(b)]):       # +-inf in the same places
        return False
    if fin.sum() == 0:
        return True
    return float((a[fin] - b[fin]).abs().max()) <= tol * (1.0 + float(b[fin].abs().max()))


def _cases():
    meta = json.loads(bytes(np.load(GOLD)["meta"]).decode())
    return list(enumerate(tuple(c) for c in meta["cases"]))


@pytest.mark.parametrize("ci,case", _cases())
def test_training_losses_match_reference(ci, case):
    from gaussiananything_b200 import transport as tr
    z, meta, x0, model = _load()
    t = tr.create_transport(case[0], case[1], None, meta["eps"], meta["eps"], "uniform")
    torch.manual_seed(100 + ci)
    loss = t.training_losses(model, x0)["loss"]
    assert _close(loss, z["loss_%d" % ci]), (case, loss, z["loss_%d" % ci])


@pytest.mark.parametrize("ci,case", _cases())
def test_sample_sde_matches_reference(ci, case):
    from gaussiananything_b200 import transport as tr
    z, meta, x0, model = _load()
    s = tr.Sampler(tr.create_transport(case[0], case[1], None, meta["eps"], meta["eps"], "uniform"))
    raises = meta.get("sde_raises", {})
    seen = 0
    for si, (method, form, norm, last) in enumerate(meta["sde"]):
        torch.manual_seed(200 + 10 * ci + si)
        run = lambda: s.sample_sde(sampling_method=method, diffusion_form=form, diffusion_norm=norm, last_step=last,
                                   last_step_size=0.04, num_steps=meta["sde_steps"])(x0, model)
        key = "%d_%d" % (ci, si)
        if key in raises:
            exc = {"TypeError": TypeError, "NotImplementedError": NotImplementedError}[raises[key]]
            with pytest.raises(exc):
                run()
            continue
        xs = torch.stack(run(), 0)
        ref = z["sde_" + key]
        assert xs.shape == ref.shape == (meta["sde_steps"],) + tuple(x0.shape)
        assert _close(xs, ref), (case, method, form, last, float((xs - torch.tensor(ref)).abs().max()))
        seen += 1
    assert seen >= 5


@pytest.mark.parametrize("ci,case", _cases())
def test_likelihood_ode_matches_reference(ci, case):
    from gaussiananything_b200 import transport as tr
    z, meta, x0, model = _load()
    s = tr.Sampler(tr.create_transport(case[0], case[1], None, meta["eps"], meta["eps"], "uniform"))
    for mi, method in enumerate(("euler", "heun2")):
        torch.manual_seed(300 + ci)
        with torch.no_grad():
            logp, zz = s.sample_ode_likelihood(sampling_method=method, num_steps=meta["like_steps"])(x0.clone(), model)
        assert _close(zz, z["like_%d_%d_z" % (ci, mi)]), (case, method)
        assert _close(logp, z["like_%d_%d_logp" % (ci, mi)], 1e-4), (case, method, logp, z["like_%d_%d_logp" % (ci, mi)])
