"""Generates tests/golden/transport_small.npz by running the REFERENCE's own transport package
(/root/reference/transport/{__init__,transport,path,integrators,utils}.py) on CPU with a small analytic model: training
losses, SDE sampling (Euler-Maruyama / Heun, every diffusion form, every last-step rule) and the Hutchinson likelihood ODE
(`sample_sde` / `sample_ode_likelihood` were untested in round 1).  torchdiffeq is absent from
this image: _ref_stubs.py restates its fixed-grid solvers (tuple states flattened like torchdiffeq does).  The random
draws (`th.randn` per SDE step, `th.randint` per likelihood-drift evaluation, `th.randn_like` / `th.rand` in
`Transport.sample`) are made under `torch.manual_seed`, so a mirror that draws in the same order reproduces the numbers.
Only runs inside the build container (needs /root/reference).
    python tests/golden/make_transport_golden.py
"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
from _ref_stubs import *  # noqa: F401,F403
import json
import numpy as np
import torch
import transport as ref_transport

OUT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "transport_small.npz")
CASES = [("Linear", "velocity"), ("GVP", "velocity"), ("VP", "noise"), ("Linear", "score"), ("GVP", "noise"), ("VP", "velocity")]
SDE = [("Euler", "SBDM", 1.0, "Mean"), ("Heun", "sigma", 0.7, "Tweedie"), ("Euler", "constant", 0.5, "Euler"),
       ("Euler", "linear", 1.0, None), ("Heun", "decreasing", 1.0, "Mean"), ("Euler", "increasing-decreasing", 0.3, "Mean"),
       ("Euler", "inccreasing-decreasing", 0.3, "Mean")]     # the reference's key is spelled with two c's; the right spelling raises


EPS = 1e-2            # train_eps = sample_eps: keeps t away from the end points where score <-> velocity conversions divide by 0


def make_model(A):
    return lambda x, t, **kw: torch.tanh(x @ A) * (1.0 + t[:, None]) - 0.3 * x


def main():
    g = torch.Generator().manual_seed(7)
    A = torch.randn(6, 6, generator=g) * 0.5
    x0 = torch.randn(3, 6, generator=g)
    model = make_model(A)
    save = {"A": A.numpy(), "x0": x0.numpy()}
    meta = {"cases": CASES, "sde": SDE, "sde_steps": 6, "like_steps": 5, "eps": EPS}
    for ci, (path_type, pred) in enumerate(CASES):
        tr = ref_transport.create_transport(path_type, pred, None, EPS, EPS, "uniform")
        sampler = ref_transport.Sampler(tr)
        torch.manual_seed(100 + ci)
        terms = tr.training_losses(model, x0)
        save["loss_%d" % ci] = terms["loss"].detach().numpy()
        for si, (method, form, norm, last) in enumerate(SDE):
            torch.manual_seed(200 + 10 * ci + si)
            try:
                xs = sampler.sample_sde(sampling_method=method, diffusion_form=form, diffusion_norm=norm, last_step=last,
                                        last_step_size=0.04, num_steps=6)(x0, model)
                save["sde_%d_%d" % (ci, si)] = torch.stack(xs, 0).detach().numpy()
            except Exception as e:                 # the reference's own behaviour for this setting: recorded, mirrored
                meta.setdefault("sde_raises", {})["%d_%d" % (ci, si)] = type(e).__name__
        for mi, method in enumerate(("euler", "heun2")):
            torch.manual_seed(300 + ci)
            with torch.no_grad():              # as the reference's eval loops run it (the drift re-enables grad itself)
                logp, z = sampler.sample_ode_likelihood(sampling_method=method, num_steps=5)(x0.clone(), model)
            save["like_%d_%d_logp" % (ci, mi)] = logp.detach().numpy()
            save["like_%d_%d_z" % (ci, mi)] = z.detach().numpy()
    save["meta"] = np.frombuffer(json.dumps(meta).encode(), dtype=np.uint8)
    np.savez_compressed(OUT, **save)
    print("wrote", OUT, len(save), "arrays")


if __name__ == "__main__":
    main()
