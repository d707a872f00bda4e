"""ODE / SDE integrators behind the `ode` / `sde` classes of
/root/reference/transport/integrators.py:8-124, with the time integration done
in-tree (the reference delegates to torchdiffeq.odeint, an unpinned third-party
package, integrators.py:4,111).

Fixed-grid methods ('euler', 'midpoint', 'heun2'/'heun', 'rk4') step exactly on
t = linspace(t0, t1, num_steps), i.e. num_steps - 1 steps, and return the
stacked trajectory [num_steps, *x.shape] like odeint.  'dopri5' is an adaptive
Dormand-Prince 5(4) pair that lands on every grid point (same tolerances as
the reference call; the step-size controller is not bit-compatible with
torchdiffeq's, so for reproducible parity use a fixed-grid method).
"""
import torch as th

FIXED = ("euler", "midpoint", "heun2", "heun", "rk4")


def _fixed_step(method, f, t0, t1, y):
    dt = t1 - t0
    if method == "euler":
        return y + dt * f(t0, y)
    if method == "midpoint":
        return y + dt * f(t0 + 0.5 * dt, y + 0.5 * dt * f(t0, y))
    if method in ("heun2", "heun"):
        k1 = f(t0, y)
        return y + 0.5 * dt * (k1 + f(t1, y + dt * k1))
    if method == "rk4":                     # 3/8 rule, as torchdiffeq's fixed-grid rk4
        k1 = f(t0, y)
        k2 = f(t0 + dt / 3, y + dt * k1 / 3)
        k3 = f(t0 + dt * 2 / 3, y + dt * (k2 - k1 / 3))
        k4 = f(t1, y + dt * (k1 - k2 + k3))
        return y + dt * 0.125 * (k1 + 3 * (k2 + k3) + k4)
    raise NotImplementedError(method)


_DP_C = (1 / 5, 3 / 10, 4 / 5, 8 / 9, 1.0, 1.0)
_DP_A = ((1 / 5,), (3 / 40, 9 / 40), (44 / 45, -56 / 15, 32 / 9),
         (19372 / 6561, -25360 / 2187, 64448 / 6561, -212 / 729),
         (9017 / 3168, -355 / 33, 46732 / 5247, 49 / 176, -5103 / 18656),
         (35 / 384, 0.0, 500 / 1113, 125 / 192, -2187 / 6784, 11 / 84))
_DP_E = (35 / 384 - 1951 / 21600, 0.0, 500 / 1113 - 22642 / 50055, 125 / 192 - 451 / 720,
         -2187 / 6784 + 12231 / 42400, 11 / 84 - 649 / 6300, -1 / 60)


def _dopri5(f, y, ts, atol, rtol):
    out = [y]
    t = float(ts[0])
    h = float(ts[1] - ts[0])
    k1 = f(th.as_tensor(t, device=y.device), y)
    for target in [float(v) for v in ts[1:]]:
        while target - t > 1e-12:
            h = min(h, target - t)
            ks = [k1]
            for c, row in zip(_DP_C, _DP_A):
                yi = y + h * sum(a * k for a, k in zip(row, ks))
                ks.append(f(th.as_tensor(t + c * h, device=y.device), yi))
            y5 = y + h * sum(a * k for a, k in zip(_DP_A[-1], ks[:6]))
            err = h * sum(e * k for e, k in zip(_DP_E, ks))
            tol = atol + rtol * th.maximum(y.abs(), y5.abs())
            ratio = float(th.sqrt(th.mean((err.float() / tol.float()) ** 2)))
            if ratio <= 1.0:
                t, y, k1 = t + h, y5, ks[6]
            h = h * min(10.0, max(0.2, 0.9 * (ratio + 1e-10) ** -0.2))
        out.append(y)
    return th.stack(out, 0)


def odeint(f, y0, ts, method="dopri5", atol=1e-6, rtol=1e-3):
    """Integrates dy/dt = f(t, y); returns the state at every ts (stacked)."""
    if method in FIXED:
        ys, y = [y0], y0
        for i in range(len(ts) - 1):
            y = _fixed_step(method, f, ts[i], ts[i + 1], y)
            ys.append(y)
        return th.stack(ys, 0)
    if method == "dopri5":
        return _dopri5(f, y0, ts, atol, rtol)
    raise NotImplementedError("sampling_method %r" % (method,))


class ode:
    """ODE solver class (reference integrators.py:78-119)."""

    def __init__(self, drift, *, t0, t1, sampler_type, num_steps, atol, rtol):
        assert t0 < t1, "ODE sampler has to be in forward time"
        self.drift = drift
        self.t = th.linspace(t0, t1, num_steps)
        self.atol, self.rtol, self.sampler_type = atol, rtol, sampler_type

    def sample(self, x, model, **model_kwargs):
        if isinstance(x, tuple):
            raise NotImplementedError("tuple states (likelihood ODE) need sample_ode_likelihood")
        device = x.device

        def _fn(t, xx):
            tb = th.ones(xx.size(0), device=device) * t
            return self.drift(xx, tb, model, **model_kwargs)

        return odeint(_fn, x, self.t.to(device), method=self.sampler_type, atol=self.atol, rtol=self.rtol)


class sde:
    """Euler-Maruyama / Heun SDE sampler (reference integrators.py:8-76)."""

    def __init__(self, drift, diffusion, *, t0, t1, num_steps, sampler_type):
        assert t0 < t1, "SDE sampler has to be in forward time"
        self.num_timesteps = num_steps
        self.t = th.linspace(t0, t1, num_steps)
        self.dt = self.t[1] - self.t[0]
        self.drift, self.diffusion, self.sampler_type = drift, diffusion, sampler_type

    def _euler_maruyama(self, x, mean_x, t, model, **kw):
        w = th.randn(x.size()).to(x)
        tb = th.ones(x.size(0)).to(x) * t
        mean_x = x + self.drift(x, tb, model, **kw) * self.dt
        return mean_x + th.sqrt(2 * self.diffusion(x, tb)) * (w * th.sqrt(self.dt)), mean_x

    def _heun(self, x, _, t, model, **kw):
        w = th.randn(x.size()).to(x)
        tb = th.ones(x.size(0)).to(x) * t
        xhat = x + th.sqrt(2 * self.diffusion(x, tb)) * (w * th.sqrt(self.dt))
        k1 = self.drift(xhat, tb, model, **kw)
        k2 = self.drift(xhat + self.dt * k1, tb + self.dt, model, **kw)
        return xhat + 0.5 * self.dt * (k1 + k2), xhat

    def sample(self, init, model, **model_kwargs):
        try:
            step = {"Euler": self._euler_maruyama, "Heun": self._heun}[self.sampler_type]
        except KeyError:
            raise NotImplementedError("Smapler type not implemented.")
        x, mean_x, samples = init, init, []
        for ti in self.t[:-1]:
            with th.no_grad():
                x, mean_x = step(x, mean_x, ti, model, **model_kwargs)
                samples.append(x)
        return samples
