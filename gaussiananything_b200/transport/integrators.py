"""ODE / SDE integrators behind the `ode` / `sde` classes of
/root/reference/transport/integrators.py:8-124, with the time integration done
in-tree (the reference delegates to torchdiffeq.odeint, an unpinned third-party
package, integrators.py:4,111).

Fixed-grid methods ('euler', 'midpoint', 'heun2'/'heun', 'rk4') step exactly on
t = linspace(t0, t1, num_steps), i.e. num_steps - 1 steps, and return the
stacked trajectory [num_steps, *x.shape] like odeint.  'dopri5' is the adaptive
Dormand-Prince 5(4) pair with torchdiffeq's published controller (initial-step
heuristic, safety 0.9, growth 10 / shrink 0.2, RMS norm) and dense output:
steps are NOT clamped to the output grid, grid states are interpolated.
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
_DP_E = (35 / 384 - 1951 / 21600, 0.0, 500 / 1113 - 22642 / 50085, 125 / 192 - 451 / 720,
         -2187 / 6784 + 12231 / 42400, 11 / 84 - 649 / 6300, -1 / 60)
# mid-point weights of the Dormand-Prince pair (Shampine's 4th-order dense output)
_DP_MID = (6025192743 / 30085553152 / 2, 0.0, 51252292925 / 65400821598 / 2, -2691868925 / 45128329728 / 2,
           187940372067 / 1594534317056 / 2, -1776094331 / 19743644256 / 2, 11237099 / 235043384 / 2)


def _rms(x):
    return float(th.sqrt(th.mean(x.to(th.float64 if x.dtype == th.float64 else th.float32) ** 2)))


def _initial_step(f, tt, t0, y0, f0, order, rtol, atol):
    """Hairer-Norsett-Wanner II.4 starting step, as torchdiffeq's adaptive solvers choose it."""
    scale = atol + y0.abs() * rtol
    d0, d1 = _rms(y0 / scale), _rms(f0 / scale)
    h0 = 1e-6 if (d0 < 1e-5 or d1 < 1e-5) else 0.01 * d0 / d1
    f1 = f(tt(t0 + h0), y0 + h0 * f0)
    d2 = _rms((f1 - f0) / scale) / h0
    if d1 <= 1e-15 and d2 <= 1e-15:
        h1 = max(1e-6, h0 * 1e-3)
    else:
        h1 = (0.01 / max(d1, d2)) ** (1.0 / (order + 1))
    return min(100 * h0, h1)


def _dopri5(f, y, ts, atol, rtol, max_steps=2 ** 31 - 1):
    """Adaptive Dormand-Prince 5(4) with the step-size controller and the dense output of torchdiffeq's
    `dopri5` (the solver the reference calls by default, transport/integrators.py:111,
    nsr/lsgm/flow_matching_trainer.py:715): steps are chosen from the error estimate only -- NOT clamped to the
    output grid -- and the state at every grid point comes from the 4th-order interpolant over the step that
    contains it.  With the reference's num_steps=250 this takes tens of steps (6 NFE each), not >= 249.  As in
    torchdiffeq the last step may reach past ts[-1].  Raises on a non-finite error estimate or step underflow."""
    t = float(ts[0])
    dev = y.device
    tdt = ts.dtype if th.is_tensor(ts) and ts.is_floating_point() else th.float32
    tt = lambda v: th.as_tensor(v, device=dev, dtype=tdt)       # time keeps the grid's dtype, as in torchdiffeq
    f0 = f(tt(t), y)
    h = _initial_step(f, tt, t, y, f0, 4, rtol, atol)
    if not (h == h and h > 0.0):
        raise RuntimeError("dopri5: non-finite initial step (is the model output finite?)")
    out = [y]
    interp = None                       # (t_lo, t_hi, coefficients) of the last accepted step
    n = 0
    for target in [float(v) for v in ts[1:]]:
        while interp is None or target > interp[1]:
            n += 1
            if n > max_steps:
                raise RuntimeError("dopri5: max_num_steps exceeded")
            if not (t + h > t):
                raise RuntimeError("dopri5: underflow in the step size at t=%g (h=%g)" % (t, h))
            ks = [f0]
            for c, row in zip(_DP_C, _DP_A):
                yi = y + h * sum(a * k for a, k in zip(row, ks) if a != 0.0)
                ks.append(f(tt(t + c * h), yi))
            y1 = y + h * sum(a * k for a, k in zip(_DP_A[-1], ks[:6]) if a != 0.0)   # == the state of stage 7 (FSAL)
            err = h * sum(e * k for e, k in zip(_DP_E, ks) if e != 0.0)
            tol = atol + rtol * th.maximum(y.abs(), y1.abs())
            ratio = _rms(err / tol)
            if ratio != ratio or ratio == float("inf"):
                raise RuntimeError("dopri5: non-finite error estimate at t=%g (h=%g)" % (t, h))
            if ratio <= 1.0:
                f1 = ks[6]
                ymid = y + h * sum(m * k for m, k in zip(_DP_MID, ks) if m != 0.0)
                a = 2 * h * (f1 - f0) - 8 * (y1 + y) + 16 * ymid
                b = h * (5 * f0 - 3 * f1) + 18 * y + 14 * y1 - 32 * ymid
                c = h * (f1 - 4 * f0) - 11 * y - 5 * y1 + 16 * ymid
                interp = (t, t + h, (a, b, c, h * f0, y))
                t, y, f0 = t + h, y1, f1
            # step-size controller (safety 0.9, growth <= 10, shrink >= 0.2; an accepted step never shrinks)
            if ratio == 0.0:
                h = h * 10.0
            else:
                dfactor = 1.0 if ratio < 1.0 else 0.2
                h = h * min(10.0, max(0.9 / ratio ** 0.2, dfactor))
        lo, hi, (a, b, c, d, e) = interp
        x = (target - lo) / (hi - lo)
        out.append(e + x * (d + x * (c + x * (b + x * a))))
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
