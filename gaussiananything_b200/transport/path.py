"""Interpolant ("coupling plan") classes with the method names of
/root/reference/transport/path.py (ICPlan :18-136, VPCPlan :139-171, GVPCPlan
:174-191).  x_t = alpha_t x1 + sigma_t x0, t: 0 (noise) -> 1 (data).
Each plan only states its coefficient functions; everything else derives from them.
"""
import math

import torch as th


def expand_t_like_x(t, x):
    """[B] -> [B,1,...,1] so that t broadcasts against x."""
    return t.view(t.size(0), *([1] * (x.dim() - 1)))


class ICPlan:
    """Linear interpolant: alpha = t, sigma = 1 - t."""

    def __init__(self, sigma=0.0):
        self.sigma = sigma

    # -- coefficients: (value, time derivative)
    def compute_alpha_t(self, t):
        return t, 1

    def compute_sigma_t(self, t):
        return 1 - t, -1

    def compute_d_alpha_alpha_ratio_t(self, t):
        return 1 / t

    # -- SDE pieces in score parametrisation
    def compute_drift(self, x, t):
        t = expand_t_like_x(t, x)
        ratio = self.compute_d_alpha_alpha_ratio_t(t)
        sigma, d_sigma = self.compute_sigma_t(t)
        return -(ratio * x), ratio * (sigma ** 2) - sigma * d_sigma

    def compute_diffusion(self, x, t, form="constant", norm=1.0):
        t = expand_t_like_x(t, x)
        if form == "constant":
            return norm
        if form == "SBDM":
            return norm * self.compute_drift(x, t)[1]
        if form == "sigma":
            return norm * self.compute_sigma_t(t)[0]
        if form == "linear":
            return norm * (1 - t)
        if form == "decreasing":
            return 0.25 * (norm * th.cos(math.pi * t) + 1) ** 2
        if form == "inccreasing-decreasing":
            return norm * th.sin(math.pi * t) ** 2
        raise NotImplementedError(f"Diffusion form {form} not implemented")

    # -- conversions between parametrisations
    def _coeffs(self, x, t):
        t = expand_t_like_x(t, x)
        return self.compute_alpha_t(t) + self.compute_sigma_t(t)

    def get_score_from_velocity(self, velocity, x, t):
        alpha, d_alpha, sigma, d_sigma = self._coeffs(x, t)
        r = alpha / d_alpha
        return (r * velocity - x) / (sigma ** 2 - r * d_sigma * sigma)

    def get_noise_from_velocity(self, velocity, x, t):
        alpha, d_alpha, sigma, d_sigma = self._coeffs(x, t)
        r = alpha / d_alpha
        return (r * velocity - x) / (r * d_sigma - sigma)

    def get_velocity_from_score(self, score, x, t):
        drift, var = self.compute_drift(x, t)
        return var * score - drift

    # -- training pairs
    def compute_mu_t(self, t, x0, x1):
        t = expand_t_like_x(t, x1)
        return self.compute_alpha_t(t)[0] * x1 + self.compute_sigma_t(t)[0] * x0

    def compute_xt(self, t, x0, x1):
        return self.compute_mu_t(t, x0, x1)

    def compute_ut(self, t, x0, x1, xt):
        t = expand_t_like_x(t, x1)
        return self.compute_alpha_t(t)[1] * x1 + self.compute_sigma_t(t)[1] * x0

    def plan(self, t, x0, x1):
        xt = self.compute_xt(t, x0, x1)
        return t, xt, self.compute_ut(t, x0, x1, xt)


class VPCPlan(ICPlan):
    """Variance preserving path, beta(t) linear between sigma_max (t=0) and sigma_min (t=1)."""

    def __init__(self, sigma_min=0.1, sigma_max=20.0):
        self.sigma_min, self.sigma_max = sigma_min, sigma_max

    def log_mean_coeff(self, t):
        return -0.25 * ((1 - t) ** 2) * (self.sigma_max - self.sigma_min) - 0.5 * (1 - t) * self.sigma_min

    def d_log_mean_coeff(self, t):
        return 0.5 * (1 - t) * (self.sigma_max - self.sigma_min) + 0.5 * self.sigma_min

    def compute_alpha_t(self, t):
        a = th.exp(self.log_mean_coeff(t))
        return a, a * self.d_log_mean_coeff(t)

    def compute_sigma_t(self, t):
        a2 = th.exp(2 * self.log_mean_coeff(t))
        s = th.sqrt(1 - a2)
        return s, a2 * (2 * self.d_log_mean_coeff(t)) / (-2 * s)

    def compute_d_alpha_alpha_ratio_t(self, t):
        return self.d_log_mean_coeff(t)

    def compute_drift(self, x, t):
        t = expand_t_like_x(t, x)
        beta = self.sigma_min + (1 - t) * (self.sigma_max - self.sigma_min)
        return -0.5 * beta * x, beta / 2


class GVPCPlan(ICPlan):
    """Trigonometric (generalised VP) path: alpha = sin(pi t / 2), sigma = cos(pi t / 2)."""

    def compute_alpha_t(self, t):
        return th.sin(t * math.pi / 2), math.pi / 2 * th.cos(t * math.pi / 2)

    def compute_sigma_t(self, t):
        return th.cos(t * math.pi / 2), -math.pi / 2 * th.sin(t * math.pi / 2)

    def compute_d_alpha_alpha_ratio_t(self, t):
        return math.pi / (2 * th.tan(t * math.pi / 2))
