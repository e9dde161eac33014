"""Stencil coefficients of a stationary kernel (host side, once per kernel object).

Restates ``get_coeffs`` / ``coverage_diff`` / ``binary_search`` / ``DiscretizedKernelFN`` of the reference
(gpytorch_lattice_kernel/bilateral_kernel.py:14-56, 162-181) and its kernel profiles ``rbf``, ``Matern``,
``matern`` (:202-245).  The recipe: sample the profile ``k(t^2)`` on 10^4 points of [-30, 30]; find, by
bisection on the sample spacing ``s`` in (0.1, 9) to 1e-4, the spacing at which the fraction of the profile's
mass inside the stencil's spatial support ``|t| <= s(2r+1)/2`` equals the fraction of its spectrum below the
Nyquist frequency ``pi/s``; the coefficients are ``k((s j)^2) / k(0)``, ``j = -r..r``.

The numbers are independent of the lengthscale and of the data, so this stays in NumPy/PyTorch on the host.
Every floating-point operation is performed in the same precision and order as the reference so that the
coefficients (and through ``variance(coeffs)`` every lattice key) agree to the bit; tests/golden holds the
reference's values.
"""
from __future__ import annotations

import numpy as np
import torch
import torch.nn as nn
from torch.autograd import Function

__all__ = ["get_coeffs", "DiscretizedKernelFN", "rbf", "matern", "Matern", "StencilSearch"]

_NUM_SAMPLES = 10 ** 4
_HALF_WIDTH = 30
_SPACING_BOUNDS = (0.1, 9)
_SPACING_TOL = 1e-4
_MAX_BISECTIONS = 500


class StencilSearch:
    """Coverage-matching search for one profile ``profile(t)`` (a function of the distance ``t``, not ``t^2``)."""

    def __init__(self, profile):
        self.profile = profile
        n = _NUM_SAMPLES
        self.t = np.linspace(-_HALF_WIDTH, _HALF_WIDTH, n)
        self.k_t = profile(torch.from_numpy(self.t).float()).cpu().data.numpy()
        self.omega = 2 * np.pi * np.fft.fftfreq(n, 2 * _HALF_WIDTH / n)
        self.k_omega = np.absolute(np.fft.fft(self.k_t) / (2 * np.pi * np.sqrt(n)))
        self._k_t_mass = self.k_t.sum()
        self._k_omega_mass = self.k_omega.sum()

    def coverage_gap(self, spacing: float, order: int) -> float:
        """spatial coverage minus spectral coverage for sample spacing ``spacing`` (bilateral_kernel.py:30-39)."""
        taps = 2 * order + 1
        half_support = spacing * taps / 2
        nyquist = np.pi / spacing
        inside_t = (-half_support <= self.t) & (self.t <= half_support)
        inside_w = (-nyquist <= self.omega) & (self.omega <= nyquist)
        spatial = self.k_t[inside_t].sum() / self.k_t.sum()
        spectral = self.k_omega[inside_w].sum() / self.k_omega.sum()
        return spatial - spectral

    def spacing(self, order: int) -> float:
        """Zero of ``coverage_gap`` by bisection (bilateral_kernel.py:41-56)."""
        lo, hi = _SPACING_BOUNDS
        steps = 0
        while hi - lo > _SPACING_TOL:
            mid = (hi + lo) / 2
            if self.coverage_gap(mid, order) < 0:
                lo = mid
            else:
                hi = mid
            steps += 1
            if steps > _MAX_BISECTIONS:
                raise RuntimeError("stencil spacing search did not converge")
        return (hi + lo) / 2

    def coefficients(self, order: int) -> torch.Tensor:
        s = self.spacing(order)
        taps = self.profile(s * torch.arange(-order, order + 1).float())
        return taps / taps[order]


def get_coeffs(kernel_fn, order: int) -> torch.Tensor:
    """Discrete filter coefficients ``float32[2*order+1]`` of the profile ``kernel_fn(t)`` (bilateral_kernel.py:14-28)."""
    return StencilSearch(kernel_fn).coefficients(order)


class DiscretizedKernelFN(nn.Module):
    """Forward and derivative stencils of a kernel given as a function of the squared distance
    (bilateral_kernel.py:162-181).  ``get_coeffs()`` feeds the forward MVM, ``get_deriv_coeffs()`` the
    lengthscale-gradient filter."""

    def __init__(self, kernel_fn, order: int, verbose: bool = False):
        super().__init__()
        self.kernel_fn = kernel_fn
        self.order = order

        def profile(t):
            return self.kernel_fn(t ** 2)

        def d_profile(t):
            # derivative of the kernel with respect to the squared distance, evaluated at t^2
            with torch.autograd.enable_grad():
                z = t ** 2 + torch.zeros_like(t, requires_grad=True)
                (g,) = torch.autograd.grad(self.kernel_fn(z).sum(), z)
            return g

        self._forward_coeffs = get_coeffs(profile, order).detach()
        self._deriv_coeffs = get_coeffs(d_profile, order).detach()
        if verbose:
            print(f"Discretized kernel coeffs: {self._forward_coeffs}")
            print(f"Discretized kernel deriv coeffs: {self._deriv_coeffs}")

    def get_coeffs(self) -> torch.Tensor:
        return self._forward_coeffs

    def get_deriv_coeffs(self) -> torch.Tensor:
        return self._deriv_coeffs


def rbf(d2: torch.Tensor) -> torch.Tensor:
    """RBF profile as a function of the squared distance (bilateral_kernel.py:202-203)."""
    return torch.exp(-d2)


class Matern(Function):
    """Matern-nu profile in the squared distance with an explicit derivative that is finite at 0
    (bilateral_kernel.py:207-232).  nu in {1.5, 2.5}."""

    @staticmethod
    def forward(ctx, d2, nu):
        dist = d2.abs().sqrt()
        decay = torch.exp(-np.sqrt(nu * 2) * dist)
        if nu == 1.5:
            poly = (np.sqrt(3) * dist).add(1)
        elif nu == 2.5:
            poly = (np.sqrt(5) * dist).add(1).add(5.0 / 3.0 * dist ** 2)
        else:
            raise NotImplementedError
        if any(ctx.needs_input_grad):
            ctx.nu = nu
            ctx.save_for_backward(dist, decay)
        return poly * decay

    @staticmethod
    def backward(ctx, grad_output):
        if ctx.needs_input_grad[1]:
            raise NotImplementedError  # no gradient with respect to nu
        dist, decay = ctx.saved_tensors
        if ctx.nu == 1.5:
            slope = -(3 / 2)
        elif ctx.nu == 2.5:
            slope = -(5 / 6) * (1 + dist * np.sqrt(5))
        else:
            raise NotImplementedError
        return grad_output * slope * decay, None


def matern(d2: torch.Tensor, nu: float = 0.5) -> torch.Tensor:
    """Plain (autograd-differentiated) Matern profile, nu in {0.5, 1.5, 2.5} (bilateral_kernel.py:234-245)."""
    dist = d2.abs().sqrt()
    decay = torch.exp(-np.sqrt(nu * 2) * dist)
    if nu == 0.5:
        poly = 1
    elif nu == 1.5:
        poly = (np.sqrt(3) * dist).add(1)
    elif nu == 2.5:
        poly = (np.sqrt(5) * dist).add(1).add(5.0 / 3.0 * dist ** 2)
    else:
        raise NotImplementedError
    return poly * decay

