"""Closed forms of the Sersic and Moffat profiles (reference: `utils/conversions/functions.py:7-237`): b(n), total flux
<-> central / effective intensity, inverse profile.  One implementation serves numpy and torch inputs; the ``_np`` /
``_torch`` names of the reference are aliases."""
import numpy as np
import torch
from scipy.special import gammaln as _gammaln_np


def _is_torch(*xs):
    return any(isinstance(x, torch.Tensor) for x in xs)


def _gamma(x):
    if isinstance(x, torch.Tensor):
        return torch.exp(torch.lgamma(x))
    return np.exp(_gammaln_np(x))


def _exp(x):
    return torch.exp(x) if isinstance(x, torch.Tensor) else np.exp(x)


def _log(x):
    return torch.log(x) if isinstance(x, torch.Tensor) else np.log(x)


def sersic_n_to_b(n):
    """b(n) such that Re encloses half of the light (asymptotic series in 1/n)."""
    x = 1 / n
    return 2 * n - 1 / 3 + x * (4 / 405 + x * (46 / 25515 + x * (131 / 1148175 - x * 2194697 / 30690717750)))


def _central_norm(n, R, q):
    # integral of exp(-(r/R)^(1/n)) over the plane, axis ratio q
    return 2 * np.pi * q * n * R**2 * _gamma(2 * n)


def _effective_norm(n, R, q):
    b = sersic_n_to_b(n)
    return _central_norm(n, R, q) * (_exp(b) * b ** (-2 * n))


def sersic_I0_to_flux(I0, n, R, q):
    return I0 * _central_norm(n, R, q)


def sersic_flux_to_I0(flux, n, R, q):
    return flux / _central_norm(n, R, q)


def sersic_Ie_to_flux(Ie, n, R, q):
    return Ie * _effective_norm(n, R, q)


def sersic_flux_to_Ie(flux, n, R, q):
    return flux / _effective_norm(n, R, q)


def sersic_inv(I, n, Re, Ie):
    """Radius at which a Sersic profile has intensity ``I``."""
    return Re * (1 - _log(I / Ie) / sersic_n_to_b(n)) ** n


sersic_I0_to_flux_np = sersic_I0_to_flux_torch = sersic_I0_to_flux
sersic_flux_to_I0_np = sersic_flux_to_I0_torch = sersic_flux_to_I0
sersic_Ie_to_flux_np = sersic_Ie_to_flux_torch = sersic_Ie_to_flux
sersic_flux_to_Ie_np = sersic_flux_to_Ie_torch = sersic_flux_to_Ie
sersic_inv_np = sersic_inv_torch = sersic_inv


def moffat_I0_to_flux(I0, n, rd, q):
    """Total flux of I0 / (1 + (r/rd)^2)^n  (n > 1)."""
    return I0 * np.pi * rd**2 * q / (n - 1)
