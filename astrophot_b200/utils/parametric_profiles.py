"""The radial profiles as plain functions of radius (reference: `utils/parametric_profiles.py:7-177`) for building
synthetic data and plotting; the device evaluates its own copies (csrc/apb_internal.cuh ``eval_point``).  One
implementation for numpy and torch inputs; ``_np`` / ``_torch`` are aliases, except where the reference's two
signatures differ (``exponential_np`` takes (R, Ie, Re))."""
import numpy as np
import torch

from .conversions.functions import sersic_n_to_b


def _exp(x):
    return torch.exp(x) if isinstance(x, torch.Tensor) else np.exp(x)


def sersic_torch(R, n, Re, Ie):
    return Ie * _exp(-sersic_n_to_b(n) * ((R / Re) ** (1 / n) - 1))


def sersic_np(R, n, Re, Ie):
    """As ``sersic_torch``; non-positive parameters give a large constant (a wall for optimisers)."""
    if np.any(np.array([n, Re, Ie]) <= 0):
        return np.ones(len(R)) * 1e6
    return sersic_torch(R, n, Re, Ie)


def gaussian_torch(R, sigma, I0):
    return (I0 / (2 * np.pi * sigma**2) ** 0.5) * _exp(-0.5 * (R / sigma) ** 2)


gaussian_np = gaussian_torch


def exponential_torch(R, Re, Ie):
    return Ie * _exp(-sersic_n_to_b(1.0) * (R / Re - 1.0))


def exponential_np(R, Ie, Re):
    return exponential_torch(R, Re, Ie)


def moffat_torch(R, n, Rd, I0):
    return I0 / (1 + (R / Rd) ** 2) ** n


moffat_np = moffat_torch


def nuker_torch(R, Rb, Ib, alpha, beta, gamma):
    x = R / Rb
    return Ib * 2 ** ((beta - gamma) / alpha) * x ** (-gamma) * (1 + x**alpha) ** ((gamma - beta) / alpha)


nuker_np = nuker_torch
