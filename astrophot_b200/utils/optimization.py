"""chi^2 of two arrays (reference: `utils/optimization.py:6-48`).  A convenience for scripts on tensors they already
hold; the fit's own chi^2 is computed on the device (``apb_chi2``)."""
import torch


def chi_squared(target, model, mask=None, variance=None):
    """Sum of squared residuals, divided by ``variance`` when given, over the pixels where ``mask`` is False."""
    r2 = (target - model) ** 2
    if variance is not None:
        r2 = r2 / variance
    if mask is not None:
        r2 = r2[torch.logical_not(mask)]
    return torch.sum(r2)


def reduced_chi_squared(target, model, params, mask=None, variance=None):
    """chi^2 per degree of freedom: unmasked pixels minus ``params``."""
    n = target.numel() if mask is None else torch.sum(torch.logical_not(mask))
    return chi_squared(target, model, mask, variance) / (n - params)
