"""Rotation and axis-ratio stretch of plane coordinates (reference: `utils/conversions/coordinates.py:5-56`);
numpy or torch inputs."""
import numpy as np
import torch


def _sc(theta):
    if isinstance(theta, torch.Tensor):
        return torch.sin(theta), torch.cos(theta)
    return np.sin(theta), np.cos(theta)


def Rotate_Cartesian(theta, X, Y=None):
    """Rotate (X, Y) counter-clockwise by ``theta``; ``X`` may hold both coordinates when ``Y`` is omitted."""
    if Y is None:
        X, Y = X[0], X[1]
    s, c = _sc(theta)
    return c * X - s * Y, s * X + c * Y


Rotate_Cartesian_np = Rotate_Cartesian


def Axis_Ratio_Cartesian(q, X, Y, theta=0.0, inv_scale=False):
    """R(theta) diag(1, f) R(-theta) applied to (X, Y), f = q (or 1/q with ``inv_scale``): the component along the
    direction theta + 90 deg is scaled by f."""
    f = (1 / q if inv_scale else q) - 1
    s, c = _sc(theta)
    half_s2 = f * s * c
    return (1 + f * s * s) * X - half_s2 * Y, -half_s2 * X + (1 + f * c * c) * Y


Axis_Ratio_Cartesian_np = Axis_Ratio_Cartesian
