"""Statistics of angles through unit phasors (reference: `utils/angle_operations.py:5-56`)."""
import numpy as np

from ..init_heuristics import moment_position_angle


def _phasor(a):
    return np.exp(1j * np.asarray(a, dtype=np.float64))


def Angle_Average(a):
    """Direction of the mean phasor."""
    return np.angle(np.mean(_phasor(a)))


def Angle_Median(a):
    """Direction of (median cos, median sin)."""
    z = _phasor(a)
    return np.angle(np.median(z.real) + 1j * np.median(z.imag))


def Angle_Scatter(a):
    """16-84 percentile range of the angles about their mean direction (measured from a quarter turn away from it, so
    that the wrap falls opposite the mean)."""
    z = _phasor(a)
    lo, hi = np.percentile(np.angle(1j * z / np.mean(z)), [16, 84])
    return hi - lo


def Angle_COM_PA(flux, X=None, Y=None):
    """Position angle (mod pi) of the flux-weighted second angular moment; coordinates default to pixel indices about
    the middle of the array."""
    if X is None:
        h, w = flux.shape
        X, Y = np.meshgrid(np.arange(w) - w / 2, np.arange(h) - h / 2, indexing="xy")
    return moment_position_angle(flux, X, Y)
