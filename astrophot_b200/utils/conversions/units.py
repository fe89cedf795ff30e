"""Flux <-> magnitude / surface brightness (reference: `utils/conversions/units.py:6-128`)."""
import numpy as np

deg_to_arcsec = 3600.0
arcsec_to_deg = 1.0 / deg_to_arcsec
_K = 2.5 / np.log(10)


def flux_to_mag(flux, zeropoint, fluxe=None):
    mag = zeropoint - 2.5 * np.log10(flux)
    return mag if fluxe is None else (mag, _K * fluxe / flux)


def mag_to_flux(mag, zeropoint, mage=None):
    flux = 10 ** ((zeropoint - mag) / 2.5)
    return flux if mage is None else (flux, flux * mage / _K)


def flux_to_sb(flux, pixel_area, zeropoint):
    """mag / arcsec^2 of ``flux`` collected over ``pixel_area``."""
    return flux_to_mag(flux, zeropoint) + 2.5 * np.log10(pixel_area)


def sb_to_flux(sb, pixel_area, zeropoint):
    return pixel_area * mag_to_flux(sb, zeropoint)


def _ellipse_area(a, b, R, A):
    if R is not None:
        return np.pi * R**2
    if A is not None:
        return A
    assert a is not None and b is not None, "give an area A, a radius R or semi-axes a, b"
    return np.pi * a * b


def magperarcsec2_to_mag(mu, a=None, b=None, A=None):
    """Total magnitude of a uniform surface brightness ``mu`` over an area (or ellipse a, b)."""
    return mu - 2.5 * np.log10(_ellipse_area(a, b, None, A))


def mag_to_magperarcsec2(m, a=None, b=None, R=None, A=None):
    return m + 2.5 * np.log10(_ellipse_area(a, b, R, A))


def PA_shift_convention(pa, unit="rad"):
    """Between position angles measured from the x axis and from the y axis (mod half a turn)."""
    half_turn = {"rad": np.pi, "deg": 180.0}[unit]
    return (pa - half_turn / 2) % half_turn
