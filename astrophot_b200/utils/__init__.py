"""Small host-side helpers for building synthetic inputs (the reference keeps
equivalents under `utils/initialize/construct_psf.py:8-65`)."""
import numpy as np

__all__ = ["gaussian_psf", "moffat_psf", "initialize", "conversions", "parametric_profiles", "angle_operations", "optimization"]


def _grid(img_width, pixelscale, upsample):
    assert img_width % 2 == 1, "psf images should have an odd shape"
    n = img_width * upsample
    half = (img_width * pixelscale) / 2
    edges = np.linspace(-half, half, n + 1)
    mids = 0.5 * (edges[1:] + edges[:-1])
    return np.meshgrid(mids, mids, indexing="xy")


def _bin(z, img_width, upsample):
    return z.reshape(img_width, upsample, img_width, upsample).sum(axis=(1, 3))


def gaussian_psf(sigma, img_width, pixelscale, upsample=4, normalize=True):
    """Pixel-integrated (mean of ``upsample``^2 sub-samples) circular Gaussian, normalised to unit sum by default."""
    X, Y = _grid(img_width, pixelscale, upsample)
    z = _bin(np.exp(-0.5 * (X**2 + Y**2) / sigma**2), img_width, upsample)
    return z / z.sum() if normalize else z / upsample**2


def moffat_psf(n, Rd, img_width, pixelscale, upsample=4, normalize=True):
    """Pixel-integrated circular Moffat, normalised to unit sum by default."""
    X, Y = _grid(img_width, pixelscale, upsample)
    z = _bin(1.0 / (1.0 + (X**2 + Y**2) / Rd**2) ** n, img_width, upsample)
    return z / z.sum() if normalize else z / upsample**2


from . import angle_operations, conversions, initialize, optimization, parametric_profiles  # noqa: E402  (ap.utils.<module>.*, as in the reference)

initialize.gaussian_psf, initialize.moffat_psf = gaussian_psf, moffat_psf      # where the reference keeps them
