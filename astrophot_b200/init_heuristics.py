"""Start values from the data: the input prep that feeds the fit.

Host-side numpy; nothing here is on the device path.  ``models.py`` calls these from the ``initialize()`` methods so
that a reference script (``model.initialize(); ap.fit.LM(model).fit()``) runs unchanged.  The recipes are the
reference's, restated -- including two of its quirks, kept because the start values are part of what a fit reproduces:

* centroid search on a small box                 `utils/initialize/center.py:8-58`
* position angle from the second angular moment  `utils/angle_operations.py:35-56`
* axis ratio from the m=2 Fourier amplitude of isophotes  `utils/initialize/initialize.py:8-113`,
  `utils/isophote/extract.py:9-27,105-225`, `utils/interpolate.py:232-273`
* radial profile in log-spaced bins + Nelder-Mead fit of the profile  `models/_shared_methods.py:70-164`

Quirk 1: the isophote centre is handed over as (pixel-y, pixel-x) and then used as (x, y)
(`galaxy_model_object.py:103-106` -> `initialize.py:40`).  Quirk 2: the exponential profile used for the start-value fit
takes its arguments swapped (`exponential_model.py:36-37` vs `parametric_profiles.py:77`).
"""
import numpy as np
from scipy.optimize import minimize
from scipy.stats import binned_statistic

from . import AP_config

LN10 = np.log(10.0)


def half_spread(v, lo=16.0, hi=84.0):
    """Half of the lo..hi percentile range: the robust sigma the reference uses everywhere
    (``scipy.stats.iqr(v, rng=(16, 84)) / 2``)."""
    a, b = np.percentile(v, [lo, hi])
    return (b - a) / 2


def edge_pixels(dat):
    return np.concatenate((dat[:, 0], dat[:, -1], dat[0, :], dat[-1, :]))


# ---------------------------------------------------------------------------
# centre
# ---------------------------------------------------------------------------
def light_centroid(start, img, box=None):
    """Iterated centre of light in a small box around ``start`` = (row, col); gives up (returns ``start``) when the box
    reaches the image edge.  Returns the centre the last box was placed on."""
    if box is None:
        box = max(min(int(min(img.shape) / 10), 30), 6)
    box += box % 2
    cc, rr = np.meshgrid(np.arange(box), np.arange(box))
    c = start
    for _ in range(100):
        r0, r1 = int(round(c[0]) - box / 2), int(round(c[0]) + box / 2)
        c0, c1 = int(round(c[1]) - box / 2), int(round(c[1]) + box / 2)
        if r0 < 0 or c0 < 0 or r1 >= img.shape[0] or c1 >= img.shape[1]:
            AP_config.ap_logger.warning("Image edge!")
            return start
        cut = img[r0:r1, c0:c1]
        tot = np.sum(cut)
        new = np.array([r0 + np.sum(cut * rr) / tot, c0 + np.sum(cut * cc) / tot])
        if np.sum(np.abs(np.array(c) - new)) < 0.1:
            break
        c = new
    return c


# ---------------------------------------------------------------------------
# position angle, axis ratio
# ---------------------------------------------------------------------------
def moment_position_angle(flux, X, Y):
    """Angle (mod pi) of the flux-weighted second angular moment."""
    th2 = 2 * np.arctan2(Y, X)
    tot = np.sum(flux)
    return np.arctan2(np.sum(flux * np.sin(th2)) / tot, np.sum(flux * np.cos(th2)) / tot) / 2 % np.pi


def upper_clip_limit(v, iterations=10, nsigma=5):
    """Median + nsigma robust sigmas, iterated on the values below the previous limit."""
    v = np.sort(v)
    lim, old = np.inf, 0
    k = 0
    while k < iterations and old != lim:
        kept = v[v < lim]
        old = lim
        lim = np.median(kept) + half_spread(kept) * nsigma
        k += 1
    return lim


def lanczos_at(img, X, Y, a):
    """Lanczos-``a`` interpolation of ``img`` at the points (X, Y) (pixel units); taps outside the image are dropped
    and the weights renormalised."""
    H, W = img.shape
    k = np.arange(-a + 1, a + 1)
    fx, fy = np.floor(X), np.floor(Y)

    def taps(f, pos, n):
        t = (k[None, :] - pos[:, None]) + f[:, None]
        idx = f[:, None].astype(np.int64) + k[None, :]
        ok = (idx >= 0) & (idx < n)
        return np.where(ok, np.sinc(t) * np.sinc(t / a), 0.0), np.clip(idx, 0, n - 1)

    Lx, ix = taps(fx, X, W)
    Ly, iy = taps(fy, Y, H)
    vals = img[iy[:, :, None], ix[:, None, :]]                       # (N, 2a, 2a)
    num = np.einsum("nab,na,nb->n", vals, Ly, Lx)
    return num / (Lx.sum(axis=1) * Ly.sum(axis=1))


def isophote_samples(img, sma, q, pa, cx, cy, clip_nsigma=None, fill_clipped=False):
    """Flux along the ellipse (semi-major axis ``sma`` pixels, axis ratio ``q``, angle ``pa``) about (cx, cy):
    Lanczos-5 below 30 px, nearest pixel above; optional upper sigma clip whose rejected samples are either dropped or
    re-filled by periodic interpolation in angle."""
    n = max(15, int(0.9 * 2 * np.pi * sma))
    th = np.linspace(0, 2 * np.pi * (1.0 - 1.0 / n), n)
    th = np.arctan(q * np.tan(th)) + np.pi * (np.cos(th) < 0)
    rs = q / (np.abs(q * np.cos(th)) ** 2 + np.abs(np.sin(th)) ** 2) ** (1.0 / 2)
    ex, ey = sma * (rs * np.cos(th)), sma * (rs * np.sin(th))
    s, c = np.sin(pa), np.cos(pa)
    X, Y = c * ex - s * ey + cx, c * ey + s * ex + cy
    th = (th + pa) % (2 * np.pi)
    inside = (X >= 0) & (X < img.shape[1] - 1) & (Y >= 0) & (Y < img.shape[0] - 1)
    X, Y, th = X[inside], Y[inside], th[inside]
    if sma < 30:
        flux = lanczos_at(img, X, Y, 5)
    else:
        flux = img[np.rint(Y).astype(np.int32), np.rint(X).astype(np.int32)]
    if clip_nsigma is not None and len(flux) > 30:
        keep = flux < upper_clip_limit(flux, 10, clip_nsigma)
        if np.sum(keep) <= 0:
            AP_config.ap_logger.warning("Entire Isophote was Masked!")
        elif fill_clipped:
            flux[~keep] = np.interp(th[~keep], th[keep], flux[keep], period=2 * np.pi)
        else:
            flux = flux[keep]
    return flux


def axis_ratio_scan(img, cx, cy, threshold, pa, q_samples):
    """The q of ``q_samples`` whose isophote (at the radius where the light drops to ``threshold``) has the smallest
    m=2 Fourier amplitude relative to its flux."""
    radii = [1.0]
    while radii[-1] < max(img.shape) / 2:
        radii.append(radii[-1] * 1.2)
        iso = isophote_samples(img, radii[-1], np.max(q_samples), pa, cx, cy, clip_nsigma=3)
        if len(iso) < 3:
            continue
        if np.quantile(iso, 0.8) < threshold and len(radii) > 4:
            break
    R = radii[-1]
    amp2 = []
    for q in q_samples:
        iso = isophote_samples(img, R, q, pa, cx, cy, clip_nsigma=3, fill_clipped=True)
        if len(iso) < 3:
            amp2.append(None)
            continue
        coef = np.fft.fft(iso)
        amp2.append(np.abs(coef[2]) / (len(iso) * (max(0, np.median(iso)) + half_spread(iso))))
    good = [a for a in amp2 if a is not None]
    if not good:
        raise ValueError("Unable to recover any isophotes, try on a better band or manually provide values")
    amp2 = [good[-1] if a is None else a for a in amp2]
    return q_samples[int(np.argmin(amp2))]


# ---------------------------------------------------------------------------
# radial profile and profile fit
# ---------------------------------------------------------------------------
def radial_profile(dat, mask, R, pixel_area, rad_bins=None):
    """Median surface brightness (log10) and its robust scatter in radial bins of the pixel radii ``R``; the median of
    the edge pixels is taken off first, non-positive bins are floored and the outer bins forced to decline."""
    dat = dat.copy()
    if mask is not None:
        dat[mask] = np.median(dat[~mask])
    dat -= np.median(edge_pixels(dat))
    R = R.ravel()
    if rad_bins is None:
        rad_bins = np.logspace(np.log10(R.min() * 0.9), np.log10(R.max() * 1.1), 11)
    else:
        rad_bins = np.array(rad_bins)
    flat = dat.ravel()
    I = binned_statistic(R, flat, statistic="median", bins=rad_bins)[0] / pixel_area
    S = binned_statistic(R, flat, statistic=lambda d: half_spread(d), bins=rad_bins)[0] / pixel_area
    Rm = (rad_bins[:-1] + rad_bins[1:]) / 2
    I[I <= 0] = np.min(I[np.logical_and(np.isfinite(I), I > 0)])
    for i in range(5, len(I)):
        if I[i] >= I[i - 1] and np.isfinite(I[i - 1]):
            I[i] = I[i - 1] - np.abs(I[i - 1] * 0.1)
    S = S / (I * LN10)
    I = np.log10(I)
    ok = np.isfinite(I)
    if not np.all(ok):
        I[~ok] = np.interp(Rm[~ok], Rm[ok], I[ok])
    ok = np.isfinite(S)
    if not np.all(ok):
        S[~ok] = np.abs(np.interp(Rm[~ok], Rm[ok], S[ok]))
    return Rm, I, S


def _sersic_b(n):
    return (2 * n - 1 / 3 + 4 / (405 * n) + 46 / (25515 * n**2) + 131 / (1148175 * n**3)
            - 2194697 / (30690717750 * n**4))


def _p_sersic(R, n, Re, Ie):
    if n <= 0 or Re <= 0 or 10**Ie <= 0:
        return np.ones(len(R)) * 1e6
    return 10**Ie * np.exp(-_sersic_b(n) * ((R / Re) ** (1 / n) - 1))


def _p_exponential(R, Re, Ie):
    # quirk 2 (module docstring): the reference's fit evaluates  Re * exp(-b1 (R / 10^Ie - 1))
    return Re * np.exp(-_sersic_b(1.0) * (R / 10**Ie - 1.0))


def _p_gaussian(R, sigma, flux):
    return (10**flux / np.sqrt(2 * np.pi * sigma**2)) * np.exp(-0.5 * ((R / sigma) ** 2))


def _p_moffat(R, n, Rd, I0):
    return 10**I0 / (1 + (R / Rd) ** 2) ** n


# kind -> (parameter names, profile, start values from the binned profile)
PROFILE_FITS = {
    "sersic": (("n", "Re", "Ie"), _p_sersic, lambda R, I: (2.0, R[4], I[4])),
    "exponential": (("Re", "Ie"), _p_exponential, lambda R, I: (R[4], I[4])),
    "gaussian": (("sigma", "flux"), _p_gaussian, lambda R, I: (R[4], I[0])),
    "moffat": (("n", "Rd", "I0"), _p_moffat, lambda R, I: (2.0, R[4], I[0])),
}


def fit_profile(R, I, prof, x0):
    """Nelder-Mead fit of log10(profile) to the binned profile, the two worst bins ignored; ten bootstrap refits (drawn
    from numpy's global generator, as the reference does) give the start uncertainties.
    Returns (x, success, per-parameter std of the bootstrap fits)."""
    def cost(x, r, f):
        res = (f - np.log10(prof(r, *x))) ** 2
        return np.mean(np.sort(res)[:-2])

    with np.errstate(all="ignore"):
        best = minimize(cost, x0=x0, args=(R, I), method="Nelder-Mead")
        boots = []
        for _ in range(10):
            pick = np.random.randint(0, len(R), len(R))
            boots.append(minimize(cost, x0=x0, args=(R[pick], I[pick]), method="Nelder-Mead").x)
    return best.x, bool(best.success), np.std(np.array(boots), axis=0)
