"""CPU oracle for the AstroPhot forward-model-and-fit hot path.

TEST INFRASTRUCTURE ONLY.  Nothing under ``astrophot_b200/`` imports this
module; only ``tests/``, ``__graft_entry__.smoke()`` and the ``cpu_baseline`` /
``--impl reference`` legs of ``bench.py`` may.  It is a plain numpy restatement
of the reference's algorithm, written from SURVEY.md Appendix B and the
reference sources cited next to each function, operating on the same flat
``Scene`` tables (``astrophot_b200/scene.py``) that are handed to the CUDA
library.

Parity status: PINNED.  ``oracle/make_golden.py`` imports the reference
itself (``/root/reference/astrophot``, CPU torch) in the build container, runs
it on seeded synthetic scenes and stores its outputs under ``tests/golden/``;
``tests/test_oracle_golden.py`` checks this oracle against those fixtures
(model images, forward-AD Jacobians, LM histories).

Derivatives are analytic (forward-mode by hand); the reference obtains them
with ``torch.func`` forward-mode AD over the same arithmetic
(`models/_model_methods.py:325-340`), so the two agree to rounding.
"""
import math

import numpy as np
from numpy.polynomial.legendre import leggauss

from astrophot_b200 import scene as sc

LN10 = math.log(10.0)

# Host threads for the timed CPU legs of bench.py (numpy releases the GIL inside its array loops): sources of a
# scene and the planes of a PSF convolution are independent.  1 (the default, used by the tests) = plain loops.
N_THREADS = 1


def set_threads(n):
    global N_THREADS
    N_THREADS = max(1, int(n))


def _pmap(fn, items):
    items = list(items)
    if N_THREADS <= 1 or len(items) <= 1:
        return [fn(it) for it in items]
    from concurrent.futures import ThreadPoolExecutor
    with ThreadPoolExecutor(max_workers=min(N_THREADS, len(items))) as ex:
        return list(ex.map(fn, items))


def _host(a, dtype=np.float64):
    """numpy view of an array or (possibly CUDA) torch tensor."""
    if hasattr(a, "detach"):
        a = a.detach().cpu().numpy()
    return np.asarray(a, dtype=dtype)


# ---------------------------------------------------------------------------
# parameters: representation <-> value  (utils/conversions/optimization.py:6-54)
# ---------------------------------------------------------------------------
def rep_to_val(x, transform, lo, hi):
    """Returns (values, dvalue/drep) for a representation-space vector."""
    x = np.asarray(x, dtype=np.float64)
    val = x.copy()
    dv = np.ones_like(x)
    for p in range(len(x)):
        t = transform[p]
        if t == sc.TR_LOWER:
            d = x[p] - lo[p]
            rt = math.sqrt(d * d + 4)
            val[p] = 0.5 * (x[p] + lo[p] + rt)
            dv[p] = 0.5 + 0.5 * d / rt
        elif t == sc.TR_UPPER:
            d = x[p] - hi[p]
            rt = math.sqrt(d * d + 4)
            val[p] = 0.5 * (x[p] + hi[p] - rt)
            dv[p] = 0.5 - 0.5 * d / rt
        elif t == sc.TR_BOTH:
            val[p] = (math.atan(x[p]) + math.pi / 2) * (hi[p] - lo[p]) / math.pi + lo[p]
            dv[p] = (hi[p] - lo[p]) / (math.pi * (x[p] * x[p] + 1))
        elif t == sc.TR_CYCLIC:
            val[p] = lo[p] + np.remainder(x[p] - lo[p], hi[p] - lo[p])
    return val, dv


def val_to_rep(v, transform, lo, hi):
    v = np.asarray(v, dtype=np.float64)
    out = v.copy()
    for p in range(len(v)):
        t = transform[p]
        if t == sc.TR_LOWER:
            out[p] = v[p] - 1.0 / (v[p] - lo[p])
        elif t == sc.TR_UPPER:
            out[p] = v[p] - 1.0 / (v[p] - hi[p])
        elif t == sc.TR_BOTH:
            out[p] = math.tan((v[p] - lo[p]) * math.pi / (hi[p] - lo[p]) - math.pi / 2)
        elif t == sc.TR_CYCLIC:
            out[p] = lo[p] + np.remainder(v[p] - lo[p], hi[p] - lo[p])
    return out


def source_elements(src, vals):
    """Natural-unit element values of one source from the global value vector."""
    return np.array([vals[s] if s >= 0 else c for s, c in zip(src.slot, src.cval)], dtype=np.float64)


# ---------------------------------------------------------------------------
# profiles  (utils/parametric_profiles.py:7-100,157-177; conversions/functions.py:7-22)
# ---------------------------------------------------------------------------
def sersic_b(n):
    return (2 * n - 1 / 3 + 4 / (405 * n) + 46 / (25515 * n**2) + 131 / (1148175 * n**3)
            - 2194697 / (30690717750 * n**4))


def sersic_db(n):
    return (2 - 4 / (405 * n**2) - 92 / (25515 * n**3) - 393 / (1148175 * n**4)
            + 4 * 2194697 / (30690717750 * n**5))


def sersic_total_flux(Ie_lin, n, Re, q):
    """conversions/functions.py:168-190 (used as integration reference)."""
    bn = sersic_b(n)
    return 2 * math.pi * Ie_lin * Re**2 * q * n * (math.exp(bn) * bn ** (-2 * n)) * math.exp(math.lgamma(2 * n))


def _spline_eval(R, prof, v, extend=True):
    """log10 brightness s(R) and its derivatives wrt R and every node value.
    Hermite cubic, centred-difference slopes (utils/interpolate.py:31-62) with
    the reference's index arithmetic, linear extension past the last node
    (utils/parametric_profiles.py:157-177)."""
    prof = np.asarray(prof, dtype=np.float64)
    v = np.asarray(v, dtype=np.float64)
    K = len(prof)
    h = prof[1:] - prof[:-1]
    delta = (v[1:] - v[:-1]) / h
    m = np.concatenate([delta[[0]], (delta[1:] + delta[:-1]) / 2, delta[[-1]]])
    # dm/dv  (K x K)
    ddelta = np.zeros((K - 1, K))
    for k in range(K - 1):
        ddelta[k, k + 1] = 1 / h[k]
        ddelta[k, k] = -1 / h[k]
    dm = np.concatenate([ddelta[[0]], (ddelta[1:] + ddelta[:-1]) / 2, ddelta[[-1]]])
    Rf = R.reshape(-1)
    idx = np.searchsorted(prof[:-1], Rf, side="left") - 1   # -1 wraps like torch indexing
    i0 = np.mod(idx, K)
    i1 = idx + 1
    dx = prof[i1] - prof[i0]
    t = (Rf - prof[i0]) / dx
    h00 = 1 - 3 * t**2 + 2 * t**3
    h10 = t - 2 * t**2 + t**3
    h01 = 3 * t**2 - 2 * t**3
    h11 = -(t**2) + t**3
    s = h00 * v[i0] + h10 * m[i0] * dx + h01 * v[i1] + h11 * m[i1] * dx
    d00 = (-6 * t + 6 * t**2) / dx
    d10 = (1 - 4 * t + 3 * t**2) / dx
    d01 = (6 * t - 6 * t**2) / dx
    d11 = (-2 * t + 3 * t**2) / dx
    dsdR = d00 * v[i0] + d10 * m[i0] * dx + d01 * v[i1] + d11 * m[i1] * dx
    dsdv = np.zeros((K, Rf.size))
    np.add.at(dsdv, (i0, np.arange(Rf.size)), h00)
    np.add.at(dsdv, (i1, np.arange(Rf.size)), h01)
    dsdv += dm[i0].T * (h10 * dx) + dm[i1].T * (h11 * dx)
    beyond = Rf > prof[-1]
    if np.any(beyond):
        slope = (v[-1] - v[-2]) / (prof[-1] - prof[-2])
        s[beyond] = v[-2] + (Rf[beyond] - prof[-2]) * slope
        dsdR[beyond] = slope
        dsdv[:, beyond] = 0
        dsdv[K - 2, beyond] = 1 - (Rf[beyond] - prof[-2]) / (prof[-1] - prof[-2])
        dsdv[K - 1, beyond] = (Rf[beyond] - prof[-2]) / (prof[-1] - prof[-2])
    return s.reshape(R.shape), dsdR.reshape(R.shape), dsdv.reshape((K,) + R.shape), beyond.reshape(R.shape)


def eval_profile(src, el, X, Y, area, want_grad):
    """Brightness at plane offsets (X, Y) from the centre, and (optionally)
    its derivative with respect to every element (natural units).

    Follows `_shared_methods.py:286-307` (rotate by -(PA - pi/2), divide y by
    q), `_model_methods.py:38-40` (softened radius) and the radial profiles.
    Returns (I, dI) with dI of shape (n_elem,) + X.shape or None.
    """
    kind = src.kind
    ne = len(el)
    if kind == sc.KIND_FLAT_SKY:
        I = np.full(X.shape, area * 10.0 ** el[2])
        if not want_grad:
            return I, None
        dI = np.zeros((ne,) + X.shape)
        dI[2] = LN10 * I
        return I, dI
    if kind == sc.KIND_PLANE_SKY:
        # planesky_model.py:65-74:  pixel_area F + X dx + Y dy  (X, Y relative to the centre)
        I = area * el[2] + X * el[3] + Y * el[4]
        if not want_grad:
            return I, None
        dI = np.zeros((ne,) + X.shape)
        dI[0], dI[1] = -el[3], -el[4]
        dI[2], dI[3], dI[4] = area, X, Y
        return I, dI
    if kind == sc.KIND_POINT:
        raise ValueError("point sources are not profile-evaluated")
    q, PA = el[2], el[3]
    if src.flags & sc.FLAG_RADIAL:
        xp, yp = X, Y
        c = s = None
    else:
        theta = -(PA - math.pi / 2)
        s, c = math.sin(theta), math.cos(theta)
        xp = c * X - s * Y
        yp = (s * X + c * Y) / q
    R = np.sqrt(xp**2 + yp**2 + src.softening**2)
    dI = None
    if kind == sc.KIND_SERSIC:
        n, Re, Ie = el[4], el[5], el[6]
        bn = sersic_b(n)
        u = np.power(R / Re, 1 / n)
        I = (area * 10.0**Ie) * np.exp(-bn * (u - 1))
        if want_grad:
            dIdR = -I * bn * u / (n * R)
            dI = np.zeros((ne,) + X.shape)
            dI[4] = I * (-sersic_db(n) * (u - 1) + bn * u * np.log(R / Re) / n**2)
            dI[5] = I * bn * u / (n * Re)
            dI[6] = LN10 * I
    elif kind == sc.KIND_EXPONENTIAL:
        Re, Ie = el[4], el[5]
        b1 = sersic_b(1.0)
        I = (area * 10.0**Ie) * np.exp(-b1 * (R / Re - 1.0))
        if want_grad:
            dIdR = -I * b1 / Re
            dI = np.zeros((ne,) + X.shape)
            dI[4] = I * b1 * R / Re**2
            dI[5] = LN10 * I
    elif kind == sc.KIND_GAUSSIAN:
        sig, fl = el[4], el[5]
        I = ((area * 10.0**fl) / math.sqrt(2 * math.pi * sig**2)) * np.exp(-0.5 * (R / sig) ** 2)
        if want_grad:
            dIdR = -I * R / sig**2
            dI = np.zeros((ne,) + X.shape)
            dI[4] = I * (-1 / sig + R**2 / sig**3)
            dI[5] = LN10 * I
    elif kind == sc.KIND_MOFFAT:
        n, Rd, I0 = el[4], el[5], el[6]
        t = 1 + (R / Rd) ** 2
        I = (area * 10.0**I0) / t**n
        if want_grad:
            dIdR = -I * n * 2 * R / (Rd**2 * t)
            dI = np.zeros((ne,) + X.shape)
            dI[4] = -I * np.log(t)
            dI[5] = I * n * 2 * R**2 / (Rd**3 * t)
            dI[6] = LN10 * I
    elif kind == sc.KIND_SPLINE:
        sv, dsdR, dsdv, _ = _spline_eval(R, src.prof, el[4:])
        I = area * 10.0**sv
        if want_grad:
            dIdR = LN10 * I * dsdR
            dI = np.zeros((ne,) + X.shape)
            dI[4:] = LN10 * I * dsdv
    else:
        raise ValueError(f"unknown kind {kind}")
    if want_grad:
        if src.flags & sc.FLAG_RADIAL:
            dRdX, dRdY = xp / R, yp / R
        else:
            dRdX = (xp * c + yp * s / q) / R
            dRdY = (-xp * s + yp * c / q) / R
            dI[2] = dIdR * (-(yp**2) / (q * R))
            dI[3] = dIdR * (xp * yp * (q - 1 / q) / R)
        dI[0] = -dIdR * dRdX
        dI[1] = -dIdR * dRdY
    return I, dI


# ---------------------------------------------------------------------------
# quadrature tables (utils/operations.py:94-120)
# ---------------------------------------------------------------------------
def quad_nodes(n, S):
    """Offsets (dx, dy) in the plane and weights of the n x n Gauss-Legendre
    rule over one pixel with pixelscale matrix S."""
    a, w = leggauss(n)
    k = np.arange(n * n)
    ax, ay = a[k % n], a[k // n]
    off = S @ (np.stack((ax, ay)) / 2.0)
    W = (w[k // n] * w[k % n]) / 4.0
    return off[0], off[1], W


def sub_offsets(N, S):
    d = np.linspace(-(N - 1) / (2 * N), (N - 1) / (2 * N), N)
    k = np.arange(N * N)
    off = S @ np.stack((d[k % N], d[k // N]))
    return off[0], off[1]


def _gl(src, el, X, Y, S, area, n, want_grad):
    """Gauss-Legendre integral of each point's pixel; also the centre-node value."""
    ox, oy, W = quad_nodes(n, S)
    Xs = X[..., None] + ox
    Ys = Y[..., None] + oy
    I, dI = eval_profile(src, el, Xs, Ys, area, want_grad)
    ref = I[..., (n * n) // 2]
    res = (I * W).sum(axis=-1)
    dres = (dI * W).sum(axis=-1) if want_grad else None
    return res, ref, dres


def grid_integrate(src, el, X, Y, S, area, depth, reference, want_grad):
    """Adaptive sub-pixel integration (utils/operations.py:150-247).
    Returns integrated flux (and derivatives) for flat arrays X, Y; also counts
    the profile evaluations in ``grid_integrate.spe``."""
    res, ref, dres = _gl(src, el, X, Y, S, area, src.quad_level, want_grad)
    grid_integrate.spe += X.size * src.quad_level**2
    grid_integrate.queued[depth] = grid_integrate.queued.get(depth, 0) + X.size
    if depth >= src.max_depth:
        return res, dres
    select = np.abs(res - ref) > reference
    if not np.any(select):
        return res, dres
    N = src.gridding
    sx, sy = sub_offsets(N, S)
    subX = (X[select][:, None] + sx).reshape(-1)
    subY = (Y[select][:, None] + sy).reshape(-1)
    sres, sdres = grid_integrate(src, el, subX, subY, S / N, area / N**2, depth + 1,
                                 reference * N**2, want_grad)
    out = res.copy()
    out[select] = sres.reshape(-1, N * N).sum(axis=-1)
    if want_grad:
        dout = dres.copy()
        dout[:, select] = sdres.reshape(sdres.shape[0], -1, N * N).sum(axis=-1)
        return out, dout
    return out, None


grid_integrate.spe = 0
grid_integrate.queued = {}


# ---------------------------------------------------------------------------
# PSF shift  (_model_methods.py:187-230, utils/interpolate.py:282-329)
# ---------------------------------------------------------------------------
def shift_psf_bilinear(psf, shift, keep_pad=True, want_grad=False):
    """Bilinear resample of the 1-px zero-padded PSF at (i - sx, j - sy).
    Returns stamp (and d/dsx, d/dsy)."""
    im = np.pad(np.asarray(psf, dtype=np.float64), 1)
    h, w = im.shape
    x = np.arange(w, dtype=np.float64) - shift[0]
    y = np.arange(h, dtype=np.float64) - shift[1]
    x0 = np.floor(x).astype(np.int64)
    y0 = np.floor(y).astype(np.int64)
    x1 = np.clip(x0 + 1, 1, w - 1)
    y1 = np.clip(y0 + 1, 1, h - 1)
    x0 = np.clip(x0, 0, w - 2)
    y0 = np.clip(y0, 0, h - 2)
    wx0, wx1 = (x1 - x)[None, :], (x - x0)[None, :]
    wy0, wy1 = (y1 - y)[:, None], (y - y0)[:, None]
    fa = im[np.ix_(y0, x0)]
    fb = im[np.ix_(y1, x0)]
    fc = im[np.ix_(y0, x1)]
    fd = im[np.ix_(y1, x1)]
    out = fa * (wx0 * wy0) + fb * (wx0 * wy1) + fc * (wx1 * wy0) + fd * (wx1 * wy1)
    grads = None
    if want_grad:
        # x = i - sx  =>  d(x1-x)/dsx = +1, d(x-x0)/dsx = -1
        dsx = fa * wy0 + fb * wy1 - fc * wy0 - fd * wy1
        dsy = fa * wx0 - fb * wx0 + fc * wx1 - fd * wx1
        grads = (dsx, dsy)
    if not keep_pad:
        out = out[1:-1, 1:-1]
        if want_grad:
            grads = (grads[0][1:-1, 1:-1], grads[1][1:-1, 1:-1])
    return out, grads


def _lanczos_taps(d, k, want_grad):
    """Lx[i] = sinc(t) sinc(t / k), t = i + d, i = -k..k, divided by its sum (utils/interpolate.py:145-163 called with
    -shift: sinc is even, so the sign flips there cancel), and its derivative wrt d."""
    t = np.arange(-k, k + 1, dtype=np.float64) + d
    L = np.sinc(t) * np.sinc(t / k)
    S = L.sum()
    if not want_grad:
        return L / S, None

    def dsinc(u):
        small = np.abs(u) < 1e-2
        us = np.where(small, 1.0, u)
        big = (np.cos(np.pi * us) - np.sinc(us)) / us
        ser = -(np.pi**2 / 3) * u + (np.pi**4 / 30) * u**3 - (np.pi**6 / 840) * u**5
        return np.where(small, ser, big)

    dL = dsinc(t) * np.sinc(t / k) + np.sinc(t) * dsinc(t / k) / k
    return L / S, dL / S - L * (dL.sum() / S**2)


def shift_psf_lanczos(psf, shift, k, keep_pad=True, want_grad=False):
    """_model_methods.py:209-227: the PSF zero-padded by k, cross-correlated ('same') with the normalised separable
    Lanczos-k kernel of the shift.  Returns stamp (and d/dsx, d/dsy)."""
    im = np.pad(np.asarray(psf, dtype=np.float64), k)
    Lx, dLx = _lanczos_taps(shift[0], k, want_grad)
    Ly, dLy = _lanczos_taps(shift[1], k, want_grad)

    def corr(kx, ky):
        wide = np.pad(im, k)
        tmp = sum(kx[i] * wide[:, i : i + im.shape[1]] for i in range(2 * k + 1))
        return sum(ky[j] * tmp[j : j + im.shape[0], :] for j in range(2 * k + 1))

    out = corr(Lx, Ly)
    grads = (corr(dLx, Ly), corr(Lx, dLy)) if want_grad else None
    if not keep_pad:
        out = out[k:-k, k:-k]
        if want_grad:
            grads = (grads[0][k:-k, k:-k], grads[1][k:-k, k:-k])
    return out, grads


def _shift_psf(psf, shift, method, keep_pad, want_grad):
    if method == sc.SHIFT_BILINEAR:
        return shift_psf_bilinear(psf, shift, keep_pad, want_grad)
    if method > sc.SHIFT_LANCZOS:
        return shift_psf_lanczos(psf, shift, method - sc.SHIFT_LANCZOS, keep_pad, want_grad)
    raise NotImplementedError(f"sub-pixel shift method {method}")


def conv_same_circular(img, ker):
    """The reference's FFT convolution on the pre-padded image (utils/operations.py:9-36 with img_prepadded=True): a
    CIRCULAR convolution of the image's own size.  For the bilinear shift the kernel reaches exactly to the PSF border,
    nothing wraps and this equals the linear convolution; the Lanczos-k stamp is 2 (k - 1) pixels wider, and its outer
    taps bring pixels of the opposite edge into the k - 1 outermost rows / columns of the cropped result."""
    H, W = img.shape
    kh, kw = ker.shape
    f = np.fft.irfft2(np.fft.rfft2(img) * np.fft.rfft2(ker, s=(H, W)), s=(H, W))
    return np.roll(f, (-((kh - 1) // 2), -((kw - 1) // 2)), axis=(0, 1))


def normalized_shifted_psf(psf, shift, method, keep_pad, want_grad):
    """Shifted PSF divided by its sum (_model_methods.py:239-243), with the
    quotient-rule derivative wrt the shift."""
    if method == sc.SHIFT_NONE or shift is None:
        p = np.asarray(psf, dtype=np.float64)
        return p / p.sum(), None
    st, g = _shift_psf(psf, shift, method, keep_pad, want_grad)
    tot = st.sum()
    out = st / tot
    if not want_grad:
        return out, None
    dout = tuple(gi / tot - st * (gi.sum() / tot**2) for gi in g)
    return out, dout


def shifted_psf_param_derivative(psf, dpsf, shift, method, keep_pad):
    """d/d theta of the shifted, normalised stamp when the PSF itself depends on theta (auxiliary PSF
    model): the shift is linear in the PSF, so  d P = S(dpsf)/T - S(psf) sum(S(dpsf))/T^2,  T = sum S(psf)."""
    if method == sc.SHIFT_NONE or shift is None:
        st, dst = np.asarray(psf, dtype=np.float64), np.asarray(dpsf, dtype=np.float64)
    else:
        st, _ = _shift_psf(psf, shift, method, keep_pad, False)
        dst, _ = _shift_psf(dpsf, shift, method, keep_pad, False)
    tot = st.sum()
    return dst / tot - st * (dst.sum() / tot**2)


def aux_psf(scene, ps, vals, mode, want_grad):
    """Stamp of an auxiliary PSF model (model_object.py:307-310: ``psf = self.psf(parameters=...)``), sampled
    on its own grid like any PSF model (psf_model_object.py:180-264), with its derivatives (natural units)."""
    psrc = scene.sources[ps.source]
    el = source_elements(psrc, vals)
    r = sample_source(scene, ps.source, el, mode, want_grad)
    return psrc, r


def conv_same(img, ker):
    """Linear 'same' convolution, direct sum (what the reference's FFT /
    conv2d paths compute up to rounding, _model_methods.py:245-255)."""
    kh, kw = ker.shape
    ph, pw = (kh - 1) // 2, (kw - 1) // 2
    H, W = img.shape
    pad = np.zeros((H + kh - 1, W + kw - 1))
    pad[ph : ph + H, pw : pw + W] = img
    out = np.zeros_like(img, dtype=np.float64)
    for a in range(kh):
        for b in range(kw):
            k = ker[a, b]
            if k != 0.0:
                out += k * pad[kh - 1 - a : kh - 1 - a + H, kw - 1 - b : kw - 1 - b + W]
    return out


def conv_same_fft(img, ker):
    """Same result via scipy FFTs (used for the timed CPU baseline, like the
    reference's default psf_convolve_mode='fft', utils/operations.py:9-36)."""
    from scipy.signal import fftconvolve
    return fftconvolve(img, ker, mode="same")


# ---------------------------------------------------------------------------
# one source
# ---------------------------------------------------------------------------
def _affine(img, i, j):
    """Plane coordinates of pixel indices (wcs.py:561-584)."""
    di = i - img.rij[0]
    dj = j - img.rij[1]
    return (img.S[0, 0] * di + img.S[0, 1] * dj + img.rxy[0],
            img.S[1, 0] * di + img.S[1, 1] * dj + img.rxy[1])


def _psf_border(psf):
    h, w = psf.shape
    return int(math.ceil((1 + w) / 2)), int(math.ceil((1 + h) / 2))


class SourceResult:
    """Output-window stamps of one source: value (oh, ow) and derivatives
    (n_elem, oh, ow) in natural units (None when not requested)."""

    def __init__(self, value, grad, extra=None):
        self.value = value
        self.grad = grad
        self.extra = extra or []   # [(slot, plane)]: derivatives wrt parameters of an auxiliary PSF model (natural units)


def sample_source(scene, si, el, mode, want_grad, conv="direct", stats=None, vals=None):
    """One component model on its output window, times the model's own mask (model_object.py:370-371,
    point_source.py:184-185, psf_model_object.py:259-260: ``working_image.data * logical_not(self.mask)`` -- the
    reference differentiates through that product, so the derivative planes are masked too)."""
    r = _sample_source_unmasked(scene, si, el, mode, want_grad, conv, stats, vals)
    src = scene.sources[si]
    m = getattr(src, "mask", None)
    if m is None:
        return r
    mx0, my0 = src.mask_origin
    ox, oy, ow, oh = src.out
    keep = ~np.asarray(m, dtype=bool)[oy - my0 : oy - my0 + oh, ox - mx0 : ox - mx0 + ow]
    r.value = r.value * keep
    if r.grad is not None:
        r.grad = r.grad * keep[None]
    r.extra = [(sl, pl * keep) for sl, pl in r.extra]
    return r


def _sample_source_unmasked(scene, si, el, mode, want_grad, conv="direct", stats=None, vals=None):
    """One component model on its output window.

    mode: "fwd" (working window = group window; group_model_object.py:211-227
    passes ``window=use_window``) or "jac" (working window = own window,
    _model_methods.py:294-299).  Pipeline per model_object.py:258-375.
    """
    src = scene.sources[si]
    img = scene.images[src.image]
    S = np.asarray(img.S, dtype=np.float64)
    Sinv = np.linalg.inv(S)
    area = abs(np.linalg.det(S))
    ox, oy, ow, oh = src.out
    wx, wy, ww, wh = src.fwd if mode == "fwd" else src.jac
    ne = src.n_elem

    if src.kind == sc.KIND_FLAT_SKY:
        X = np.zeros((oh, ow))
        I, dI = eval_profile(src, el, X, X, area, want_grad)
        return SourceResult(I, dI)

    up = int(getattr(src, "upscale", 1) or 1)
    if up > 1:
        # Super-sampled PSF (model_object.py:312-315,348-349; point_source.py:147-149,181): the working window is
        # rescaled to the PSF's pixels (window_object.py:233-239: pixelscale / up, reference_imageij -> (rij + 0.5) up - 0.5),
        # the model is sampled / integrated / convolved (a point source: its PSF dropped) on that grid, and the result is
        # block-summed back (image_object.py:331-376 ``reduce``).  Here: the same source on a virtual fine image.
        import dataclasses
        fine_img = dataclasses.replace(img, H=img.H * up, W=img.W * up, S=S / up,
                                       rij=(np.asarray(img.rij, dtype=np.float64) + 0.5) * up - 0.5,
                                       data=None, weight=None, mask=None)
        fine_src = dataclasses.replace(src, upscale=1, image=len(scene.images), mask=None,
                                       out=tuple(v * up for v in src.out), fwd=tuple(v * up for v in src.fwd),
                                       jac=tuple(v * up for v in src.jac))
        srcs = list(scene.sources)
        srcs[si] = fine_src
        scene1 = dataclasses.replace(scene, images=list(scene.images) + [fine_img], sources=srcs)
        r = _sample_source_unmasked(scene1, si, el, mode, want_grad, conv, stats, vals)
        red = lambda a: a.reshape(a.shape[:-2] + (oh, up, ow, up)).sum(axis=(-3, -1))
        return SourceResult(red(r.value), None if r.grad is None else red(r.grad), [(sl, red(pl)) for sl, pl in r.extra])

    if src.kind == sc.KIND_POINT:
        return _sample_point(scene, src, img, el, want_grad)

    if src.flags & sc.FLAG_AMP:
        # Point source drawn from a PSF *model* (point_source.py:122-140): the PSF model sampled on the working window
        # shifted by -centre (its own sampling / integration knobs, normalised over that window when normalize_psf),
        # then scaled by 10^flux.  The amplitude is the last element; the profile never sees it.
        import dataclasses
        base = dataclasses.replace(src, slot=list(src.slot[:-1]), cval=list(src.cval[:-1]),
                                   flags=src.flags & ~sc.FLAG_AMP)
        scene1 = dataclasses.replace(scene, sources=[base])
        r = sample_source(scene1, 0, el[:-1], mode, want_grad, conv, stats, vals)
        A = 10.0 ** el[-1]
        val = A * r.value
        grad = None
        if want_grad:
            grad = np.concatenate([A * r.grad, (LN10 * val)[None]], axis=0)
        return SourceResult(val, grad)

    has_psf = src.psf >= 0
    bx = by = 0
    psf_src = psf_res = None
    if has_psf:
        ps = scene.psfs[src.psf]
        if getattr(ps, "source", -1) >= 0:
            psf_src, psf_res = aux_psf(scene, ps, vals, mode, want_grad)
            psf = psf_res.value
        else:
            psf = _host(ps.data)
        bx, by = _psf_border(psf)
    # evaluation region: output window + psf border; working region likewise
    ex0, ey0, ew, eh = ox - bx, oy - by, ow + 2 * bx, oh + 2 * by
    rx0, ry0, rw, rh = wx - bx, wy - by, ww + 2 * bx, wh + 2 * by
    cx, cy = el[0], el[1]
    shift = None
    if has_psf and src.psf_shift != sc.SHIFT_NONE:
        pc = Sinv @ (np.array([cx, cy]) - img.rxy) + img.rij       # pixel coords of the centre
        rnd = np.round(pc)
        shift = pc - rnd
        # grid re-centred on the source (model_object.py:320-323): X = S.(pix - round(pc))
        def coords(i, j):
            di, dj = i - rnd[0], j - rnd[1]
            return S[0, 0] * di + S[0, 1] * dj, S[1, 0] * di + S[1, 1] * dj
        center_in_grid = False
    else:
        def coords(i, j):
            px, py = _affine(img, i, j)
            return px - cx, py - cy
        center_in_grid = True

    # ---- first pass over the evaluation region (+1 px ring for the curvature stencil)
    mx0, my0 = max(ex0 - 1, rx0), max(ey0 - 1, ry0)
    mx1, my1 = min(ex0 + ew + 1, rx0 + rw), min(ey0 + eh + 1, ry0 + rh)
    jj, ii = np.meshgrid(np.arange(my0, my1, dtype=np.float64), np.arange(mx0, mx1, dtype=np.float64), indexing="ij")
    X, Y = coords(ii, jj)
    nspe = 0
    if src.sampling_mode in (sc.SAMPLE_MIDPOINT, sc.SAMPLE_TRAPEZOID):
        if src.sampling_mode == sc.SAMPLE_MIDPOINT:
            deep, ddeep = eval_profile(src, el, X, Y, area, want_grad)
            nspe += X.size
        else:
            # trapezoid (_model_methods.py:124-143): mean of the four pixel-corner values, then the same
            # curvature proxy as midpoint on that image
            deep, ddeep = _trapezoid(src, el, X, Y, S, area, want_grad)
            nspe += 4 * X.size
        # curvature: valid 3x3 Laplacian, replicate-padded *over the working region*
        # (_model_methods.py:87-98): evaluate the stencil at the index clamped to the interior
        def lap_at(i, j):   # absolute pixel indices (arrays)
            ic = np.clip(i, rx0 + 1, rx0 + rw - 2) - mx0
            jc = np.clip(j, ry0 + 1, ry0 + rh - 2) - my0
            return (deep[jc - 1, ic] + deep[jc + 1, ic] + deep[jc, ic - 1] + deep[jc, ic + 1] - 4 * deep[jc, ic])
        EJ, EI = np.meshgrid(np.arange(ey0, ey0 + eh), np.arange(ex0, ex0 + ew), indexing="ij")
        if rw >= 3 and rh >= 3:
            err = np.abs(lap_at(EI, EJ))
        else:
            err = np.zeros((eh, ew))
        sl = (slice(ey0 - my0, ey0 - my0 + eh), slice(ex0 - mx0, ex0 - mx0 + ew))
        deep_e = deep[sl].copy()
        ddeep_e = ddeep[(slice(None),) + sl].copy() if want_grad else None
        Xe, Ye = X[sl], Y[sl]
        mean_src = deep
    elif src.sampling_mode == sc.SAMPLE_QUAD:
        sl = (slice(ey0 - my0, ey0 - my0 + eh), slice(ex0 - mx0, ex0 - mx0 + ew))
        Xe, Ye = X[sl], Y[sl]
        deep_e, refc, ddeep_e = _gl(src, el, Xe, Ye, S, area, src.quad_init, want_grad)
        nspe += Xe.size * src.quad_init**2
        err = np.abs(deep_e - refc)
        mean_src = deep_e
    elif src.sampling_mode == sc.SAMPLE_SIMPSONS:
        sl = (slice(ey0 - my0, ey0 - my0 + eh), slice(ex0 - mx0, ex0 - mx0 + ew))
        Xe, Ye = X[sl], Y[sl]
        # (2h+1) x (2w+1) half-pixel lattice, 3x3 Simpson weights stride 2 (_model_methods.py:99-109)
        wts = np.array([[1, 4, 1], [4, 16, 4], [1, 4, 1]], dtype=np.float64) / 36.0
        deep_e = np.zeros((eh, ew))
        ddeep_e = np.zeros((ne, eh, ew)) if want_grad else None
        mid = None
        for a in (-1, 0, 1):
            for b in (-1, 0, 1):
                dxp = S[0, 0] * (0.5 * b) + S[0, 1] * (0.5 * a)
                dyp = S[1, 0] * (0.5 * b) + S[1, 1] * (0.5 * a)
                I, dI = eval_profile(src, el, Xe + dxp, Ye + dyp, area, want_grad)
                deep_e += wts[a + 1, b + 1] * I
                if want_grad:
                    ddeep_e += wts[a + 1, b + 1] * dI
                if a == 0 and b == 0:
                    mid = I
        nspe += 9 * Xe.size
        err = np.abs(deep_e - mid)
        mean_src = deep_e
    else:
        raise NotImplementedError("sampling mode not in oracle")

    # ---- threshold integration (_model_methods.py:155-184)
    if src.integrate_mode == sc.INTEGRATE_THRESHOLD:
        if src.ref_mode == sc.REF_SERSIC_FLUX:
            ref = sersic_total_flux(10.0 ** el[6], el[4], el[5], el[2]) / (rw * rh)
        else:
            ringed = src.sampling_mode in (sc.SAMPLE_MIDPOINT, sc.SAMPLE_TRAPEZOID)
            if (mx0, my0, mx1, my1) == (rx0, ry0, rx0 + rw, ry0 + rh) or not ringed and (ex0, ey0, ew, eh) == (rx0, ry0, rw, rh):
                ref = mean_src.sum() / (rw * rh)
            else:
                ref = _mean_over_region(src, el, coords, S, area, rx0, ry0, rw, rh)
        thr = src.tolerance * ref
        select = err > thr
        if np.any(select):
            grid_integrate.spe = 0
            grid_integrate.queued = {}
            ires, idres = grid_integrate(src, el, Xe[select], Ye[select], S, area, 1, thr, want_grad)
            nspe += grid_integrate.spe
            deep_e[select] = ires
            if want_grad:
                ddeep_e[:, select] = idres
            if stats is not None:
                for d, c in grid_integrate.queued.items():
                    stats.setdefault("queued", {}).setdefault(d, 0)
                    stats["queued"][d] += c
    if stats is not None:
        stats["spe"] = stats.get("spe", 0) + nspe

    if src.flags & sc.FLAG_NORMALIZE:
        tot = deep_e.sum()
        if want_grad:
            ddeep_e = ddeep_e / tot - deep_e[None] * (ddeep_e.sum(axis=(1, 2), keepdims=True) / tot**2)
        deep_e = deep_e / tot

    if not has_psf:
        return SourceResult(deep_e, ddeep_e)

    # ---- PSF convolution on the padded region, then crop the border (model_object.py:342-349)
    cfn = conv_same if conv == "direct" else conv_same_fft
    if src.psf_shift > sc.SHIFT_LANCZOS:
        cfn = conv_same_circular     # the wider Lanczos stamp wraps around the padded image in the reference
    P, dP = normalized_shifted_psf(psf, shift, src.psf_shift, True, want_grad and shift is not None)
    crop = (slice(by, by + oh), slice(bx, bx + ow))
    val = cfn(deep_e, P)[crop]
    grad = None
    if want_grad:
        grad = np.zeros((ne, oh, ow))
        first = 0 if center_in_grid else 2
        todo = [e for e in range(first, ne) if src.slot[e] >= 0]
        for e, plane in zip(todo, _pmap(lambda e: cfn(ddeep_e[e], P)[crop], todo)):
            grad[e] = plane
        if not center_in_grid:
            # centre enters only through the PSF shift: d shift / d centre = S^-1
            gsx = cfn(deep_e, dP[0])[crop]
            gsy = cfn(deep_e, dP[1])[crop]
            grad[0] = Sinv[0, 0] * gsx + Sinv[1, 0] * gsy
            grad[1] = Sinv[0, 1] * gsx + Sinv[1, 1] * gsy
    extra = []
    if want_grad and psf_src is not None:
        for e, sl in enumerate(psf_src.slot):
            if sl >= 0:
                dPe = shifted_psf_param_derivative(psf, psf_res.grad[e], shift, src.psf_shift, True)
                extra.append((sl, cfn(deep_e, dPe)[crop]))
    return SourceResult(val, grad, extra)


def _trapezoid(src, el, X, Y, S, area, want_grad):
    """Mean of the profile at the four corners of every pixel (2x2 box filter on the corner lattice)."""
    tot, dtot = 0.0, 0.0
    for a in (-0.5, 0.5):
        for b in (-0.5, 0.5):
            I, dI = eval_profile(src, el, X + (S[0, 0] * b + S[0, 1] * a), Y + (S[1, 0] * b + S[1, 1] * a), area, want_grad)
            tot = tot + 0.25 * I
            if want_grad:
                dtot = dtot + 0.25 * dI
    return tot, (dtot if want_grad else None)


def _mean_over_region(src, el, coords, S, area, rx0, ry0, rw, rh):
    """Mean of the first-pass image over the whole working region (the default
    ``_integrate_reference``, _model_methods.py:151-152), in row blocks."""
    tot = 0.0
    for j0 in range(ry0, ry0 + rh, 256):
        j1 = min(j0 + 256, ry0 + rh)
        jj, ii = np.meshgrid(np.arange(j0, j1, dtype=np.float64), np.arange(rx0, rx0 + rw, dtype=np.float64), indexing="ij")
        X, Y = coords(ii, jj)
        if src.sampling_mode == sc.SAMPLE_MIDPOINT:
            I, _ = eval_profile(src, el, X, Y, area, False)
        elif src.sampling_mode == sc.SAMPLE_TRAPEZOID:
            I, _ = _trapezoid(src, el, X, Y, S, area, False)
        elif src.sampling_mode == sc.SAMPLE_QUAD:
            I, _, _ = _gl(src, el, X, Y, S, area, src.quad_init, False)
        elif src.sampling_mode == sc.SAMPLE_SIMPSONS:
            # 3x3 Simpson weights on the half-pixel lattice (_model_methods.py:99-109)
            I = np.zeros(X.shape)
            for a in (-1, 0, 1):
                for b in (-1, 0, 1):
                    wt = (4.0 if a == 0 else 1.0) * (4.0 if b == 0 else 1.0) / 36.0
                    dxp = S[0, 0] * (0.5 * b) + S[0, 1] * (0.5 * a)
                    dyp = S[1, 0] * (0.5 * b) + S[1, 1] * (0.5 * a)
                    I += wt * eval_profile(src, el, X + dxp, Y + dyp, area, False)[0]
        else:
            raise NotImplementedError
        tot += I.sum()
    return tot / (rw * rh)


def _sample_point(scene, src, img, el, want_grad):
    """Point source with a PSF image (models/point_source.py:145-175): the PSF,
    sub-pixel shifted and normalised, times 10^flux, dropped at the rounded
    pixel of the centre and clipped to the output window."""
    S = np.asarray(img.S, dtype=np.float64)
    Sinv = np.linalg.inv(S)
    psf = _host(scene.psfs[src.psf].data)
    ph, pw = psf.shape
    ox, oy, ow, oh = src.out
    pc = Sinv @ (np.array([el[0], el[1]]) - img.rxy) + img.rij
    rnd = np.round(pc)
    shift = pc - rnd
    P, dP = normalized_shifted_psf(psf, shift, src.psf_shift, False, want_grad)
    F = 10.0 ** el[2]
    val = np.zeros((oh, ow))
    ne = src.n_elem
    grad = np.zeros((ne, oh, ow)) if want_grad else None
    # psf pixel (a, b) lands on image pixel (rnd_x - (pw-1)/2 + b, rnd_y - (ph-1)/2 + a)
    x_lo = int(rnd[0]) - (pw - 1) // 2
    y_lo = int(rnd[1]) - (ph - 1) // 2
    ix0, ix1 = max(ox, x_lo), min(ox + ow, x_lo + pw)
    iy0, iy1 = max(oy, y_lo), min(oy + oh, y_lo + ph)
    if ix1 > ix0 and iy1 > iy0:
        ps = (slice(iy0 - y_lo, iy1 - y_lo), slice(ix0 - x_lo, ix1 - x_lo))
        os_ = (slice(iy0 - oy, iy1 - oy), slice(ix0 - ox, ix1 - ox))
        val[os_] = F * P[ps]
        if want_grad:
            if dP is not None:
                gsx, gsy = F * dP[0][ps], F * dP[1][ps]
                grad[0][os_] = Sinv[0, 0] * gsx + Sinv[1, 0] * gsy
                grad[1][os_] = Sinv[0, 1] * gsx + Sinv[1, 1] * gsy
            grad[2][os_] = LN10 * F * P[ps]
    return SourceResult(val, grad)


# ---------------------------------------------------------------------------
# whole scene
# ---------------------------------------------------------------------------
def _values(scene, x, as_rep):
    if as_rep:
        return rep_to_val(x, scene.transform, scene.lo, scene.hi)
    x = np.asarray(x, dtype=np.float64)
    return x.copy(), np.ones_like(x)


def sample(scene, x, as_rep=True, mode="fwd", conv="direct", stats=None):
    """Model image per scene image: sum of all sources (group_model_object.py:183-231)."""
    vals, _ = _values(scene, x, as_rep)
    out = [np.zeros((im.H, im.W)) for im in scene.images if not getattr(im, "aux", False)]
    todo = [si for si, src in enumerate(scene.sources) if not getattr(scene.images[src.image], "aux", False)]
    # (auxiliary PSF models are sampled by the sources that use them)
    st = stats if N_THREADS <= 1 else None        # the counters are not thread-safe
    res = _pmap(lambda si: sample_source(scene, si, source_elements(scene.sources[si], vals), mode, False, conv, st,
                                         vals=vals), todo)
    for si, r in zip(todo, res):
        ox, oy, ow, oh = scene.sources[si].out
        out[scene.sources[si].image][oy : oy + oh, ox : ox + ow] += r.value
    return out


def jacobian(scene, x, as_rep=True, conv="direct", stats=None):
    """Dense (H, W, P) Jacobian per image (group_model_object.py:233-283).
    Only for small scenes."""
    vals, dv = _values(scene, x, as_rep)
    P = scene.n_par
    out = [np.zeros((im.H, im.W, P)) for im in scene.images if not getattr(im, "aux", False)]
    todo = []
    for si, src in enumerate(scene.sources):
        if getattr(scene.images[src.image], "aux", False):
            continue
        aux = src.psf >= 0 and getattr(scene.psfs[src.psf], "source", -1) >= 0
        if all(s < 0 for s in src.slot) and not aux:
            continue
        todo.append(si)
    st = stats if N_THREADS <= 1 else None
    res = _pmap(lambda si: sample_source(scene, si, source_elements(scene.sources[si], vals), "jac", True, conv, st,
                                         vals=vals), todo)
    for si, r in zip(todo, res):
        src = scene.sources[si]
        ox, oy, ow, oh = src.out
        for e, s in enumerate(src.slot):
            if s >= 0:
                out[src.image][oy : oy + oh, ox : ox + ow, s] += r.grad[e] * dv[s]
        for s, plane in r.extra:
            out[src.image][oy : oy + oh, ox : ox + ow, s] += plane * dv[s]
    return out


def flat_targets(scene):
    """Y, W, keep-mask as flat vectors over all images (lm.py:191-222)."""
    ims = [im for im in scene.images if not getattr(im, "aux", False)]
    Y = np.concatenate([_host(im.data).reshape(-1) for im in ims])
    W = np.concatenate([(np.ones(im.H * im.W) if im.weight is None else _host(im.weight).reshape(-1))
                        for im in ims])
    keep = np.concatenate([(np.ones(im.H * im.W, dtype=bool) if im.mask is None else ~_host(im.mask, bool).reshape(-1))
                           for im in ims])
    return Y, W, keep


def normal_eq(scene, x, conv="direct", same_geometry=False, stats=None):
    """J^T W J, J^T W (Y - Y0), sum W (Y - Y0)^2 over unmasked pixels
    (lm.py:256-260,373-399)."""
    Y, W, keep = flat_targets(scene)
    Y0 = np.concatenate([m.reshape(-1) for m in sample(scene, x, True, "fwd", conv, stats)])
    J = np.concatenate([j.reshape(-1, scene.n_par) for j in jacobian(scene, x, True, conv, stats)])
    Jk, Wk = J[keep], W[keep]
    H = Jk.T @ (Wk[:, None] * Jk)
    g = Jk.T @ (Wk * (Y[keep] - Y0[keep]))
    chi2 = np.sum(Wk * (Y[keep] - Y0[keep]) ** 2)
    return H, g, chi2, (J, Y0)


def chi2(scene, x, conv="direct"):
    Y, W, keep = flat_targets(scene)
    Y1 = np.concatenate([m.reshape(-1) for m in sample(scene, x, True, "fwd", conv)])
    return np.sum((W * (Y - Y1) ** 2)[keep])


def lm_solve(H, g, L):
    """Damped step (lm.py:359-371)."""
    P = len(g)
    I = np.eye(P)
    D = np.ones_like(H) - I
    return np.linalg.solve(H * (I + D / (1 + L)) + L * I * (1 + np.diag(H)), g)


class OptimizeStop(Exception):
    pass


def lm_fit(scene, x0, max_iter=100, relative_tolerance=1e-5, ndf=None, max_step_iter=10,
           curvature_limit=1.0, Lup=11.0, Ldn=9.0, L0=1.0, acceleration=0.0, conv="direct", verbose=0, stop=None):
    """Levenberg-Marquardt loop, control flow of fit/lm.py:248-357,428-493.
    ``stop``: optional callable(loss_history, L) -> bool checked after every iteration (bench.py ends a fit with the same
    rule as its GPU arm); the result carries the wall-clock seconds of every iteration in ``iter_seconds``."""
    import time as _time
    x = np.asarray(x0, dtype=np.float64).copy()
    Y, W, keep = flat_targets(scene)
    if ndf is None:
        ndf = max(1.0, float(keep.sum()) - len(x))
    L = L0

    def up(L):
        return min(1e9, L * Lup)

    def dn(L):
        return max(1e-9, L / Ldn)

    def fwd(xx):
        return np.concatenate([m.reshape(-1) for m in sample(scene, xx, True, "fwd", conv)])

    def c2(Yp):
        return float(np.sum((W * (Y - Yp) ** 2)[keep]) / ndf)

    loss = [c2(fwd(x))]
    L_hist = [L]
    x_hist = [x.copy()]
    trials_hist = []
    message = ""
    iter_seconds = []
    for it in range(max_iter):
        _t0 = _time.perf_counter()
        # ---- step (lm.py:248-357)
        Y0 = fwd(x)
        J = np.concatenate([j.reshape(-1, scene.n_par) for j in jacobian(scene, x, True, conv)])
        Jk, Wk = J[keep], W[keep]
        r = Wk * (Y0[keep] - Y[keep])
        H = Jk.T @ (Wk[:, None] * Jk)
        g = -Jk.T @ r
        init = loss[-1]
        nostep = True
        best = (np.zeros_like(x), init, L)
        scary = (None, init, L)
        direction = "none"
        d = 0.1
        ntr = 0
        for k in range(max_step_iter):
            ntr += 1
            if k > max_step_iter / 2 and L < 1e-3:
                L = 1.0
            h = lm_solve(H, g, L)
            Y1 = fwd(x + d * h)
            rh = Wk * (Y1[keep] - Y[keep])
            rpp = Jk.T @ ((2 / d) * ((rh - r) / d - Wk * (Jk @ h)))
            a = -lm_solve(H, rpp, L) / 2 if L > 1e-4 else np.zeros_like(h)
            ha = h + a * acceleration
            chi = c2(fwd(x + ha))
            if verbose > 1:
                print(f"  sub step L: {L}, Chi^2/DoF: {chi}")
            if not np.isfinite(chi):
                L = up(L)
                if direction == "better":
                    break
                direction = "worse"
                continue
            if chi <= scary[1]:
                scary = (ha, chi, L)
            rho = np.linalg.norm(a) / np.linalg.norm(h)
            if rho > curvature_limit:
                L = up(L)
                if direction == "better":
                    break
                direction = "worse"
                continue
            if chi < best[1]:
                best = (ha, chi, L)
                nostep = False
                L = dn(L)
                if L <= 1e-8 or direction == "worse":
                    break
                direction = "better"
            elif chi > best[1] and direction in ("none", "worse"):
                L = up(L)
                if L == 1e9:
                    break
                direction = "worse"
            else:
                break
            if (best[1] - init) / init < -0.1:
                break
        trials_hist.append(ntr)
        iter_seconds.append(_time.perf_counter() - _t0)
        if nostep:
            if scary[0] is not None:
                res = scary
            else:
                message += "fail. Could not find step to improve Chi^2"
                break
        else:
            res = best
        L = res[2]
        x = x + res[0]
        L_hist.append(L)
        loss.append(res[1])
        x_hist.append(x.copy())
        L = dn(L)
        if verbose:
            print(f"Chi^2/DoF: {loss[-1]}, L: {L}")
        if stop is not None and stop(loss, L):
            message += "stopped"
            break
        if len(loss) >= 3 and (loss[-3] - loss[-1]) / loss[-1] < relative_tolerance and L < 0.1:
            message += "success"
            break
        if len(loss) > 10 and (loss[-10] - loss[-1]) / loss[-1] < relative_tolerance:
            message += "success by immobility. Convergence not guaranteed"
            break
    else:
        message += "fail. Maximum iterations"
    return {"x": x, "loss_history": loss, "L_history": L_hist, "lambda_history": x_hist,
            "message": message, "trials": trials_hist, "iter_seconds": iter_seconds}
