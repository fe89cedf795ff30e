"""ap.utils.conversions / parametric_profiles / angle_operations / the PSF builders: the helper surface scripts written
for the reference use (its own tests/test_utils.py checks of these pass against this package too).  Known answers here
come from direct numerical integration and from round trips."""
import numpy as np
import torch
from scipy.integrate import quad

import astrophot_b200 as ap

F = ap.utils.conversions.functions
U = ap.utils.conversions.units
C = ap.utils.conversions.coordinates
P = ap.utils.parametric_profiles


def test_sersic_and_moffat_total_flux_are_the_integrals_of_the_profiles():
    for n, Re, Ie, q in [(1.0, 3.0, 2.0, 1.0), (2.5, 5.0, 0.7, 0.6), (4.0, 2.0, 1.3, 0.8)]:
        num = 2 * np.pi * q * quad(lambda r: r * P.sersic_np(np.array([r]), n, Re, Ie)[0], 0, np.inf, limit=400)[0]
        # b(n) is an asymptotic series: the closed form is exact for that b
        assert abs(F.sersic_Ie_to_flux_np(Ie, n, Re, q) - num) / num < 1e-7
        assert abs(F.sersic_flux_to_Ie_np(num, n, Re, q) - Ie) / Ie < 1e-7
        half = 2 * np.pi * q * quad(lambda r: r * P.sersic_np(np.array([r]), n, Re, Ie)[0], 0, Re, limit=400)[0]
        assert abs(half / num - 0.5) < 2e-4            # Re encloses half of the light
        I0 = Ie * np.exp(F.sersic_n_to_b(n))
        assert abs(F.sersic_I0_to_flux_np(I0, n, Re / F.sersic_n_to_b(n) ** n, q) - num) / num < 1e-7
        assert abs(F.sersic_flux_to_I0_np(num, n, Re / F.sersic_n_to_b(n) ** n, q) - I0) / I0 < 1e-7
        assert abs(F.sersic_inv_np(P.sersic_np(np.array([1.7 * Re]), n, Re, Ie)[0], n, Re, Ie) - 1.7 * Re) < 1e-9
        t = lambda v: torch.tensor(v, dtype=torch.float64)
        assert abs(float(F.sersic_Ie_to_flux_torch(t(Ie), t(n), t(Re), t(q))) - F.sersic_Ie_to_flux_np(Ie, n, Re, q)) < 1e-9 * num
        assert abs(float(F.sersic_inv_torch(t(0.3 * Ie), t(n), t(Re), t(Ie))) - F.sersic_inv_np(0.3 * Ie, n, Re, Ie)) < 1e-10
    for n, Rd, I0, q in [(2.5, 3.0, 1.0, 1.0), (1.8, 2.0, 0.4, 0.5)]:
        num = 2 * np.pi * q * quad(lambda r: r * P.moffat_np(r, n, Rd, I0), 0, np.inf, limit=400)[0]
        assert abs(F.moffat_I0_to_flux(I0, n, Rd, q) - num) / num < 1e-7


def test_profiles_numpy_torch_and_conventions():
    R = np.linspace(0.1, 20, 50)
    Rt = torch.as_tensor(R)
    np.testing.assert_allclose(P.sersic_torch(Rt, 2.0, 5.0, 1.5).numpy(), P.sersic_np(R, 2.0, 5.0, 1.5), rtol=1e-14)
    assert np.all(P.sersic_np(R, -1.0, 5.0, 1.0) == 1e6)                       # the optimiser wall
    np.testing.assert_allclose(P.sersic_np(R, 1.0, 4.0, 2.0), P.exponential_torch(Rt, 4.0, 2.0).numpy(), rtol=1e-12)
    np.testing.assert_allclose(P.exponential_np(R, 2.0, 4.0), P.exponential_torch(Rt, 4.0, 2.0).numpy(), rtol=1e-14)
    assert abs(2 * np.pi * quad(lambda r: r * P.gaussian_np(r, 1.7, 3.0), 0, np.inf)[0] - 3.0 * np.sqrt(2 * np.pi) * 1.7) < 1e-8
    assert abs(P.nuker_np(4.0, 4.0, 2.5, 1.5, 2.0, 0.5) - 2.5) < 1e-14          # I(Rb) = Ib
    np.testing.assert_allclose(P.moffat_torch(Rt, 2.5, 3.0, 1.0).numpy(), P.moffat_np(R, 2.5, 3.0, 1.0), rtol=1e-14)


def test_units_round_trips():
    flux, zp, area = np.array([0.5, 20.0, 3e4]), 22.5, 0.04
    np.testing.assert_allclose(U.mag_to_flux(U.flux_to_mag(flux, zp), zp), flux, rtol=1e-13)
    np.testing.assert_allclose(U.sb_to_flux(U.flux_to_sb(flux, area, zp), area, zp), flux, rtol=1e-13)
    mag, mage = U.flux_to_mag(flux, zp, fluxe=0.01 * flux)
    np.testing.assert_allclose(mage, 2.5 * 0.01 / np.log(10), rtol=1e-13)
    f2, fe2 = U.mag_to_flux(mag, zp, mage=mage)
    np.testing.assert_allclose(fe2, 0.01 * flux, rtol=1e-12)
    assert U.flux_to_mag(1.0, zp) == zp and abs(U.flux_to_sb(1.0, 1.0, zp) - zp) < 1e-15
    assert abs(U.mag_to_magperarcsec2(U.magperarcsec2_to_mag(21.0, a=3.0, b=2.0), a=3.0, b=2.0) - 21.0) < 1e-13
    assert abs(U.mag_to_magperarcsec2(15.0, R=2.0) - (15.0 + 2.5 * np.log10(np.pi * 4))) < 1e-13
    assert abs(U.PA_shift_convention(U.PA_shift_convention(0.3)) - 0.3) < 1e-13 and U.PA_shift_convention(100.0, "deg") == 10.0


def test_coordinates():
    X, Y = np.array([1.0, 0.0, 2.0]), np.array([0.0, 1.0, -1.0])
    x, y = C.Rotate_Cartesian_np(np.pi / 2, X, Y)
    np.testing.assert_allclose([x, y], [-Y, X], atol=1e-15)
    xt, yt = C.Rotate_Cartesian(torch.tensor(0.3, dtype=torch.float64), torch.as_tensor(np.stack([X, Y])))
    xn, yn = C.Rotate_Cartesian_np(0.3, X, Y)
    np.testing.assert_allclose([xt.numpy(), yt.numpy()], [xn, yn], rtol=1e-14)
    # theta = 0: y scaled by q; theta = pi/2: x scaled by q; inverse undoes it
    x, y = C.Axis_Ratio_Cartesian_np(0.5, X, Y, 0.0)
    np.testing.assert_allclose([x, y], [X, 0.5 * Y], atol=1e-15)
    x, y = C.Axis_Ratio_Cartesian_np(0.5, X, Y, np.pi / 2)
    np.testing.assert_allclose([x, y], [0.5 * X, Y], atol=1e-15)
    x, y = C.Axis_Ratio_Cartesian_np(0.5, *C.Axis_Ratio_Cartesian_np(0.5, X, Y, 0.7), 0.7, inv_scale=True)
    np.testing.assert_allclose([x, y], [X, Y], atol=1e-14)
    xt, yt = C.Axis_Ratio_Cartesian(torch.tensor(0.5, dtype=torch.float64), torch.as_tensor(X), torch.as_tensor(Y),
                                    torch.tensor(0.7, dtype=torch.float64))
    xn, yn = C.Axis_Ratio_Cartesian_np(0.5, X, Y, 0.7)
    np.testing.assert_allclose([xt.numpy(), yt.numpy()], [xn, yn], rtol=1e-14)


def test_angle_statistics_and_psf_builders():
    A = ap.utils.angle_operations
    a = np.array([np.pi, 2 * np.pi, 3 * np.pi, 4 * np.pi])
    assert abs(A.Angle_Median(a) + np.pi / 2) < 1e-12 and abs(A.Angle_Scatter(a) - np.pi) < 1e-12
    assert abs(A.Angle_Average(np.array([0.1, 0.3, 2 * np.pi + 0.2])) - 0.2) < 1e-12
    yy, xx = np.mgrid[:41, :41] - 20.0
    pa = 0.6
    u, v = np.cos(pa) * xx + np.sin(pa) * yy, -np.sin(pa) * xx + np.cos(pa) * yy
    assert abs(A.Angle_COM_PA(np.exp(-0.5 * (u**2 / 36 + v**2 / 4)), xx, yy) - pa) < 0.05          # (a moment estimate on a pixel grid)
    for build, args in ((ap.utils.gaussian_psf, (1.5,)), (ap.utils.moffat_psf, (2.5, 2.0))):
        psf = build(*args, 15, 0.8)
        raw = build(*args, 15, 0.8, normalize=False)
        assert psf.shape == (15, 15) and abs(psf.sum() - 1) < 1e-14 and np.allclose(psf, psf.T) and np.allclose(psf, psf[::-1])
        assert abs(raw[7, 7] - 1.0) < 0.1 and np.allclose(raw / raw.sum(), psf)      # sub-sample MEAN: centre ~ profile(0)
    assert ap.utils.initialize.gaussian_psf is ap.utils.gaussian_psf and ap.utils.initialize.moffat_psf is ap.utils.moffat_psf
