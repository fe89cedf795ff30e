"""fit.LM's host-side control flow (lambda search, damping schedule, convergence bookkeeping) without a GPU: the
device plan is replaced by a stand-in that answers normal_eq / geodesic / chi2 from the CPU oracle, the damped solve
by numpy.  The histories must still be the reference's (tests/golden): this guards fit.py itself, which otherwise
only runs where a B200 is present."""
import numpy as np
import pytest
import torch

import astrophot_b200 as ap
import astrophot_oracle as orc
import scenes
from astrophot_b200 import cabi
from conftest import load_golden, golden_data


class OraclePlan:
    def __init__(self, scene, **kw):
        self.scene, self.n_par = scene, scene.n_par
        self.shapes = [(im.H, im.W) for im in scene.images if not im.aux]
        self._cache = None

    def block_doubles(self):
        return 0

    def sample(self, x, as_rep=False):
        x = np.asarray(torch.as_tensor(x).cpu().numpy(), dtype=np.float64)
        return [torch.as_tensor(m) for m in orc.sample(self.scene, x, as_rep)]

    def jacobian(self, x, as_rep=False):
        x = np.asarray(torch.as_tensor(x).cpu().numpy(), dtype=np.float64)
        return [torch.as_tensor(j) for j in orc.jacobian(self.scene, x, as_rep)]

    def reserve(self, caps=None):
        pass

    def normal_eq(self, x, as_rep=True, out=None, check=False):
        x = np.asarray(torch.as_tensor(x).cpu().numpy(), dtype=np.float64)
        if as_rep:
            H, g, chi2, (J, Y0) = orc.normal_eq(self.scene, x)
        else:       # natural units (covariance): J from the natural-unit Jacobian
            Y, W, keep = orc.flat_targets(self.scene)
            Y0 = np.concatenate([m.reshape(-1) for m in orc.sample(self.scene, x, False)])
            J = np.concatenate([j.reshape(-1, self.n_par) for j in orc.jacobian(self.scene, x, False)])
            H = J[keep].T @ (W[keep][:, None] * J[keep])
            g = J[keep].T @ (W[keep] * (Y[keep] - Y0[keep]))
            chi2 = np.sum(W[keep] * (Y[keep] - Y0[keep]) ** 2)
        self._cache = (x.copy(), J, Y0)
        res = (torch.as_tensor(H), torch.as_tensor(g), torch.tensor([chi2, 1.0], dtype=torch.float64))
        if out is not None:
            for o, r in zip(out, res):
                if o is not None:
                    o.copy_(r)
            return out
        return res

    def geodesic(self, xdh, h, d, out=None):
        x, J, Y0 = self._cache
        Y, W, keep = orc.flat_targets(self.scene)
        h = np.asarray(torch.as_tensor(h).numpy(), dtype=np.float64)
        Y1 = np.concatenate([m.reshape(-1) for m in orc.sample(self.scene, np.asarray(torch.as_tensor(xdh).numpy()))])
        r, rh = (W * (Y0 - Y))[keep], (W * (Y1 - Y))[keep]
        rpp = J[keep].T @ ((2 / d) * ((rh - r) / d - W[keep] * (J[keep] @ h)))
        t = torch.as_tensor(rpp)
        if out is not None:
            out.copy_(t)
            return out
        return t

    def chi2(self, x, out=None):
        c = orc.chi2(self.scene, np.asarray(torch.as_tensor(x).numpy(), dtype=np.float64))
        t = torch.tensor([c, 1.0 if np.isfinite(c) else 0.0], dtype=torch.float64)
        if out is not None:
            out.copy_(t)
            return out
        return t


def _solve(H, g, L, out=None, info=None):
    return torch.as_tensor(orc.lm_solve(H.numpy(), g.numpy(), float(L)))


@pytest.fixture
def host_only(monkeypatch):
    monkeypatch.setattr(cabi, "Plan", OraclePlan)
    monkeypatch.setattr(cabi, "lm_solve", _solve)
    old = ap.AP_config.ap_device
    ap.AP_config.ap_device = "cpu"
    yield
    ap.AP_config.ap_device = old


@pytest.mark.parametrize("name", ["c1_sersic", "psf_sersic", "group", "joint", "group_nosky", "masked_locked_edge",
                                  "aux_psf_moffat", "plane_sky_group", "point_psf_model", "point_psf_model_group"])
def test_lm_control_flow_reproduces_the_reference(host_only, name):
    fix = load_golden(name)
    model, _ = scenes.build(ap, name, data=golden_data(fix))
    res = ap.fit.LM(model, initial_state=fix["x0"], max_iter=8, relative_tolerance=0.0, fused_trial=False).fit()
    ref_loss = fix["loss_history"]
    n = min(len(ref_loss), len(res.loss_history))
    moving = 1
    while moving < n and abs(ref_loss[moving] - ref_loss[moving - 1]) / ref_loss[moving] > 1e-12:
        moving += 1
    assert moving >= 3
    np.testing.assert_allclose(res.loss_history[:moving], ref_loss[:moving], rtol=1e-8)
    steady = 1
    while steady < moving and abs(ref_loss[steady] - ref_loss[steady - 1]) / ref_loss[steady] > 1e-6:
        steady += 1
    np.testing.assert_allclose(res.L_history[:steady], fix["L_history"][:steady], rtol=1e-12)
    for k in range(moving):
        np.testing.assert_allclose(res.lambda_history[k], fix["lambda_history"][k], rtol=1e-8, atol=1e-8)
    # uncertainties in natural units (lm.py:408-425,495-539)
    cov = res.covariance_matrix.numpy()
    d = np.sqrt(np.abs(np.diag(fix["cov"])))
    assert np.max(np.abs(cov - fix["cov"]) / np.outer(d, d)) < 1e-6


@pytest.mark.parametrize("name", list(scenes.LM_KWARGS_SCENES))
def test_lm_non_default_knobs(host_only, name):
    """Geodesic acceleration on, another damping schedule (acceleration, curvature_limit, Lup, Ldn, L0,
    max_step_iter): the lambda search of fit.LM against the reference's with the same knobs."""
    fix = load_golden(name)
    model, _ = scenes.build(ap, name, data=golden_data(fix))
    res = ap.fit.LM(model, initial_state=fix["x0"], max_iter=6, relative_tolerance=0.0, fused_trial=False,
                    **scenes.LM_KWARGS).fit()
    ref_loss = fix["kw_loss_history"]
    n = min(len(ref_loss), len(res.loss_history))
    moving = 1
    while moving < n and abs(ref_loss[moving] - ref_loss[moving - 1]) / ref_loss[moving] > 1e-9:
        moving += 1
    assert moving >= 3
    np.testing.assert_allclose(res.loss_history[:moving], ref_loss[:moving], rtol=1e-8)
    np.testing.assert_allclose(res.L_history[:moving - 1], fix["kw_L_history"][:moving - 1], rtol=1e-12)
    for k in range(moving):
        np.testing.assert_allclose(res.lambda_history[k], fix["kw_lambda_history"][k], rtol=1e-7, atol=1e-8)


@pytest.mark.parametrize("name", list(scenes.ITER_SCENES))
def test_iter_control_flow_reproduces_the_reference(host_only, name):
    """fit.Iter (sub-models one at a time on the residual image) end to end on the stand-in plan."""
    fix = load_golden(name)
    model, _ = scenes.build(ap, name, data=golden_data(fix))
    res = ap.fit.Iter(model, initial_state=fix["x0"], max_iter=3,
                      method_kwargs={"max_iter": 4, "relative_tolerance": 0.0, "fused_trial": False}).fit()
    np.testing.assert_allclose(res.loss_history, fix["iter_loss_history"], rtol=1e-8)
    np.testing.assert_allclose(np.array(res.lambda_history), fix["iter_lambda_history"], rtol=1e-7, atol=1e-7)


def check_flux_uncertainties(name):
    """total_flux / total_flux_uncertainty / total_magnitude / total_magnitude_uncertainty (core_model.py:265-290) against
    the reference's own numbers (oracle/make_flux_golden.py), with the same seeded parameter uncertainties."""
    ref = load_golden("flux_uncertainty")[name]
    model, _ = scenes.build(ap, name)
    unc = np.random.default_rng(500 + scenes.FLUX_SCENES.index(name)).uniform(0.01, 0.1, size=len(model.parameters.vector_values()))
    model.parameters.vector_set_uncertainty(torch.as_tensor(unc))
    got = [model.total_flux(), model.total_flux_uncertainty(), model.total_magnitude(), model.total_magnitude_uncertainty()]
    np.testing.assert_allclose([float(g) for g in got], ref, rtol=1e-9)


@pytest.mark.parametrize("name", scenes.FLUX_SCENES)
def test_total_flux_and_its_uncertainty(host_only, name):
    check_flux_uncertainties(name)
