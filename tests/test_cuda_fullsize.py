"""Parity at BASELINE.json's full sizes, where the oracle is too slow to be the checker: size-independent
properties of the path (two independent convolution kernels agree, flux linearity, pixel partitions add up,
chi^2's directional derivative equals -2 g.v, block-sparse solve = dense solve) on config[1] (1024^2, 51x51 PSF)
and on the 1024^2 crowded field with 375 sources (c3s)."""
import os
import sys

import numpy as np
import pytest
import torch

from conftest import ROOT

pytestmark = pytest.mark.gpu
sys.path.insert(0, ROOT)


def _workload(wl):
    import bench
    import astrophot_b200 as ap
    from astrophot_b200.lowering import lower

    ap.AP_config.ap_device = "cuda:0"
    truth = bench.build_workload(ap, wl, 1, None)
    t = truth().data.cpu().numpy()
    datas = [bench.make_data(t, 10)]
    model = bench.build_workload(ap, wl, 1, datas)
    scene, _ = lower(model, window=model.window, for_fit=True)
    x0 = bench.start_state(model.parameters.vector_representation().numpy(), scale=bench.start_scale(wl))
    return ap, model, scene, x0, t


def test_config1_fft_and_direct_convolution_agree_and_flux_is_linear():
    from astrophot_b200.cabi import Plan
    ap, model, scene, x0, truth = _workload("c2")
    xv = model.parameters.vector_values().numpy()
    a = Plan(scene, conv="fft").sample(xv)[0]
    b = Plan(scene, conv="direct").sample(xv)[0]
    scale = float(a.abs().max())
    assert float((a - b).abs().max()) <= 1e-10 * scale          # two independent kernels, 1024^2 x 51^2 PSF
    assert abs(float(a.sum()) - float(truth.sum())) <= 1e-12 * float(truth.sum())
    # Sersic flux is linear in 10^Ie: thresholds scale with the flux, so do all refinement decisions
    x3 = xv.copy()
    x3[-1] += np.log10(3.0)
    c = Plan(scene, conv="fft").sample(x3)[0]
    assert float((c - 3.0 * a).abs().max()) <= 1e-12 * 3.0 * scale
    # PSF convolution conserves flux away from the edges: the centred galaxy keeps > 99.9 % inside the image
    assert float(a.sum()) > 0


@pytest.mark.parametrize("wl", ["c2", "c3s"])
def test_gradient_is_the_directional_derivative_of_chi2(wl):
    """chi^2(x) = sum W (Y - M(x))^2  =>  d chi^2 / dx . v = -2 g . v with g = J^T W (Y - M) from apb_normal_eq."""
    from astrophot_b200.cabi import Plan
    ap, model, scene, x0, _ = _workload(wl)
    plan = Plan(scene)
    H, g, c2 = plan.normal_eq(x0, check=True)
    g = g.cpu().numpy()
    rng = np.random.default_rng(4)
    v = rng.normal(size=len(x0))
    v /= np.linalg.norm(v)
    eps = 1e-6
    cp = plan.chi2(x0 + eps * v)[0].item()
    cm = plan.chi2(x0 - eps * v)[0].item()
    fd = (cp - cm) / (2 * eps)
    want = -2.0 * float(g @ v)
    assert abs(fd - want) <= 2e-5 * max(abs(want), np.linalg.norm(g) * 1e-3), (fd, want)
    assert abs(plan.chi2(x0)[0].item() - c2[0].item()) <= 1e-12 * c2[0].item()


def test_crowded_field_tiles_add_up_and_sparse_solve_matches_dense():
    from astrophot_b200.cabi import Plan, lm_solve
    from astrophot_b200.lowering import tile_scene
    ap, model, scene, x0, _ = _workload("c3s")
    whole = Plan(scene)
    H0, g0, c0 = [t.clone() for t in whole.normal_eq(x0, check=True)]
    cut = Plan(tile_scene(scene, 2, 4))
    H1, g1, c1 = cut.normal_eq(x0, check=True)
    d = torch.sqrt(torch.diagonal(H0))
    assert float(((H1 - H0).abs() / torch.outer(d, d)).max()) < 1e-11
    assert float((g1 - g0).abs().max()) <= 1e-10 * float(g0.abs().max())
    assert abs(c1[0].item() - c0[0].item()) <= 1e-12 * c0[0].item()
    # damped system (lm.py:359-371): block-sparse PCG on the owner blocks of the tiled plan vs a dense solve
    for L in (1.0, 1e-3):
        A = H0 / (1.0 + L)
        A.diagonal().copy_(torch.diagonal(H0) * (1.0 + L) + L)
        want = torch.linalg.solve(A, g0)
        for plan in (whole, cut):
            h, info = plan.solve_sparse(g0, L)
            its, rel = info.tolist()
            assert rel <= 1e-12 and its < 2000
            assert float((h - want).abs().max()) <= 1e-8 * float(want.abs().max())


def test_config1_lm_history_matches_the_reference_itself():
    """BASELINE config[1] at full size against the REFERENCE's own fit (oracle/time_reference.py ran
    astrophot.fit.LM on the same seeded inputs in the build container): truth image, chi^2 and state per iteration."""
    fix_path = os.path.join(ROOT, "tests", "golden", "c2_fullsize_lm.npz")
    if not os.path.exists(fix_path):
        pytest.skip("full-size reference fixture not generated")
    fix = dict(np.load(fix_path))
    ap, model, scene, x0, truth = _workload("c2")
    np.testing.assert_allclose(x0, fix["x0"], rtol=0, atol=0)
    assert abs(truth.sum() - float(fix["truth_sum"])) <= 1e-11 * float(fix["truth_sum"])
    np.testing.assert_allclose(truth[::97, ::89], fix["truth_probe"], rtol=0, atol=1e-10 * float(truth.max()))
    n = len(fix["loss_history"]) - 1
    res = ap.fit.LM(model, initial_state=x0, max_iter=n, relative_tolerance=0.0).fit()
    np.testing.assert_allclose(res.loss_history[: n + 1], fix["loss_history"], rtol=1e-8)
    np.testing.assert_allclose(res.L_history[: n + 1], fix["L_history"], rtol=1e-12)
    np.testing.assert_allclose(np.array(res.lambda_history)[: n + 1], fix["lambda_history"], rtol=1e-8, atol=1e-8)


def test_config3_band_lm_history_matches_the_reference_itself():
    """One 2048^2 band of BASELINE config[3] (25x25 Gaussian PSF; its Jacobian is taken in 3 x 3 chunks of 683 / 682
    pixels) against astrophot.fit.LM run on the same seeded inputs in the build container."""
    import bench
    import astrophot_b200 as ap
    fix_path = os.path.join(ROOT, "tests", "golden", "c4band_fullsize_lm.npz")
    if not os.path.exists(fix_path):
        pytest.skip("full-size reference fixture not generated")
    fix = dict(np.load(fix_path))
    ap.AP_config.ap_device = "cuda:0"
    truth = bench.build_c4_band(ap, None, 0)().data.cpu().numpy()
    assert abs(truth.sum() - float(fix["truth_sum"])) <= 1e-11 * float(fix["truth_sum"])
    model = bench.build_c4_band(ap, [bench.make_data(truth, 10)], 0)
    x0 = bench.start_state(model.parameters.vector_representation().numpy(), scale=bench.start_scale("c4"))
    np.testing.assert_allclose(x0, fix["x0"], rtol=0, atol=0)
    n = len(fix["loss"]) - 1
    res = ap.fit.LM(model, initial_state=x0, max_iter=n, relative_tolerance=0.0).fit()
    assert res.info.chunked and res.planF is not None and len(res.plan.scene.sources) == 9
    np.testing.assert_allclose(res.loss_history[: n + 1], fix["loss"], rtol=1e-8)
    np.testing.assert_allclose(res.L_history[: n + 1], fix["L"], rtol=1e-12)
    np.testing.assert_allclose(np.array(res.lambda_history)[: n + 1], fix["lam"], rtol=1e-8, atol=1e-8)


def test_crowded_scale_model_lm_history_matches_the_reference_itself():
    """The 512^2 scale model of BASELINE config[2] (15 PSF-convolved Sersic + 78 point sources + sky, P = 340: the
    block-sparse PCG path) against astrophot.fit.LM run on the same seeded inputs in the build container."""
    import bench
    import astrophot_b200 as ap
    fix = dict(np.load(os.path.join(ROOT, "tests", "golden", "c3t_lm.npz")))
    ap.AP_config.ap_device = "cuda:0"
    truth = bench.build_crowded(ap, "c3t", None)().data.cpu().numpy()
    assert abs(truth.sum() - float(fix["truth_sum"])) <= 1e-11 * float(fix["truth_sum"])
    model = bench.build_crowded(ap, "c3t", bench.make_data(truth, 10))
    x0 = bench.start_state(model.parameters.vector_representation().numpy(), scale=bench.start_scale("c3t"))
    np.testing.assert_allclose(x0, fix["x0"], rtol=0, atol=0)
    n = len(fix["loss"]) - 1
    res = ap.fit.LM(model, initial_state=x0, max_iter=n, relative_tolerance=0.0).fit()
    assert len(res.pcg_iterations) > 0 and res._factor is None        # solved on the source-pair blocks
    np.testing.assert_allclose(res.loss_history[: n + 1], fix["loss"], rtol=1e-8)
    np.testing.assert_allclose(res.L_history[: n + 1], fix["L"], rtol=1e-12)
    np.testing.assert_allclose(np.array(res.lambda_history)[: n + 1], fix["lam"], rtol=1e-8, atol=1e-8)


def test_mosaic_scale_model_lm_history_matches_the_reference_itself():
    """The 512^2 scale model of BASELINE config[4] (Sersic + 12-node spline galaxies + sky, no PSF: the block-sparse
    solver, the mean integration reference of spline models inside a group) against astrophot.fit.LM run on the same
    seeded inputs in the build container (oracle/make_workload_golden.py c5t)."""
    import bench
    import astrophot_b200 as ap
    fix = dict(np.load(os.path.join(ROOT, "tests", "golden", "c5t_lm.npz")))
    ap.AP_config.ap_device = "cuda:0"
    truth = bench.build_mosaic(ap, "c5t", None)().data.cpu().numpy()
    assert abs(truth.sum() - float(fix["truth_sum"][0])) <= 1e-11 * float(fix["truth_sum"][0])
    np.testing.assert_allclose(truth[::37, ::41].reshape(-1), fix["truth_probe"], rtol=1e-10, atol=1e-16 * truth.max())
    model = bench.build_mosaic(ap, "c5t", bench.make_data(truth, 10))
    x0 = bench.start_state(model.parameters.vector_representation().numpy(), scale=bench.start_scale("c5t"))
    np.testing.assert_allclose(x0, fix["x0"], rtol=0, atol=0)
    n = len(fix["loss"]) - 1
    res = ap.fit.LM(model, initial_state=x0, max_iter=n, relative_tolerance=0.0).fit()
    np.testing.assert_allclose(res.loss_history[: n + 1], fix["loss"], rtol=1e-8)
    np.testing.assert_allclose(res.L_history[: n + 1], fix["L"], rtol=1e-12)
    np.testing.assert_allclose(np.array(res.lambda_history)[: n + 1], fix["lam"], rtol=1e-8, atol=1e-8)


def test_eight_band_joint_fit_lm_history_matches_the_reference_itself():
    """BASELINE config[3] at 160^2 per band -- 8 bands in a Target_Image_List, centre / q / PA / n / Re shared, per-band Ie
    and Gaussian PSF, P = 14 -- against astrophot.fit.LM on the same seeded inputs (oracle/make_workload_golden.py joint8)."""
    import bench
    import astrophot_b200 as ap
    fix = dict(np.load(os.path.join(ROOT, "tests", "golden", "joint8_lm.npz")))
    ap.AP_config.ap_device = "cuda:0"
    size = 160
    truth = []
    for b in range(bench.C4_BANDS):
        full = bench.build_c4(ap, None, size=size)
        truth.append(list(full.models.values())[b]().data.cpu().numpy())
    np.testing.assert_allclose([t.sum() for t in truth], fix["truth_sum"], rtol=1e-11)
    model = bench.build_c4(ap, [bench.make_data(t, 10 + b) for b, t in enumerate(truth)], size=size)
    x0 = bench.start_state(model.parameters.vector_representation().numpy(), scale=bench.start_scale("c4"))
    np.testing.assert_allclose(x0, fix["x0"], rtol=0, atol=0)
    assert len(x0) == 14            # centre (2), q, PA, n, Re shared + 8 x Ie
    n = len(fix["loss"]) - 1
    res = ap.fit.LM(model, initial_state=x0, max_iter=n, relative_tolerance=0.0).fit()
    np.testing.assert_allclose(res.loss_history[: n + 1], fix["loss"], rtol=1e-8)
    np.testing.assert_allclose(res.L_history[: n + 1], fix["L"], rtol=1e-12)
    np.testing.assert_allclose(np.array(res.lambda_history)[: n + 1], fix["lam"], rtol=1e-8, atol=1e-8)
