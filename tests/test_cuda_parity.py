"""GPU parity: the sm_100a library (through the C ABI) against the CPU oracle on
the same scene tables, and against the reference goldens.  Tolerances are the
north_star's: model images 1e-10 relative (image scale), parameters / chi^2 1e-8."""
import numpy as np
import pytest
import torch

import astrophot_b200 as ap
import astrophot_oracle as orc
import scenes
from astrophot_b200.lowering import lower
from conftest import load_golden, golden_data, rel_err

pytestmark = pytest.mark.gpu


def _plan(scene, conv=None):
    from astrophot_b200.cabi import Plan
    return Plan(scene, conv=conv)


PSF_SCENES = ["psf_sersic", "psf_sersic_noshift", "group", "group_nosky", "joint", "crowded", "aux_psf_moffat",
              "aux_psf_gauss_noshift"]


@pytest.mark.parametrize("name", PSF_SCENES)
def test_fft_convolution_path(name):
    """Force the shared-memory FFT convolution (the goldens' PSFs are small enough that the
    automatic choice is the direct tile kernel) and hold it to the same bars."""
    fix = load_golden(name)
    model, _ = scenes.build(ap, name)
    scene, info = lower(model)
    plan = _plan(scene, conv="fft")
    got = [t.cpu().numpy() for t in plan.sample(fix["x_val"], as_rep=False)]
    want = orc.sample(scene, fix["x_val"], as_rep=False)
    for i, (g, w) in enumerate(zip(got, want)):
        assert rel_err(g, w) < 1e-10, (name, "oracle")
        assert rel_err(g, fix[f"img{i}"]) < 1e-10, (name, "reference golden")
    J = [t.cpu().numpy() for t in plan.jacobian(fix["x_rep"], as_rep=True)]
    Jf = np.concatenate([j.reshape(-1, j.shape[-1]) for j in J])
    ref = fix["jac_rep"]
    rs = np.maximum(np.abs(ref).max(axis=0), 1e-300)
    assert np.max(np.abs(Jf[fix["jac_idx"]] - ref) / rs) < 1e-9, (name, "reference golden")
    Jd = [t.cpu().numpy() for t in _plan(scene, conv="direct").jacobian(fix["x_rep"], as_rep=True)]
    Jdf = np.concatenate([j.reshape(-1, j.shape[-1]) for j in Jd])
    scale = np.maximum(np.abs(Jdf).max(axis=0), 1e-300)
    assert np.max(np.abs(Jf - Jdf) / scale) < 1e-12, (name, "fft vs direct")


def test_fft_convolution_large_psf_and_lm():
    """51x51 PSF on a 200x180 image (padded stamp 252 x 232 -> transform lengths 256 = 2^8 and
    240 = 2^4.3.5): automatic choice is FFT; the LM trajectory must equal the direct-convolution one."""
    rng = np.random.default_rng(5)
    psf = ap.image.PSF_Image(data=ap.utils.moffat_psf(2.5, 3.0, 51, 1.0), pixelscale=1.0)
    pars = {"center": [101.3, 88.6], "q": 0.6, "PA": 1.0, "n": 2.5, "Re": 14.0, "Ie": 1.0}

    def build(data=None, var=None):
        kw = {} if var is None else {"variance": var}
        tar = ap.image.Target_Image(data=np.zeros((180, 200)) if data is None else data, pixelscale=1.0,
                                    zeropoint=22.5, psf=psf, **kw)
        return ap.models.AstroPhot_Model(name="big", model_type="sersic galaxy model", target=tar, psf_mode="full",
                                         parameters=dict(pars))

    m = build()
    scene, _ = lower(m)
    xv = m.parameters.vector_values().numpy()
    a = _plan(scene).sample(xv)[0].cpu().numpy()            # auto -> FFT
    b = _plan(scene, conv="direct").sample(xv)[0].cpu().numpy()
    w = orc.sample(scene, xv, as_rep=False)[0]
    assert rel_err(a, w) < 1e-10 and rel_err(b, w) < 1e-10
    assert rel_err(a, b) < 1e-13
    var = 0.01 + w / 100
    data = w + rng.normal(size=w.shape) * np.sqrt(var)
    x0 = build().parameters.vector_representation().numpy() + 0.05 * rng.normal(size=7)
    r1 = ap.fit.LM(build(data, var), initial_state=x0, max_iter=5, relative_tolerance=0.0).fit()
    r2 = ap.fit.LM(build(data, var), initial_state=x0, max_iter=5, relative_tolerance=0.0, conv="direct").fit()
    np.testing.assert_allclose(r1.loss_history[:4], r2.loss_history[:4], rtol=1e-10)
    np.testing.assert_allclose(r1.lambda_history[3], r2.lambda_history[3], rtol=1e-9, atol=1e-10)


def test_refinement_queue_grows_on_overflow():
    """Start with absurdly small refinement queues: sample / jacobian / LM must notice the sticky
    overflow flag, grow the queues (apb_plan_reserve) and still reproduce the reference."""
    from astrophot_b200.cabi import Plan
    fix = load_golden("c1_sersic")
    model, _ = scenes.build(ap, "c1_sersic")
    scene, _ = lower(model)
    plan = Plan(scene, queue_capacity=64)
    got = plan.sample(fix["x_val"], as_rep=False)[0].cpu().numpy()
    assert rel_err(got, fix["img0"]) < 1e-10
    assert plan.stats()["overflow"] == 0
    J = plan.jacobian(fix["x_rep"], as_rep=True)[0].cpu().numpy().reshape(-1, 7)
    ref = fix["jac_rep"]
    rs = np.maximum(np.abs(ref).max(axis=0), 1e-300)
    assert np.max(np.abs(J[fix["jac_idx"]] - ref) / rs) < 1e-9
    model, _ = scenes.build(ap, "c1_sersic", data=golden_data(fix))
    res = ap.fit.LM(model, initial_state=fix["x0"], max_iter=4, relative_tolerance=0.0, queue_capacity=64).fit()
    np.testing.assert_allclose(res.loss_history[:4], fix["loss_history"][:4], rtol=1e-8)
    model, _ = scenes.build(ap, "c1_sersic", data=golden_data(fix))
    from astrophot_b200 import cabi
    orig = cabi.Plan.__init__
    try:   # per-depth launches (the only path whose deeper queues can overflow), inside LM
        cabi.Plan.__init__ = lambda self, *a, **k: orig(self, *a, **{**k, "fused_integration": False})
        res = ap.fit.LM(model, initial_state=fix["x0"], max_iter=4, relative_tolerance=0.0, queue_capacity=64).fit()
    finally:
        cabi.Plan.__init__ = orig
    np.testing.assert_allclose(res.loss_history[:4], fix["loss_history"][:4], rtol=1e-8)


@pytest.mark.parametrize("name", ["c1_sersic", "sersic_sheared", "spline", "moffat_psf_model", "crowded", "point_psf_model",
                                  "point_psf_model_group"])
def test_fused_integration_equals_per_depth_launches(name):
    """k_integrate (one cooperative launch, lane-parallel Gauss-Legendre nodes) against the
    per-depth k_select / k_refine / k_reduce_level / k_scatter launches: same queues, same
    decisions, sums differ only in association order."""
    from astrophot_b200.cabi import Plan
    fix = load_golden(name)
    model, _ = scenes.build(ap, name)
    scene, _ = lower(model)
    pa, pb = Plan(scene), Plan(scene, fused_integration=False)
    a = pa.sample(fix["x_val"])
    b = pb.sample(fix["x_val"])
    for u, v in zip(a, b):
        assert rel_err(u.cpu().numpy(), v.cpu().numpy()) < 1e-13
    assert pa.stats()["queued"] == pb.stats()["queued"]
    Ja = np.concatenate([j.cpu().numpy().reshape(-1, scene.n_par) for j in pa.jacobian(fix["x_rep"], as_rep=True)])
    Jb = np.concatenate([j.cpu().numpy().reshape(-1, scene.n_par) for j in pb.jacobian(fix["x_rep"], as_rep=True)])
    scale = np.maximum(np.abs(Jb).max(axis=0), 1e-300)
    assert np.max(np.abs(Ja - Jb) / scale) < 1e-12


@pytest.mark.parametrize("name", ["c1_sersic", "sersic_sheared", "spline", "moffat", "psf_sersic", "moffat_psf_model", "group",
                                  "joint", "crowded", "aux_psf_moffat", "point_psf_model", "point_psf_model_group"])
def test_pooled_integration_equals_lane_shared(name):
    """k_integrate_pool (throughput form for long queues: a lane per cell, children of all failing
    entries pooled over the CTA) against k_integrate: same nodes, same decisions, same counts; sums
    differ only in the order the children of an entry are added."""
    from astrophot_b200.cabi import Plan
    fix = load_golden(name)
    model, _ = scenes.build(ap, name)
    scene, _ = lower(model)
    pa, pb = Plan(scene), Plan(scene, pooled_integration=True)
    a = pa.sample(fix["x_val"])
    b = pb.sample(fix["x_val"])
    for u, v in zip(a, b):
        assert rel_err(u.cpu().numpy(), v.cpu().numpy()) < 1e-13
    assert pa.stats()["queued"] == pb.stats()["queued"]
    assert sum(pa.stats()["queued"]) > 0
    Ja = np.concatenate([j.cpu().numpy().reshape(-1, scene.n_par) for j in pa.jacobian(fix["x_rep"], as_rep=True)])
    Jb = np.concatenate([j.cpu().numpy().reshape(-1, scene.n_par) for j in pb.jacobian(fix["x_rep"], as_rep=True)])
    assert pa.stats()["queued"] == pb.stats()["queued"]
    scale = np.maximum(np.abs(Jb).max(axis=0), 1e-300)
    assert np.max(np.abs(Ja - Jb) / scale) < 1e-12


@pytest.mark.parametrize("name", scenes.SAMPLE_SCENES)
def test_sample_vs_oracle_and_reference(name):
    fix = load_golden(name)
    model, _ = scenes.build(ap, name)
    scene, info = lower(model)
    plan = _plan(scene)
    got = [t.cpu().numpy() for t in plan.sample(fix["x_val"], as_rep=False)]
    want = orc.sample(scene, fix["x_val"], as_rep=False)
    has_psf = any(s.psf >= 0 for s in scene.sources)
    for i, (g, w) in enumerate(zip(got, want)):
        assert rel_err(g, w) < 1e-10, (name, "oracle")
        assert rel_err(g, fix[f"img{i}"]) < 1e-10, (name, "reference golden")
        if not has_psf:
            # no convolution in between: the north star's 1e-10 holds pixel by pixel (the reference's FFT convolution
            # has absolute, not relative, rounding noise, so PSF scenes are held to the image scale)
            ref = fix[f"img{i}"]
            np.testing.assert_allclose(g, ref, rtol=1e-10, atol=1e-16 * np.abs(ref).max(), err_msg=name)
            np.testing.assert_allclose(g, w, rtol=1e-10, atol=1e-16 * np.abs(ref).max(), err_msg=name)
    # representation-space entry gives the same image
    got2 = [t.cpu().numpy() for t in plan.sample(fix["x_rep"], as_rep=True)]
    for g, g2 in zip(got, got2):
        assert rel_err(g2, g) < 1e-12
    st = plan.stats()
    assert st["overflow"] == 0


@pytest.mark.parametrize("name,conv", [("psf_sersic_up3_direct", None), ("psf_sersic_up3_direct", "fft"),
                                       ("psf_sersic_up2", "direct"), ("psf_sersic_up2", "fft"),
                                       ("group_up2", "direct"), ("group_up2", "fft"),
                                       ("aux_psf_up2", "direct"), ("aux_psf_up2", "fft"), ("point_psf_model_up2", None)])
def test_supersampled_psf_both_convolutions(name, conv):
    """Super-sampled PSFs (model_object.py:312-315,348-349; point_source.py:147-149,181): fine-grid sampling and
    convolution, block-summed back by k_reduce_up, through both convolution kernels, against the oracle (for
    psf_upscale = 3 the oracle is the only exact checker: scenes.ODD_UPSCALE_SCENES)."""
    fix = load_golden(name)
    model, _ = scenes.build(ap, name)
    scene, info = lower(model)
    assert max(s.upscale for s in scene.sources) == (3 if "up3" in name else 2)
    plan = _plan(scene, conv=conv)
    got = plan.sample(fix["x_val"], as_rep=False)[0].cpu().numpy()
    want = orc.sample(scene, fix["x_val"], as_rep=False)[0]
    assert rel_err(got, want) < 1e-10
    if "up3" in name:
        assert rel_err(got[:-1, :-1], fix["img0"][:-1, :-1]) < 1e-6
    else:
        assert rel_err(got, fix["img0"]) < 1e-10
    for as_rep, x in ((True, fix["x_rep"]), (False, fix["x_val"])):
        J = plan.jacobian(x, as_rep=as_rep)[0].cpu().numpy()
        Jo = orc.jacobian(scene, x, as_rep=as_rep)[0]
        scale = np.maximum(np.abs(Jo).reshape(-1, Jo.shape[-1]).max(axis=0), 1e-300)
        assert np.max(np.abs(J - Jo) / scale) < 1e-9
    assert plan.stats()["overflow"] == 0


@pytest.mark.parametrize("k", range(8))
def test_supersampled_psf_random_groups(k):
    """Random groups on targets with 2x / 4x super-sampled PSFs (scenes.upscale_fuzz_builders; the oracle agrees with the
    reference on the very same scenes to 2.2e-14, oracle/fuzz_reference_upscale.py): CUDA against the oracle."""
    desc, up, build = scenes.upscale_fuzz_builders()[k]
    model = build(ap)
    scene, _ = lower(model)
    assert max(s.upscale for s in scene.sources) == up
    x = model.parameters.vector_values().numpy()
    plan = _plan(scene)
    got = plan.sample(x, as_rep=False)[0].cpu().numpy()
    want = orc.sample(scene, x, as_rep=False)[0]
    assert rel_err(got, want) < 1e-10, desc
    J = plan.jacobian(x, as_rep=False)[0].cpu().numpy()
    Jo = orc.jacobian(scene, x, as_rep=False)[0]
    scale = np.maximum(np.abs(Jo).reshape(-1, Jo.shape[-1]).max(axis=0), 1e-300)
    assert np.max(np.abs(J - Jo) / scale) < 1e-9, desc
    assert plan.stats()["overflow"] == 0


@pytest.mark.parametrize("name", scenes.SAMPLE_SCENES)
@pytest.mark.parametrize("tag", ["rep", "nat"])
def test_jacobian_vs_oracle_and_reference(name, tag):
    fix = load_golden(name)
    model, _ = scenes.build(ap, name)
    scene, info = lower(model)
    plan = _plan(scene)
    x = fix["x_rep"] if tag == "rep" else fix["x_val"]
    J = [t.cpu().numpy() for t in plan.jacobian(x, as_rep=(tag == "rep"))]
    Jo = orc.jacobian(scene, x, as_rep=(tag == "rep"))
    Jf = np.concatenate([j.reshape(-1, j.shape[-1]) for j in J])
    Jof = np.concatenate([j.reshape(-1, j.shape[-1]) for j in Jo])
    scale = np.maximum(np.abs(Jof).max(axis=0), 1e-300)
    assert np.max(np.abs(Jf - Jof) / scale) < 1e-9, (name, "oracle")
    ref = fix[f"jac_{tag}"]
    rs = np.maximum(np.abs(ref).max(axis=0), 1e-300)
    assert np.max(np.abs(Jf[fix["jac_idx"]] - ref) / rs) < 1e-9, (name, "reference golden")


@pytest.mark.parametrize("name", list(scenes.LM_SCENES))
def test_normal_equations_and_geodesic(name):
    fix = load_golden(name)
    model, _ = scenes.build(ap, name, data=golden_data(fix))
    scene, info = lower(model, for_fit=True)
    plan = _plan(scene)
    x0 = fix["x0"]
    H, g, c2 = plan.normal_eq(x0)
    H, g = H.cpu().numpy(), g.cpu().numpy()
    d = np.sqrt(np.diag(fix["hess0"]))
    assert np.max(np.abs(H - fix["hess0"]) / np.outer(d, d)) < 1e-9
    assert np.max(np.abs(g - fix["grad0"])) / np.abs(fix["grad0"]).max() < 1e-9
    Ho, go, chio, (J, Y0) = orc.normal_eq(scene, x0)
    assert abs(c2[0].item() - chio) / chio < 1e-11
    assert abs(plan.chi2(x0)[0].item() - chio) / chio < 1e-11
    # geodesic term against the oracle's dense-J formula (lm.py:401-406)
    Y, W, keep = orc.flat_targets(scene)
    h = orc.lm_solve(Ho, go, 1.0)
    dstep = 0.1
    Y1 = np.concatenate([m.reshape(-1) for m in orc.sample(scene, x0 + dstep * h)])
    r = (W * (Y0 - Y))[keep]
    rh = (W * (Y1 - Y))[keep]
    Jk = J[keep]
    rpp_o = Jk.T @ ((2 / dstep) * ((rh - r) / dstep - W[keep] * (Jk @ h)))
    plan.normal_eq(x0)   # (re)cache the stamp Jacobian at x0
    rpp = plan.geodesic(x0 + dstep * h, h, dstep).cpu().numpy()
    assert np.max(np.abs(rpp - rpp_o)) / np.abs(rpp_o).max() < 1e-7
    # damped solve
    from astrophot_b200.cabi import lm_solve
    hd = lm_solve(torch.as_tensor(Ho, device="cuda"), torch.as_tensor(go, device="cuda"), 1.0).cpu().numpy()
    np.testing.assert_allclose(hd, h, rtol=1e-9, atol=1e-12 * np.abs(h).max())


@pytest.mark.parametrize("name", ["joint", "crowded", "group", "aux_psf_moffat"])
def test_normal_equations_bit_reproducible(name):
    """Entries of J^T W J / J^T W r that several blocks add into (linked parameters of joint fits, the sky under every
    model) are gathered in a fixed order (k_block_gather), not with atomics: repeated builds and geodesic terms agree to
    the last bit."""
    fix = load_golden(name)
    model, _ = scenes.build(ap, name, data=golden_data(fix))
    scene, info = lower(model, for_fit=True)
    plan = _plan(scene)
    x0 = fix["x0"]
    H0, g0, _ = plan.normal_eq(x0)
    H0, g0 = H0.clone(), g0.clone()
    h = torch.linalg.solve(H0 + torch.diag(1.0 + torch.diagonal(H0)), g0).cpu().numpy()
    r0 = plan.geodesic(x0 + 0.1 * h, h, 0.1).clone()
    for _ in range(5):
        H, g, _ = plan.normal_eq(x0)
        assert torch.equal(H, H0) and torch.equal(g, g0)
        assert torch.equal(plan.geodesic(x0 + 0.1 * h, h, 0.1), r0)


@pytest.mark.parametrize("name", list(scenes.LM_SCENES))
def test_lm_fit_matches_reference(name):
    fix = load_golden(name)
    model, _ = scenes.build(ap, name, data=golden_data(fix))
    res = ap.fit.LM(model, initial_state=fix["x0"], max_iter=8, relative_tolerance=0.0).fit()
    ref_loss = fix["loss_history"]
    n = min(len(ref_loss), len(res.loss_history))
    moving = 1
    while moving < n and abs(ref_loss[moving] - ref_loss[moving - 1]) / ref_loss[moving] > 1e-12:
        moving += 1
    assert moving >= 3
    np.testing.assert_allclose(res.loss_history[:moving], ref_loss[:moving], rtol=1e-8)
    # the damping path is decided by chi^2 comparisons: it is only reproducible while chi^2 still moves by much more
    # than its rounding: inside an iteration successive lambda-trials are compared, whose chi^2 differ far less than the
    # iteration's gain (the reference itself flips these decisions under a 1e-14 perturbation)
    steady = 1
    while steady < moving and abs(ref_loss[steady] - ref_loss[steady - 1]) / ref_loss[steady] > 1e-6:
        steady += 1
    np.testing.assert_allclose(res.L_history[:steady], fix["L_history"][:steady], rtol=1e-12)
    for k in range(moving):
        np.testing.assert_allclose(res.lambda_history[k], fix["lambda_history"][k], rtol=1e-8, atol=1e-8)
    assert abs(min(res.loss_history) - ref_loss.min()) / ref_loss.min() < 1e-8
    # fitted parameters were written back to the model (lm.py:491)
    np.testing.assert_allclose(model.parameters.vector_representation().numpy(), res.res(), rtol=1e-10, atol=1e-10)


@pytest.mark.parametrize("name", list(scenes.LM_KWARGS_SCENES))
@pytest.mark.parametrize("fused", [True, False])
def test_lm_non_default_knobs_on_the_device(name, fused):
    """Geodesic acceleration on and another damping schedule (fit/lm.py:281-290 with acceleration != 0: the trial's
    chi^2 is taken at x + h + acceleration * a, so the concurrent chi^2 pass of apb_lm_trial is off and the second
    solve feeds the forward pass), through apb_lm_trial and through the separate calls, against the reference's fit
    with the same knobs."""
    fix = load_golden(name)
    model, _ = scenes.build(ap, name, data=golden_data(fix))
    res = ap.fit.LM(model, initial_state=fix["x0"], max_iter=6, relative_tolerance=0.0, fused_trial=fused,
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


@pytest.mark.parametrize("scene_name", scenes.FLUX_SCENES)
def test_total_flux_and_its_uncertainty(scene_name):
    """core_model.py:265-290 on the device: total_flux / total_flux_uncertainty / total_magnitude(_uncertainty)."""
    from test_lm_host_logic import check_flux_uncertainties
    check_flux_uncertainties(scene_name)


def test_getting_started_flow_on_the_device(tmp_path):
    """The tutorial flow end to end on the GPU: host start values (initialize, variance="auto") into the device LM."""
    from test_tutorial_flow import getting_started_flow
    getting_started_flow(tmp_path)


@pytest.mark.parametrize("name", ["psf_sersic", "group"])
def test_lm_trial_pieces_equal_fused_trial(name):
    """apb_lm_trial (solve, geodesic, solve, chi^2 in one call) against the same trial made of
    separate calls with torch vector algebra in between (the path distributed fits use)."""
    fix = load_golden(name)
    m1, _ = scenes.build(ap, name, data=golden_data(fix))
    m2, _ = scenes.build(ap, name, data=golden_data(fix))
    r1 = ap.fit.LM(m1, initial_state=fix["x0"], max_iter=5, relative_tolerance=0.0).fit()
    r2 = ap.fit.LM(m2, initial_state=fix["x0"], max_iter=5, relative_tolerance=0.0, fused_trial=False).fit()
    assert r1._fused_trial and r1.plan2 is not None and not r2._fused_trial
    m3, _ = scenes.build(ap, name, data=golden_data(fix))
    r3 = ap.fit.LM(m3, initial_state=fix["x0"], max_iter=5, relative_tolerance=0.0, overlap_trial=False).fit()
    assert r3._fused_trial and r3.plan2 is None
    np.testing.assert_allclose(r1.loss_history, r3.loss_history, rtol=1e-13)
    # the two-halves form sharded fits use (all-reduce between them), here on one GPU
    m4, _ = scenes.build(ap, name, data=golden_data(fix))
    r4 = ap.fit.LM(m4, initial_state=fix["x0"], max_iter=5, relative_tolerance=0.0, split_trial=True).fit()
    np.testing.assert_allclose(r1.loss_history, r4.loss_history, rtol=1e-13)
    np.testing.assert_allclose(r1.L_history, r4.L_history, rtol=1e-12)
    np.testing.assert_allclose(r1.loss_history, r2.loss_history, rtol=1e-11)
    np.testing.assert_allclose(r1.L_history, r2.L_history, rtol=1e-12)
    np.testing.assert_allclose(r1.lambda_history[-1], r2.lambda_history[-1], rtol=1e-9, atol=1e-10)


@pytest.mark.parametrize("sparse", [True, False])
def test_lm_large_system_solvers_match_reference(sparse):
    """Large systems (P > 159: crowded fields) solve the damped system by block-sparse PCG on the source-pair
    blocks (apb_lm_solve_sparse) or, as fallback, a library Cholesky factored once per (H, L); both forced
    here on the crowded golden (P = 99) and held to the reference's LM history."""
    fix = load_golden("crowded")
    m, _ = scenes.build(ap, "crowded", data=golden_data(fix))
    r = ap.fit.LM(m, initial_state=fix["x0"], max_iter=6, relative_tolerance=0.0, small_solver_max=0,
                  sparse_solver=sparse).fit()
    assert not r._fused_trial
    if sparse:
        assert r._factor is None and len(r.pcg_iterations) > 0 and max(r.pcg_iterations) < 500
    else:
        assert r._factor is not None and not r.pcg_iterations
    ref_loss = fix["loss_history"]
    n = min(len(ref_loss), len(r.loss_history))
    moving = 1
    while moving < n and abs(ref_loss[moving] - ref_loss[moving - 1]) / ref_loss[moving] > 1e-12:
        moving += 1
    assert moving >= 3
    np.testing.assert_allclose(r.loss_history[:moving], ref_loss[:moving], rtol=1e-8)
    np.testing.assert_allclose(r.L_history[:moving], fix["L_history"][:moving], rtol=1e-12)
    for k in range(moving):
        np.testing.assert_allclose(r.lambda_history[k], fix["lambda_history"][k], rtol=1e-8, atol=1e-8)


@pytest.mark.parametrize("name", ["crowded", "group", "c1_sersic"])
def test_sparse_pcg_solve_vs_dense(name):
    """apb_lm_solve_sparse against numpy's dense solve of the damped matrix (lm.py:359-371) built from the
    J^T W J of the same apb_normal_eq call."""
    fix = load_golden(name)
    model, _ = scenes.build(ap, name, data=golden_data(fix))
    scene, _ = lower(model, for_fit=True)
    plan = _plan(scene)
    H, g, _ = plan.normal_eq(fix["x0"], as_rep=True, check=True)
    Hn, gn = H.cpu().numpy(), g.cpu().numpy()
    rng = np.random.default_rng(5)
    for L in (1e-6, 1e-2, 1.0, 50.0):
        for rhs in (gn, rng.normal(size=len(gn))):
            res = plan.solve_sparse(torch.as_tensor(rhs, device="cuda"), L)
            assert res is not None
            h, info = res
            its, rel = info.tolist()
            want = orc.lm_solve(Hn, rhs, L)
            assert rel <= 1e-12 and its < 1000, (L, its, rel)
            np.testing.assert_allclose(h.cpu().numpy(), want, rtol=1e-8, atol=1e-11 * np.abs(want).max())
            # warm start from the solution at another damping (what consecutive lambda-trials do): same answer
            h2, info2 = plan.solve_sparse(torch.as_tensor(rhs, device="cuda"), L / 9.0, x0=h)
            its2, rel2 = info2.tolist()
            want2 = orc.lm_solve(Hn, rhs, L / 9.0)
            assert rel2 <= 1e-12 and its2 < 1000
            np.testing.assert_allclose(h2.cpu().numpy(), want2, rtol=1e-8, atol=1e-11 * np.abs(want2).max())
            # and from the exact solution: nothing left to do
            h3, info3 = plan.solve_sparse(torch.as_tensor(rhs, device="cuda"), L / 9.0, x0=h2)
            assert info3.tolist()[0] <= 2


def test_sparse_pcg_refuses_shared_parameters():
    fix = load_golden("joint")
    model, _ = scenes.build(ap, "joint", data=golden_data(fix))
    scene, _ = lower(model, for_fit=True)
    plan = _plan(scene)
    H, g, _ = plan.normal_eq(fix["x0"], as_rep=True, check=True)
    assert plan.solve_sparse(g, 1.0) is None


@pytest.mark.parametrize("P", [1, 31, 32, 33, 160, 333, 1000, 2049])
def test_dense_cholesky_solver(P):
    """apb_chol_factor / apb_chol_solve (the dense damped system beyond the single-CTA solver, fit/lm.py:359-371)
    against a library solve of the same matrix; one factor serves several right-hand sides; a non-finite matrix is
    reported, not solved."""
    from astrophot_b200.cabi import chol_factor, chol_solve
    g = torch.Generator(device="cuda").manual_seed(100 + P)
    J = torch.randn(2 * P + 7, P, dtype=torch.float64, device="cuda", generator=g) * torch.logspace(-2, 2, P, dtype=torch.float64, device="cuda")
    H = J.T @ J
    work = None
    for L in (1.0, 1e-3, 1e-7):
        A = H / (1.0 + L)
        d = torch.diagonal(H)
        A.diagonal().copy_(d + L * (1.0 + d))
        work, info = chol_factor(H, L, work=work)
        assert int(info.item()) == 0
        Lref = torch.linalg.cholesky(A)
        F = work[:P * P].reshape(P, P).tril()
        if L == 1.0:     # (well conditioned: the two factorisations agree far below the matrix's own rounding)
            assert float((F - Lref).abs().max() / Lref.abs().max()) < 1e-9
        for k in range(2):
            rhs = torch.randn(P, dtype=torch.float64, device="cuda", generator=g)
            x = chol_solve(work, rhs)
            resid = A @ x - rhs
            assert float(resid.abs().max()) <= 1e-10 * float((A.abs() @ x.abs()).max())
            want = torch.cholesky_solve(rhs.reshape(-1, 1), Lref).reshape(-1)
            assert float((x - want).abs().max()) <= 1e-7 * float(want.abs().max()) * max(1.0, 1e-4 / L)
            y = rhs.clone()
            chol_solve(work, y, out=y)       # in place
            assert torch.equal(x, y)
    H[P // 2, P // 2] = float("nan")
    _, info = chol_factor(H, 1.0, work=work)
    assert int(info.item()) == 1


def test_public_api_sample_and_jacobian():
    model, _ = scenes.build(ap, "group")
    img = model()
    fix = load_golden("group")
    assert rel_err(img.data.cpu().numpy(), fix["img0"]) < 1e-10
    J = model.jacobian(as_representation=True)
    assert tuple(J.data.shape) == (96, 96, 34)
    assert list(J.parameters) == list(model.parameters.vector_identities())
    # adding into a supplied image (seam 1 semantics)
    base = model.target[model.window].model_image()
    base.data += 1.0
    out = model(image=base)
    assert rel_err(out.data.cpu().numpy(), fix["img0"] + 1.0) < 1e-10


def test_unknown_modes_raise():
    tar = ap.image.Target_Image(data=np.zeros((16, 16)), pixelscale=1.0)
    m = ap.models.AstroPhot_Model(name="bad", model_type="sersic galaxy model", target=tar, sampling_mode="nope",
                                  parameters={"center": [8, 8], "q": 0.5, "PA": 0.1, "n": 1, "Re": 2, "Ie": 0})
    with pytest.raises(ap.errors.SpecificationConflict):
        m()


# ---------------------------------------------------------------------------
# one image cut into tiles: lowering.tile_scene, owner-level block-sparse J^T W J
# ---------------------------------------------------------------------------
def test_tiled_plan_with_auxiliary_psf_model():
    """Tiles + an auxiliary PSF model: the PSF grid is never cut, every tile's piece of the host uses it."""
    from astrophot_b200.lowering import tile_scene
    fix = load_golden("aux_psf_moffat")
    model, _ = scenes.build(ap, "aux_psf_moffat", data=golden_data(fix))
    scene, _ = lower(model, for_fit=True)
    cut = _plan(tile_scene(scene, 2, 2))
    H1, g1, c1 = [t.cpu().numpy() for t in cut.normal_eq(fix["x0"], check=True)]
    d = np.sqrt(np.diag(fix["hess0"]))
    assert np.max(np.abs(H1 - fix["hess0"]) / np.outer(d, d)) < 1e-9
    assert np.max(np.abs(g1 - fix["grad0"])) / np.abs(fix["grad0"]).max() < 1e-9
    assert cut.bind_blocks() is None      # PSF parameters are shared by all pieces: dense solve only


@pytest.mark.parametrize("name,tiles", [("crowded", (2, 2)), ("crowded", (3, 2)), ("group", (1, 2)), ("group_nosky", (2, 1))])
def test_tiled_plan_equals_whole_image(name, tiles):
    from astrophot_b200.lowering import tile_scene
    fix = load_golden(name)
    model, _ = scenes.build(ap, name, data=golden_data(fix))
    scene, _ = lower(model, for_fit=True)
    tiled = tile_scene(scene, *tiles)
    whole, cut = _plan(scene), _plan(tiled)
    x0 = fix["x0"]
    H0, g0, c0 = [t.cpu().numpy() for t in whole.normal_eq(x0, check=True)]
    H1, g1, c1 = [t.cpu().numpy() for t in cut.normal_eq(x0, check=True)]
    d = np.sqrt(np.diag(fix["hess0"]))
    assert np.max(np.abs(H1 - fix["hess0"]) / np.outer(d, d)) < 1e-9
    assert np.max(np.abs(H1 - H0) / np.outer(d, d)) < 1e-12
    assert np.max(np.abs(g1 - g0)) / np.abs(g0).max() < 1e-11
    assert abs(c1[0] - c0[0]) / c0[0] < 1e-12
    # model image: the tiles stitched together
    img0 = whole.sample(x0, as_rep=True)[0].cpu().numpy()
    parts = [t.cpu().numpy() for t in cut.sample(x0, as_rep=True)]
    seen = np.zeros(img0.shape, dtype=int)
    for im, part in zip(tiled.images, parts):     # (cost-balanced cuts: a tile's origin is the shift of its reference pixel)
        tx, ty = (np.round(np.asarray(scene.images[0].rij) - np.asarray(im.rij))).astype(int)
        assert rel_err(part, img0[ty:ty + im.H, tx:tx + im.W]) < 1e-12
        seen[ty:ty + im.H, tx:tx + im.W] += 1
    assert np.all(seen == 1)
    # the pieces of a model share its parameters; the sparse matrix is laid out on the owners and still solves
    rng = np.random.default_rng(11)
    for L in (1e-3, 1.0):
        for rhs in (g1, rng.normal(size=len(g1))):
            res = cut.solve_sparse(torch.as_tensor(rhs, device="cuda"), L)
            assert res is not None
            h, info = res
            its, rel = info.tolist()
            want = orc.lm_solve(H1, rhs, L)
            assert rel <= 1e-12 and its < 1000
            np.testing.assert_allclose(h.cpu().numpy(), want, rtol=1e-8, atol=1e-11 * np.abs(want).max())
    # caller-owned block array (what a tile-sharded fit all-reduces), no dense copy
    blk = cut.bind_blocks()
    assert blk is not None and blk.numel() > len(g1)
    _, g2, _ = cut.normal_eq(x0, out=(None, torch.empty_like(torch.as_tensor(g1, device="cuda")), torch.empty(2, dtype=torch.float64, device="cuda")))
    torch.cuda.synchronize()
    np.testing.assert_allclose(blk[-len(g1):].cpu().numpy(), np.diag(H1), rtol=1e-12)
    h, info = cut.solve_sparse(g2, 1.0)
    np.testing.assert_allclose(h.cpu().numpy(), orc.lm_solve(H1, g1, 1.0), rtol=1e-8, atol=1e-11)


@pytest.mark.parametrize("small", [159, 0])
def test_lm_tiled_fit_matches_reference(small):
    """LM(tiles=(2, 2)) on one GPU: same LM history as the reference's fit of the whole image, through the fused
    trial (P = 99) and through the owner-level block-sparse PCG (small_solver_max=0)."""
    fix = load_golden("crowded")
    m, _ = scenes.build(ap, "crowded", data=golden_data(fix))
    r = ap.fit.LM(m, initial_state=fix["x0"], max_iter=6, relative_tolerance=0.0, tiles=(2, 2),
                  small_solver_max=small).fit()
    assert len(r.plan.shapes) == 4
    if small == 0:
        assert r._factor is None and len(r.pcg_iterations) > 0
    ref_loss = fix["loss_history"]
    n = min(len(ref_loss), len(r.loss_history))
    moving = 1
    while moving < n and abs(ref_loss[moving] - ref_loss[moving - 1]) / ref_loss[moving] > 1e-12:
        moving += 1
    assert moving >= 3
    np.testing.assert_allclose(r.loss_history[:moving], ref_loss[:moving], rtol=1e-8)
    np.testing.assert_allclose(r.L_history[:moving], fix["L_history"][:moving], rtol=1e-12)
    for k in range(moving):
        np.testing.assert_allclose(r.lambda_history[k], fix["lambda_history"][k], rtol=1e-8, atol=1e-8)


def _nccl_tile_worker(rank, world, port, out_dir, small):
    import os, sys
    import torch.distributed as dist
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    ap.AP_config.ap_device = f"cuda:{rank}"
    fix = load_golden("crowded")
    m, _ = scenes.build(ap, "crowded", data=golden_data(fix))
    r = ap.fit.LM(m, initial_state=fix["x0"], max_iter=6, relative_tolerance=0.0, tiles=(2, 2), distributed=True,
                  small_solver_max=small).fit()
    assert len(r.plan.shapes) == 4 // world
    if small == 0:
        assert r._blk is not None and len(r.pcg_iterations) > 0 and r._factor is None
    assert r._peer is not None          # one node: the exchange is the peer-memory kernel, not NCCL
    if rank == 0:
        np.savez(os.path.join(out_dir, f"lm_{small}.npz"), loss=np.array(r.loss_history), L=np.array(r.L_history),
                 lam=np.array(r.lambda_history))
    dist.barrier()
    dist.destroy_process_group()


def _peer_allreduce_worker(rank, world, port, out_dir):
    import os
    import torch.distributed as dist
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    from astrophot_b200.cabi import PeerComm
    comm = PeerComm(1 << 20)
    g = torch.Generator(device="cuda").manual_seed(100 + rank)
    worst = 0.0
    # many back-to-back calls of changing size: the double-buffered slots and the sequence flags must never let a
    # fast rank overwrite or read a slot too early; no host synchronisation between the calls
    sizes = [7, 212, 1 << 20, 3, 4099, 513711, 1, 65536] * 25
    outs, refs = [], []
    for k, n in enumerate(sizes):
        t = torch.randn(n, dtype=torch.float64, device="cuda", generator=g) * (1.0 + k)
        ref = t.clone()
        comm.allreduce(t)
        dist.all_reduce(ref)
        worst = max(worst, float((t - ref).abs().max() / ref.abs().max()))
    assert worst <= 1e-15, worst       # two ranks: the sum has one order, the results are identical
    # every rank holds the same bits
    t = torch.randn(1000, dtype=torch.float64, device="cuda", generator=g)
    comm.allreduce(t)
    both = [torch.empty_like(t) for _ in range(world)]
    dist.all_gather(both, t)
    assert all(torch.equal(both[0], b) for b in both)
    if rank == 0:
        open(os.path.join(out_dir, "ok"), "w").write(str(worst))
    dist.barrier()
    dist.destroy_process_group()


def test_peer_memory_allreduce_over_two_gpus(tmp_path):
    """apb_allreduce (one kernel over NVLink peer memory, rank-order sum) against the NCCL all-reduce: 200 back-to-back
    calls of sizes 1 .. 2^20 doubles without host synchronisation in between."""
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    import os
    import torch.multiprocessing as mp
    port = 35500 + (os.getpid() % 2000)
    mp.spawn(_peer_allreduce_worker, args=(2, port, str(tmp_path)), nprocs=2, join=True)
    assert (tmp_path / "ok").exists()


@pytest.mark.parametrize("small", [159, 0])
def test_lm_tile_sharded_over_two_gpus(tmp_path, small):
    """Tiles dealt to 2 ranks over NCCL: dense all-reduce of J^T W J (fused two-halves trial) and the block-array
    all-reduce + replicated PCG; both against the reference's LM history."""
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    import os
    import torch.multiprocessing as mp
    port = 33500 + (os.getpid() % 2000) + (1 if small else 0)
    mp.spawn(_nccl_tile_worker, args=(2, port, str(tmp_path), small), nprocs=2, join=True)
    got = np.load(tmp_path / f"lm_{small}.npz")
    fix = load_golden("crowded")
    ref_loss = fix["loss_history"]
    n = min(len(ref_loss), len(got["loss"]))
    moving = 1
    while moving < n and abs(ref_loss[moving] - ref_loss[moving - 1]) / ref_loss[moving] > 1e-12:
        moving += 1
    assert moving >= 3
    np.testing.assert_allclose(got["loss"][:moving], ref_loss[:moving], rtol=1e-8)
    np.testing.assert_allclose(got["L"][:moving], fix["L_history"][:moving], rtol=1e-12)
    np.testing.assert_allclose(got["lam"][:moving], fix["lambda_history"][:moving], rtol=1e-8, atol=1e-8)


@pytest.mark.parametrize("name", list(scenes.ITER_SCENES))
def test_iter_fit_matches_reference(name):
    """fit.Iter (fit/iterative.py:19-180): sub-models fitted one at a time on the residual image, 3 sweeps with 4 LM
    iterations per sub-fit on both sides; chi^2 and state per sweep against the reference's."""
    fix = load_golden(name)
    model, _ = scenes.build(ap, name, data=golden_data(fix))
    res = ap.fit.Iter(model, initial_state=fix["x0"], max_iter=3,
                      method_kwargs={"max_iter": 4, "relative_tolerance": 0.0}).fit()
    assert len(res.loss_history) == len(fix["iter_loss_history"]) == 3
    np.testing.assert_allclose(res.loss_history, fix["iter_loss_history"], rtol=1e-8)
    np.testing.assert_allclose(np.array(res.lambda_history), fix["iter_lambda_history"], rtol=1e-7, atol=1e-7)
    assert res.message.startswith("fail max iterations")


@pytest.mark.parametrize("name", list(scenes.ITER_SCENES))
def test_iter_lm_fit_matches_reference(name):
    """fit.Iter_LM (fit/iterative.py:183-338): LM on sequential chunks of 8 parameters under Param_Mask, 2 sweeps,
    3 LM iterations per chunk on both sides."""
    fix = load_golden(name)
    model, _ = scenes.build(ap, name, data=golden_data(fix))
    model.parameters.vector_set_representation(torch.as_tensor(fix["x0"]))
    res = ap.fit.Iter_LM(model, initial_state=fix["x0"], chunks=8, method="sequential", max_iter=2,
                         LM_kwargs={"max_iter": 3, "relative_tolerance": 0.0}).fit()
    assert len(res.loss_history) == 2
    np.testing.assert_allclose(res.loss_history, fix["iterlm_loss_history"], rtol=1e-8)
    np.testing.assert_allclose(np.array(res.lambda_history), fix["iterlm_lambda_history"], rtol=1e-7, atol=1e-7)


@pytest.mark.parametrize("name", list(scenes.LM_SCENES))
def test_covariance_and_uncertainty_match_reference(name):
    """LM.covariance_matrix / update_uncertainty (lm.py:408-425,495-539): inverse of J^T W J in NATURAL parameter
    units at the fitted state, one more normal-equation build with as_representation=False."""
    fix = load_golden(name)
    if "cov" not in fix:
        pytest.skip("golden without covariance")
    model, _ = scenes.build(ap, name, data=golden_data(fix))
    res = ap.fit.LM(model, initial_state=fix["x0"], max_iter=8, relative_tolerance=0.0).fit()
    cov = res.covariance_matrix.cpu().numpy()
    ref = fix["cov"]
    d = np.sqrt(np.abs(np.diag(ref)))
    assert np.max(np.abs(cov - ref) / np.outer(d, d)) < 1e-6
    res.update_uncertainty()
    np.testing.assert_allclose(model.parameters.vector_uncertainty().numpy(), fix["uncertainty"], rtol=1e-6)


def test_float32_images_are_accepted_and_returned():
    """AP_config.ap_dtype = float32 (AP_config.py:7): images come in and go out as fp32 and the profile kernels compute
    in fp32; the fp32 bar of the north star is 1e-5 of the image scale."""
    fix = load_golden("psf_sersic")
    old = ap.AP_config.ap_dtype
    ap.AP_config.ap_dtype = torch.float32
    try:
        model, _ = scenes.build(ap, "psf_sersic", data=golden_data(fix))
        assert model.target.data.dtype == torch.float32
        img = model()
        assert img.data.dtype == torch.float32
        assert rel_err(img.data.double().cpu().numpy(), fix["img0"]) < 1e-5
        res = ap.fit.LM(model, initial_state=fix["x0"], max_iter=5, relative_tolerance=0.0).fit()
        np.testing.assert_allclose(res.loss_history[:4], fix["loss_history"][:4], rtol=1e-4)
    finally:
        ap.AP_config.ap_dtype = old


F32_SCENES = ["c1_sersic", "sersic_sheared", "exponential", "gaussian", "moffat", "spline", "psf_sersic", "group", "crowded",
              "moffat_psf_model", "group_up2", "psf_sersic_up2"]


@pytest.mark.parametrize("name", F32_SCENES)
def test_fp32_profile_kernels_match_the_reference_in_fp32(name):
    """AP_config.ap_dtype = float32 (AP_config.py:7): the profile kernels (first pass, mean reference, sub-pixel
    integration) compute in single precision.  Held to the REFERENCE run in fp32 (oracle/make_golden_f32.py) and to the
    fp64 goldens at the north star's fp32 bar, 1e-5 of the image scale; the reference's own fp32 run sits 2e-6 from its
    fp64 run on these scenes."""
    from astrophot_b200.cabi import Plan
    f32, f64 = load_golden(f"f32_{name}"), load_golden(name)
    model, _ = scenes.build(ap, name)
    scene, _ = lower(model)
    p32, p64 = Plan(scene, fp32=True), Plan(scene, fp32=False)
    assert p32.fp32 and not p64.fp32
    got = [t.cpu().numpy() for t in p32.sample(f64["x_val"], as_rep=False)]
    ref64 = [t.cpu().numpy() for t in p64.sample(f64["x_val"], as_rep=False)]
    differs = False
    for i, g in enumerate(got):
        assert rel_err(g, f32[f"img{i}"].astype(np.float64)) < 1e-5, (name, "reference fp32")
        assert rel_err(g, f64[f"img{i}"]) < 1e-5, (name, "reference fp64")
        differs |= bool(np.any(g != ref64[i]))
    assert differs, "the fp32 plan must not run the fp64 kernels"
    J = np.concatenate([t.cpu().numpy().reshape(-1, scene.n_par) for t in p32.jacobian(f64["x_rep"], as_rep=True)])
    ref = f64["jac_rep"]
    rs = np.maximum(np.abs(ref).max(axis=0), 1e-300)
    assert np.max(np.abs(J[f64["jac_idx"]] - ref) / rs) < 1e-4, (name, "jacobian vs reference fp64")
    ref32 = f32["jac_rep"].astype(np.float64)
    assert np.max(np.abs(J[f32["jac_idx"]] - ref32) / rs) < 1e-3, (name, "jacobian vs reference fp32")


def test_fp32_config_switch_runs_the_fp32_kernels_end_to_end():
    """The public switch: with ap_dtype = float32 model() and LM run on fp32 plans and stay within the fp32 bar."""
    fix = load_golden("psf_sersic")
    old = ap.AP_config.ap_dtype
    ap.AP_config.ap_dtype = torch.float32
    try:
        model, _ = scenes.build(ap, "psf_sersic", data=golden_data(fix))
        img = model()
        assert img.data.dtype == torch.float32
        assert rel_err(img.data.double().cpu().numpy(), fix["img0"]) < 1e-5
        lm = ap.fit.LM(model, initial_state=fix["x0"], max_iter=5, relative_tolerance=0.0)
        assert lm.plan.fp32
        res = lm.fit()
        np.testing.assert_allclose(res.loss_history[:4], fix["loss_history"][:4], rtol=1e-4)
    finally:
        ap.AP_config.ap_dtype = old


def test_plan_rebinds_image_data():
    """apb_plan_set_image_data: the same plan, pointed at other data / weight buffers, gives the normal equations of
    a plan built on those buffers (streaming many exposures through one plan)."""
    fix = load_golden("group")
    model, _ = scenes.build(ap, "group", data=golden_data(fix))
    scene, _ = lower(model, for_fit=True)
    plan = _plan(scene)
    H0, g0, c0 = [t.clone() for t in plan.normal_eq(fix["x0"], check=True)]
    d2 = plan.image_buffers[0]["data"] * 1.01 + 0.003
    w2 = plan.image_buffers[0]["weight"] * 0.9
    old = dict(plan.image_buffers[0])
    plan.set_image_data(0, d2, w2, plan._masks.get(0))
    H1, g1, c1 = [t.clone() for t in plan.normal_eq(fix["x0"], check=True)]
    import copy
    sc2 = copy.copy(scene)
    sc2.images = [copy.copy(scene.images[0])]
    sc2.images[0].data, sc2.images[0].weight = d2.cpu(), w2.cpu()
    H2, g2, c2 = _plan(sc2).normal_eq(fix["x0"], check=True)
    assert torch.allclose(H1, H2, rtol=1e-13, atol=0) and torch.allclose(g1, g2, rtol=1e-12, atol=1e-12 * float(g2.abs().max()))
    assert abs(c1[0].item() - c2[0].item()) <= 1e-13 * abs(c2[0].item())
    assert abs(c1[0].item() - c0[0].item()) > 1e-3 * abs(c0[0].item())
    plan.set_image_data(0, old["data"], old["weight"], plan._masks.get(0))
    H3, g3, c3 = plan.normal_eq(fix["x0"], check=True)
    assert torch.equal(H3, H0) and torch.equal(g3, g0)


def test_lm_rebinds_every_plan():
    """LM.set_image_data rebinds the main plan AND the forward-only twins: a fit of other data through the same LM
    object equals a fit of an LM built on those data (rebinding the main plan alone would build the normal equations
    from the new image and judge the trials on the old one)."""
    fix = load_golden("psf_sersic")
    data = golden_data(fix)
    m1, _ = scenes.build(ap, "psf_sersic", data=data)
    lm = ap.fit.LM(m1, initial_state=fix["x0"], max_iter=4, relative_tolerance=0.0)
    assert len(lm.all_plans) >= 2
    d2 = data[0]["data"] * 1.02 + 0.01
    m2, _ = scenes.build(ap, "psf_sersic", data={0: {"data": d2, "variance": data[0]["variance"]}})
    want = ap.fit.LM(m2, initial_state=fix["x0"], max_iter=4, relative_tolerance=0.0).fit()
    new = torch.as_tensor(d2, dtype=torch.float64, device="cuda")
    lm.set_image_data(0, new, lm.plan.image_buffers[0].get("weight"))
    got = lm.fit()
    np.testing.assert_allclose(got.loss_history, want.loss_history, rtol=1e-12)
    np.testing.assert_allclose(got.res(), want.res(), rtol=1e-12, atol=1e-14)


def test_speculative_lambda_search_gives_the_same_fit():
    """LM(speculate=True): the likely next lambda-trial runs on a second stream and a forward-only plan pair
    (apb_lm_trial_spec with the main plan as Jacobian donor); histories must be identical to the sequential search."""
    fix = load_golden("psf_sersic")
    m1, _ = scenes.build(ap, "psf_sersic", data=golden_data(fix))
    m2, _ = scenes.build(ap, "psf_sersic", data=golden_data(fix))
    r1 = ap.fit.LM(m1, initial_state=fix["x0"], max_iter=6, relative_tolerance=0.0).fit()
    r2 = ap.fit.LM(m2, initial_state=fix["x0"], max_iter=6, relative_tolerance=0.0, speculate=True).fit()
    assert r2._lanes is not None and r2.n_spec_hits > 0 and r1._lanes is None
    assert r1.loss_history == r2.loss_history and r1.L_history == r2.L_history
    np.testing.assert_array_equal(np.array(r1.lambda_history), np.array(r2.lambda_history))
