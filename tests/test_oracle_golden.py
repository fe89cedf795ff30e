"""Pin the CPU oracle against outputs of the reference itself (tests/golden/,
written by oracle/make_golden.py).  Tolerances: model images 1e-10 relative
(north_star), Jacobians 1e-9 of the column scale, LM state/chi^2 1e-8."""
import numpy as np
import pytest

import astrophot_b200 as ap
import astrophot_oracle as orc
import scenes
from astrophot_b200.lowering import lower
from conftest import load_golden, golden_data, rel_err


@pytest.mark.parametrize("name", scenes.SAMPLE_SCENES + scenes.CPU_ONLY_SCENES)
def test_sample_matches_reference(name):
    fix = load_golden(name)
    model, _ = scenes.build(ap, name)
    np.testing.assert_allclose(model.parameters.vector_values().numpy(), fix["x_val"], rtol=0, atol=0)
    np.testing.assert_allclose(model.parameters.vector_representation().numpy(), fix["x_rep"], rtol=1e-14)
    scene, info = lower(model)
    imgs = orc.sample(scene, fix["x_val"], as_rep=False)
    for i, im in enumerate(imgs):
        ref = fix[f"img{i}"]
        assert im.shape == ref.shape
        has_psf = any(s.psf >= 0 for s in scene.sources)
        if has_psf:
            assert rel_err(im, ref) < 1e-10, name      # FFT conv noise in the reference is absolute
        else:
            np.testing.assert_allclose(im, ref, rtol=1e-10, atol=1e-10 * np.abs(ref).max() * 1e-6)


@pytest.mark.parametrize("name", scenes.ODD_UPSCALE_SCENES)
def test_odd_upscale_matches_reference_interior(name):
    """psf_upscale = 3: the reference's float32 1/3 (see scenes.ODD_UPSCALE_SCENES) bounds the agreement."""
    fix = load_golden(name)
    model, _ = scenes.build(ap, name)
    scene, info = lower(model)
    assert scene.sources[0].upscale == 3
    im = orc.sample(scene, fix["x_val"], as_rep=False)[0]
    ref = fix["img0"]
    assert np.all(ref[-1] == 0) and np.all(ref[:, -1] == 0) and np.all(im[-1] > 0)
    assert rel_err(im[:-1, :-1], ref[:-1, :-1]) < 1e-6
    # the oracle's Jacobian against central differences of its own model image
    x = fix["x_val"]
    J = orc.jacobian(scene, x, as_rep=False)[0]
    for k in range(len(x)):
        h = 1e-6 * max(1.0, abs(x[k]))
        xp, xm = x.copy(), x.copy()
        xp[k] += h
        xm[k] -= h
        fd = (orc.sample(scene, xp, as_rep=False)[0] - orc.sample(scene, xm, as_rep=False)[0]) / (2 * h)
        assert np.max(np.abs(J[..., k] - fd)) < 2e-5 * np.max(np.abs(fd)), k


@pytest.mark.parametrize("name", scenes.SAMPLE_SCENES + scenes.CPU_ONLY_SCENES)
@pytest.mark.parametrize("tag", ["rep", "nat"])
def test_jacobian_matches_reference(name, tag):
    fix = load_golden(name)
    model, _ = scenes.build(ap, name)
    scene, info = lower(model)
    x = fix["x_rep"] if tag == "rep" else fix["x_val"]
    J = orc.jacobian(scene, x, as_rep=(tag == "rep"))
    Jf = np.concatenate([j.reshape(-1, j.shape[-1]) for j in J])
    ref = fix[f"jac_{tag}"]
    got = Jf[fix["jac_idx"]]
    scale = np.maximum(np.abs(ref).max(axis=0), 1e-300)
    assert np.max(np.abs(got - ref) / scale) < 1e-9, name
    jtj = Jf.T @ Jf
    d = np.sqrt(np.maximum(np.diag(fix[f"jtj_{tag}"]), 1e-300))
    assert np.max(np.abs(jtj - fix[f"jtj_{tag}"]) / np.outer(d, d)) < 1e-9, name


@pytest.mark.parametrize("name", list(scenes.ALL_LM_SCENES))
def test_lm_matches_reference(name):
    fix = load_golden(name)
    model, _ = scenes.build(ap, name, data=golden_data(fix))
    scene, info = lower(model, for_fit=True)
    H, g, chi2, _ = orc.normal_eq(scene, fix["x0"])
    d = np.sqrt(np.diag(fix["hess0"]))
    assert np.max(np.abs(H - fix["hess0"]) / np.outer(d, d)) < 1e-9
    assert np.max(np.abs(g - fix["grad0"]) / (np.abs(fix["grad0"]).max())) < 1e-9
    res = orc.lm_fit(scene, fix["x0"], max_iter=8, relative_tolerance=0.0)
    ref_loss = fix["loss_history"]
    # compare every iteration both sides did while chi^2 still moves
    n = min(len(ref_loss), len(res["loss_history"]))
    moving = 1
    while moving < n and abs(ref_loss[moving] - ref_loss[moving - 1]) / ref_loss[moving] > 1e-12:
        moving += 1
    assert moving >= 3
    np.testing.assert_allclose(res["loss_history"][:moving], ref_loss[:moving], rtol=1e-8)
    # the damping path is decided by chi^2 comparisons between lambda-trials: it is only reproducible while chi^2 still
    # moves by much more than its rounding (the reference itself flips these decisions under a 1e-14
    # perturbation of the data)
    steady = 1
    while steady < moving and abs(ref_loss[steady] - ref_loss[steady - 1]) / ref_loss[steady] > 1e-9:
        steady += 1
    np.testing.assert_allclose(res["L_history"][:steady], fix["L_history"][:steady], rtol=1e-12)
    ref_x = fix["lambda_history"]
    for k in range(moving):
        np.testing.assert_allclose(res["lambda_history"][k], ref_x[k], rtol=1e-8, atol=1e-8)
    assert abs(min(res["loss_history"]) - ref_loss.min()) / ref_loss.min() < 1e-8
