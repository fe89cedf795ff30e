"""model.initialize() (the input prep that feeds the fit) against the reference's own start values
(tests/golden/initialize.npz, oracle/make_init_golden.py): same noisy images, models built without parameter values.
Host code only; the group scene samples its sub-models through the oracle-backed stand-in plan of
test_lm_host_logic.py.  Tolerance 1e-6: the values come out of Nelder-Mead searches on binned statistics."""
import numpy as np
import pytest

import astrophot_b200 as ap
import scenes
from conftest import load_golden
from test_lm_host_logic import host_only  # noqa: F401  (fixture)


@pytest.mark.parametrize("name", scenes.INIT_SCENES)
def test_initialize_matches_reference(host_only, name):  # noqa: F811
    fix = load_golden("initialize")
    model = scenes.build_init(ap, name, scenes.init_data(name, load_golden))
    assert name == "init_group" or not model.is_initialized
    np.random.seed(900 + scenes.INIT_SCENES.index(name))
    model.initialize()
    assert model.is_initialized
    got = model.parameters.vector_values().numpy()
    want = fix[f"{name}:value"]
    np.testing.assert_allclose(got, want, rtol=1e-6, atol=1e-9)
    np.testing.assert_allclose(model.parameters.vector_uncertainty().numpy(), fix[f"{name}:uncertainty"], rtol=1e-5,
                               atol=1e-9)
    if "spline" in name:
        np.testing.assert_allclose(model["I(R)"].prof.numpy(), fix[f"{name}:prof"], rtol=1e-12)


def test_segmentation_map_utilities_match_reference():
    """ap.utils.initialize (windows, centroids, PA, q from a segmentation map; scaling, filtering and transfer of windows)
    against the reference's own output on the same map (`utils/initialize/segmentation_map.py`)."""
    import torch
    fix = load_golden("initialize")
    seg, img = scenes.segmentation_inputs(load_golden)
    I = ap.utils.initialize
    cen = I.centroids_from_segmentation_map(seg, img)
    pas = I.PA_from_segmentation_map(seg, img, cen)
    qs = I.q_from_segmentation_map(seg, img, cen, pas)
    win = I.windows_from_segmentation_map(seg)
    ids = sorted(cen)
    assert ids == list(fix["seg:ids"]) and sorted(win) == ids and len(ids) > 10
    np.testing.assert_allclose([cen[i] for i in ids], fix["seg:centroids"], rtol=1e-12)
    np.testing.assert_allclose([pas[i] for i in ids], fix["seg:PA"], rtol=1e-10)
    np.testing.assert_allclose([qs[i] for i in ids], fix["seg:q"], rtol=1e-10)
    np.testing.assert_array_equal(np.array([win[i] for i in ids]), fix["seg:windows"])
    # defaults recompute what is not passed
    np.testing.assert_allclose([I.q_from_segmentation_map(seg, img)[i] for i in ids], fix["seg:q"], rtol=1e-10)
    scaled = I.scale_windows(win, image_shape=img.shape, expand_scale=1.5, expand_border=3)
    np.testing.assert_array_equal(np.array([scaled[i] for i in ids]), fix["seg:scaled"])
    kept = I.filter_windows(scaled, min_size=12, max_size=120, min_area=200, max_area=9000, min_flux=40.0, image=img)
    assert sorted(kept) == list(fix["seg:kept_ids"]) and 0 < len(kept) < len(ids)
    base = ap.image.Target_Image(data=img, pixelscale=1.0, zeropoint=22.5)
    other = ap.image.Target_Image(data=np.zeros((300, 280)), pixelscale=torch.tensor([[0.6, 0.1], [-0.1, 0.6]]),
                                  origin=[-5.0, 3.0], zeropoint=22.5)
    moved = I.transfer_windows(kept, base, other)
    np.testing.assert_array_equal(np.array([moved[i] for i in sorted(kept)]), fix["seg:moved"])
    # the windows feed straight into models
    k = sorted(kept)[0]
    m = ap.models.AstroPhot_Model(name="w", model_type="sersic galaxy model", target=base, window=kept[k])
    assert tuple(int(v) for v in m.window.pixel_shape) == (kept[k][0][1] - kept[k][0][0], kept[k][1][1] - kept[k][1][0])


def test_auto_variance_matches_reference():
    """variance="auto": the variance map estimated from the image itself (`utils/initialize/variance.py:12-55`,
    `target_image.py:278-292`) -- the W of the fit when the user has no variance map."""
    fix = load_golden("initialize")
    _, img = scenes.segmentation_inputs(load_golden)
    amask = np.zeros(img.shape, dtype=bool)
    amask[30:50, 100:130] = True
    I = ap.utils.initialize
    np.testing.assert_allclose(I.auto_variance(img), fix["autovar:plain"], rtol=1e-10)
    got = I.auto_variance(img, amask)
    assert np.all(np.isinf(got[amask])) and np.all(np.isfinite(got[~amask])) and got[~amask].min() > 0
    np.testing.assert_allclose(got, fix["autovar:masked"], rtol=1e-10)
    np.testing.assert_allclose(I.auto_variance(img[:15, :40]), fix["autovar:small"], rtol=1e-12)      # too small: constant
    assert np.all(I.auto_variance(np.zeros((30, 30))) == 1.0)                                         # flat: ones
    ap.AP_config.ap_device = "cpu"
    t = ap.image.Target_Image(data=img, pixelscale=1.0, zeropoint=22.5, variance="auto", mask=amask)
    np.testing.assert_allclose(t.weight.numpy(), fix["autovar:weight"], rtol=1e-10)
    assert np.all(t.weight.numpy()[amask] == 0)
    t2 = ap.image.Target_Image(data=img, pixelscale=1.0, zeropoint=22.5, weight="auto")
    np.testing.assert_allclose(t2.variance.numpy(), fix["autovar:plain"], rtol=1e-10)
