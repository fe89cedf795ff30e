"""The flow of the reference's GettingStarted tutorial (docs/source/tutorials/GettingStarted.ipynb) minus FITS and
plots, written as a user of the reference would: build from data with variance="auto", initialize(), LM, uncertainties,
fluxes, a windowed model, linked parameters, save / load, set the parameter vector and sample.  On the CPU the device
plan is the oracle-backed stand-in of test_lm_host_logic.py; tests/test_cuda_unverified.py runs the same flow on the GPU."""
import numpy as np
import torch

import astrophot_b200 as ap
import scenes
from conftest import load_golden
from test_lm_host_logic import host_only  # noqa: F401  (fixture)


def getting_started_flow(tmp_path, **lm_kw):
    data = scenes.init_data("init_sersic", load_golden)          # a noisy Sersic galaxy, 100 x 100
    target = ap.image.Target_Image(data=data, pixelscale=0.262, zeropoint=22.5, variance="auto")
    model = ap.models.AstroPhot_Model(name="model with target", model_type="sersic galaxy model", target=target)
    assert not model.is_initialized
    model.initialize()
    assert model.is_initialized
    result = ap.fit.LM(model, verbose=0, **lm_kw).fit()
    assert result.message.startswith("success")
    # (loss_history is chi^2 per degree of freedom, like the reference's)
    assert 0.9 < result.loss_history[-1] < 1.5 and result.loss_history[-1] < 0.8 * result.loss_history[0]
    result.update_uncertainty()
    unc = model.parameters.vector_uncertainty().numpy()
    assert np.all(unc > 0) and np.all(unc < 0.05)
    assert list(model.parameters.vector_names()) == ["center:0", "center:1", "q", "PA", "n", "Re", "Ie"]
    assert tuple(result.covariance_matrix.shape) == (7, 7)
    # the truth behind the data (scenes c1_sersic, pixelscale 0.262 instead of 1): q 0.6, PA 1.0, n 2.0
    vals = model.parameters.vector_values().numpy()
    np.testing.assert_allclose(vals[2:5], [0.6, 1.0, 2.0], atol=0.03)
    F, dF = model.total_flux().item(), model.total_flux_uncertainty().item()
    assert F > 0 and 0 < dF < 0.05 * F
    assert abs(model.total_magnitude().item() - (22.5 - 2.5 * np.log10(F))) < 1e-9
    assert 0 < model.total_magnitude_uncertainty().item() < 0.05
    assert "Re: " in str(model) and " +- " in str(model)
    # a model on a window of the image, auto-named
    sub = ap.models.AstroPhot_Model(model_type="sersic galaxy model", target=target, window=[[20, 80], [25, 75]])
    assert sub.name.startswith("sersic galaxy model [")
    sub.initialize()
    assert ap.fit.LM(sub, verbose=0, **lm_kw).fit().message.startswith("success")
    # linked parameters
    m1 = ap.models.AstroPhot_Model(model_type="sersic galaxy model", parameters={"center": [50, 50], "PA": np.pi / 4})
    m2 = ap.models.AstroPhot_Model(model_type="exponential galaxy model")
    m2["PA"].value = m1["PA"]
    m1["PA"].value = np.pi / 3
    assert m2["PA"].value.item() == m1["PA"].value.item() == np.pi / 3
    m2["PA"].value = np.pi / 2
    assert m1["PA"].value.item() == np.pi / 2
    assert m1.parameter_order == ("center", "q", "PA", "n", "Re", "Ie")
    # save, load under a new name, move the parameters, sample
    path = str(tmp_path / "AstroPhot.yaml")
    model.save(path)
    loaded = ap.models.AstroPhot_Model(name="new name", filename=path, target=target)
    np.testing.assert_array_equal(loaded.parameters.vector_values().numpy(), vals)
    np.testing.assert_array_equal(loaded.parameters.vector_uncertainty().numpy(), unc)
    loaded.initialize()
    loaded.parameters.vector_set_values(torch.tensor([13.0, 13.5, 0.4, 20 * np.pi / 180, 3, 2.5, 0.12]))
    pixels = loaded().data.detach().cpu().numpy()
    assert pixels.shape == (100, 100) and np.all(np.isfinite(pixels)) and pixels.sum() > 0
    return result


def test_getting_started_flow(host_only, tmp_path):  # noqa: F811
    getting_started_flow(tmp_path, fused_trial=False)
