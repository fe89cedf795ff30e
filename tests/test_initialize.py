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
