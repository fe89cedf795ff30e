"""tests/golden/flux_uncertainty.npz: total_flux / total_flux_uncertainty / total_magnitude / total_magnitude_uncertainty
(core_model.py:265-290) of the REFERENCE on a few golden scenes, with seeded parameter uncertainties.
Build container only:  python oracle/make_flux_golden.py"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
from make_golden import ROOT, import_reference  # noqa: E402  (also puts tests/ on the path)


def seeded_uncertainty(n, k):
    return np.random.default_rng(500 + k).uniform(0.01, 0.1, size=n)


def main():
    import torch
    import scenes

    ap = import_reference()
    fix = {}
    for k, name in enumerate(scenes.FLUX_SCENES):
        model, _ = scenes.build(ap, name)
        unc = seeded_uncertainty(len(model.parameters.vector_values()), k)
        model.parameters.vector_set_uncertainty(torch.as_tensor(unc, dtype=ap.AP_config.ap_dtype))
        vals = [model.total_flux(), model.total_flux_uncertainty(), model.total_magnitude(),
                model.total_magnitude_uncertainty()]
        fix[name] = np.array([float(v) for v in vals])
        print(name, fix[name])
    np.savez_compressed(os.path.join(ROOT, "tests", "golden", "flux_uncertainty.npz"), **fix)


if __name__ == "__main__":
    main()
