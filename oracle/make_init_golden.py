"""tests/golden/initialize.npz: parameter values and uncertainties after the REFERENCE's ``model.initialize()`` on the
scenes of ``scenes.INIT_SCENES`` (models built without parameter values on noisy data).  The bootstrap inside the
reference's profile fit draws from numpy's global generator: it is seeded per scene, here and in the test.
Build container only:  python oracle/make_init_golden.py"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
from make_golden import ROOT, import_reference  # noqa: E402  (also puts tests/ on the path)


def load_golden(name):
    return dict(np.load(os.path.join(ROOT, "tests", "golden", f"{name}.npz")))


def main():
    import scenes

    ap = import_reference()
    fix = {}
    for k, name in enumerate(scenes.INIT_SCENES):
        model = scenes.build_init(ap, name, scenes.init_data(name, load_golden))
        np.random.seed(900 + k)
        model.initialize()
        fix[f"{name}:value"] = model.parameters.vector_values().detach().cpu().numpy()
        fix[f"{name}:uncertainty"] = model.parameters.vector_uncertainty().detach().cpu().numpy()
        if "spline" in name:
            fix[f"{name}:prof"] = model["I(R)"].prof.detach().cpu().numpy()
        print(name, fix[f"{name}:value"], fix[f"{name}:uncertainty"])
    np.savez_compressed(os.path.join(ROOT, "tests", "golden", "initialize.npz"), **fix)


if __name__ == "__main__":
    main()
