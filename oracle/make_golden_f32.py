"""fp32 goldens: the REFERENCE run with AP_config.ap_dtype = torch.float32 (AP_config.py:7) on a few golden scenes
(build container only).  tests/golden/f32_<scene>.npz holds the model image(s) and a sample of the Jacobian as the
reference computes them in single precision; the fp32 profile kernels are held to them at the north star's fp32 bar
(1e-5 of the image scale).   usage: python oracle/make_golden_f32.py [scene ...]"""
import os
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.path.insert(0, HERE)
sys.path.insert(0, os.path.join(ROOT, "tests"))
from make_golden import import_reference, _datas  # noqa: E402

import numpy as np  # noqa: E402
import torch  # noqa: E402

SCENES = ["c1_sersic", "sersic_sheared", "exponential", "gaussian", "moffat", "spline", "psf_sersic", "group", "crowded",
          "moffat_psf_model", "group_up2", "psf_sersic_up2"]


def main(names):
    import scenes
    ap = import_reference()
    ap.AP_config.ap_dtype = torch.float32
    for name in names:
        model, _ = scenes.build(ap, name)
        fix = {"x_rep": model.parameters.vector_representation().detach().cpu().numpy().astype(np.float64)}
        imgs = _datas(model())
        for i, d in enumerate(imgs):
            assert d.dtype == np.float32, d.dtype
            fix[f"img{i}"] = d
        J = _datas(model.jacobian(as_representation=True))
        Jflat = np.concatenate([j.reshape(-1, j.shape[-1]) for j in J])
        idx = np.sort(np.random.default_rng(99).choice(Jflat.shape[0], size=min(2000, Jflat.shape[0]), replace=False))
        fix["jac_idx"], fix["jac_rep"] = idx, Jflat[idx]
        path = os.path.join(ROOT, "tests", "golden", f"f32_{name}.npz")
        np.savez_compressed(path, **fix)
        print(f"wrote {path}: sum={[float(d.sum()) for d in imgs]} {os.path.getsize(path) / 1e3:.0f} kB")


if __name__ == "__main__":
    main(sys.argv[1:] or SCENES)
