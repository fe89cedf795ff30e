#!/usr/bin/env python
"""A crowded-field fit written exactly as it would be for AstroPhot, with the import swapped.

    python examples/crowded_field_fit.py [--size 512] [--gpus-tiles 1x1]

Builds a synthetic field (PSF-convolved Sersic galaxies + stars + sky), perturbs the truth and fits it with
Levenberg-Marquardt on one B200; `--tiles 2x2` cuts the image into tiles (the layout a multi-GPU run deals to its
ranks: torchrun --nproc-per-node N ... with LM(distributed=True, tiles=...)).
"""
import argparse
import time

import numpy as np

import astrophot_b200 as ap          # instead of: import astrophot as ap


def main():
    p = argparse.ArgumentParser()
    p.add_argument("--size", type=int, default=512)
    p.add_argument("--tiles", default="1x1")
    args = p.parse_args()
    ny, nx = (int(v) for v in args.tiles.split("x"))
    size, rng = args.size, np.random.default_rng(3)
    n_gal, n_star = max(2, size * size // 17000), max(4, size * size // 3400)
    psf = ap.image.PSF_Image(data=ap.utils.moffat_psf(2.5, 2.0, 25, 1.0), pixelscale=1.0)

    def build(data=None, variance=None):
        kw = {} if variance is None else {"variance": variance}
        tar = ap.image.Target_Image(data=np.zeros((size, size)) if data is None else data, pixelscale=1.0,
                                    zeropoint=22.5, psf=psf, **kw)
        r = np.random.default_rng(7)
        models = []
        for k in range(n_gal):
            cx, cy = r.uniform(40, size - 40, size=2)
            models.append(ap.models.AstroPhot_Model(
                name=f"galaxy{k}", model_type="sersic galaxy model", target=tar, psf_mode="full",
                window=[[int(cx) - 40, int(cx) + 40], [int(cy) - 40, int(cy) + 40]],
                parameters={"center": [cx, cy], "q": r.uniform(0.4, 0.9), "PA": r.uniform(0, np.pi),
                            "n": r.uniform(1, 4), "Re": r.uniform(3, 9), "Ie": r.uniform(0, 1)}))
        for k in range(n_star):
            cx, cy = r.uniform(20, size - 20, size=2)
            models.append(ap.models.AstroPhot_Model(
                name=f"star{k}", model_type="point model", target=tar,
                window=[[int(cx) - 13, int(cx) + 14], [int(cy) - 13, int(cy) + 14]],
                parameters={"center": [cx, cy], "flux": r.uniform(1, 2)}))
        sky = ap.models.AstroPhot_Model(name="sky", model_type="flat sky model", target=tar, parameters={"F": -2.0})
        sky.initialize()
        models.append(sky)
        return ap.models.AstroPhot_Model(name="field", model_type="group model", models=models, target=tar,
                                         psf_mode="full")

    truth = build()().data.cpu().numpy()
    var = 0.1**2 + truth / 100.0
    model = build(truth + rng.normal(size=truth.shape) * np.sqrt(var), var)
    x0 = model.parameters.vector_representation().numpy() + 0.02 * rng.normal(size=len(model.parameters.vector_values()))
    t0 = time.time()
    res = ap.fit.LM(model, initial_state=x0, max_iter=15, verbose=0,
                    tiles=None if ny * nx == 1 else (ny, nx)).fit()
    print(f"{n_gal} galaxies + {n_star} stars on {size}x{size}, P = {len(x0)}: {res.iteration} LM iterations in "
          f"{time.time() - t0:.2f} s, chi^2/ndf {res.loss_history[0]:.3f} -> {res.loss_history[-1]:.4f} ({res.message})")
    res.update_uncertainty()
    g0 = model.models["galaxy0"]
    print("galaxy0:", {k: (float(g0[k].value.reshape(-1)[0]), float(g0[k].uncertainty.reshape(-1)[0])) for k in ("n", "Re", "Ie")})


if __name__ == "__main__":
    main()
