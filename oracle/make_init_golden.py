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
    # windows / start values from a segmentation map (utils/initialize/segmentation_map.py)
    import torch
    seg, img = scenes.segmentation_inputs(load_golden)
    I = ap.utils.initialize
    cen = I.centroids_from_segmentation_map(seg, img)
    pas = I.PA_from_segmentation_map(seg, img, cen)
    qs = I.q_from_segmentation_map(seg, img, cen, pas)
    win = I.windows_from_segmentation_map(seg)
    scaled = I.scale_windows(win, image_shape=img.shape, expand_scale=1.5, expand_border=3)
    kept = I.filter_windows(scaled, min_size=12, max_size=120, min_area=200, max_area=9000, min_flux=40.0, image=img)
    base = ap.image.Target_Image(data=img, pixelscale=1.0, zeropoint=22.5)
    other = ap.image.Target_Image(data=np.zeros((300, 280)), pixelscale=torch.tensor([[0.6, 0.1], [-0.1, 0.6]]),
                                  origin=[-5.0, 3.0], zeropoint=22.5)
    moved = I.transfer_windows(kept, base, other)
    ids = sorted(cen)
    fix["seg:ids"] = np.array(ids)
    fix["seg:centroids"] = np.array([cen[i] for i in ids], dtype=np.float64)
    fix["seg:PA"] = np.array([pas[i] for i in ids], dtype=np.float64)
    fix["seg:q"] = np.array([qs[i] for i in ids], dtype=np.float64)
    fix["seg:windows"] = np.array([win[i] for i in ids], dtype=np.float64)
    fix["seg:scaled"] = np.array([scaled[i] for i in ids], dtype=np.float64)
    fix["seg:kept_ids"] = np.array(sorted(kept))
    fix["seg:moved"] = np.array([moved[i] for i in sorted(kept)], dtype=np.float64)
    print("segments", len(ids), "kept", len(kept))
    # variance map estimated from the image (utils/initialize/variance.py), with and without a mask
    amask = np.zeros(img.shape, dtype=bool)
    amask[30:50, 100:130] = True
    fix["autovar:plain"] = I.auto_variance(img)
    fix["autovar:masked"] = I.auto_variance(img, amask)
    fix["autovar:small"] = I.auto_variance(img[:15, :40])
    t = ap.image.Target_Image(data=img, pixelscale=1.0, zeropoint=22.5, variance="auto", mask=amask)
    fix["autovar:weight"] = t.weight.detach().cpu().numpy()
    np.savez_compressed(os.path.join(ROOT, "tests", "golden", "initialize.npz"), **fix)


if __name__ == "__main__":
    main()
