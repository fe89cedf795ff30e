"""Random scenes with SUPER-SAMPLED PSFs (psf_upscale 2 or 4; model_object.py:312-315,348-349, point_source.py:123-127,147-149,181):
groups of PSF-convolved galaxies of mixed families on their own windows, point sources (PSF image, or a PSF model on the
finer grid), unconvolved models and a sky, square or sheared pixels, bilinear / none / lanczos shifts -- sampled and
differentiated by the REFERENCE and by the oracle (through astrophot_b200.lowering).  Build container only.
python oracle/fuzz_reference_upscale.py        Recorded: 12 scenes, worst relative difference 2.2e-14 (images and Jacobians)"""
import os, sys
HERE = os.path.dirname(os.path.abspath(__file__)); sys.path[:0] = [HERE, os.path.dirname(HERE), os.path.join(os.path.dirname(HERE), "tests")]
from make_golden import import_reference, _datas
import numpy as np, torch
ref = import_reference()
import astrophot_b200 as ours, astrophot_oracle as orc
from astrophot_b200.lowering import lower
import scenes
ours.AP_config.ap_device = "cpu"
worst = 0
for k, (desc, up, build) in enumerate(scenes.upscale_fuzz_builders()):
    mr, mo = build(ref), build(ours)
    a = _datas(mr())[0]
    scene, _ = lower(mo)
    assert max(s.upscale for s in scene.sources) == up
    xo = mo.parameters.vector_values().numpy()
    xr = mr.parameters.vector_values().detach().cpu().numpy()
    assert list(mo.parameters.vector_names()) == list(mr.parameters.vector_names()) and np.array_equal(xo, xr)
    b = orc.sample(scene, xo, as_rep=False)[0]
    assert a.shape == b.shape, (a.shape, b.shape)
    e = np.abs(a - b).max() / np.abs(a).max()
    Jr = _datas(mr.jacobian())[0]; Jo = orc.jacobian(scene, xo, as_rep=False)[0]
    sc2 = np.maximum(np.abs(Jr).reshape(-1, Jr.shape[-1]).max(axis=0), 1e-300)
    ej = (np.abs(Jo - Jr).reshape(-1, Jr.shape[-1]) / sc2).max()
    worst = max(worst, e, ej)
    flag = "" if max(e, ej) < 1e-9 else "   <-- CHECK"
    print(f"{k:2d} {desc} P={len(xo)} img {e:.1e} jac {ej:.1e}{flag}", flush=True)
print("worst", worst)
