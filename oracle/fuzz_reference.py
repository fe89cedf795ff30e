"""Random single-model scenes (kind, pixel scale, image size, PSF or none, parameters) sampled and differentiated by the
REFERENCE and by the oracle (through astrophot_b200.lowering): build container only.  Recorded: 40 scenes, worst
relative difference 1.7e-14 (images and Jacobians).   python oracle/fuzz_reference.py"""
import os, sys
HERE = os.path.dirname(os.path.abspath(__file__)); sys.path[:0] = [HERE, os.path.dirname(HERE), os.path.join(os.path.dirname(HERE), "tests")]
from make_golden import import_reference, _datas
import numpy as np, torch
ref = import_reference()
import astrophot_b200 as ours, astrophot_oracle as orc
from astrophot_b200.lowering import lower
import scenes
ours.AP_config.ap_device = "cpu"
rng = np.random.default_rng(123)
worst = 0
for k in range(40):
    ps = float(rng.choice([1.0, 0.8, 0.37, 1.6]))
    H, W = int(rng.integers(30, 70)), int(rng.integers(30, 70))
    kind = rng.choice(["sersic", "exponential", "gaussian", "moffat"])
    use_psf = bool(rng.integers(0, 2))
    pw = int(rng.choice([5, 7, 9, 13]))
    pars = {"center": [float(rng.uniform(0.3, 0.7) * W * ps), float(rng.uniform(0.3, 0.7) * H * ps)], "q": float(rng.uniform(0.25, 0.95)), "PA": float(rng.uniform(0, np.pi))}
    if kind == "sersic": pars.update(n=float(rng.uniform(0.5, 6.0)), Re=float(rng.uniform(1.5, 12) * ps), Ie=float(rng.uniform(-1, 2)))
    elif kind == "exponential": pars.update(Re=float(rng.uniform(1.5, 12) * ps), Ie=float(rng.uniform(-1, 2)))
    elif kind == "gaussian": pars.update(sigma=float(rng.uniform(1.0, 8) * ps), flux=float(rng.uniform(0, 3)))
    else: pars.update(n=float(rng.uniform(1.2, 4.0)), Rd=float(rng.uniform(1.5, 8) * ps), I0=float(rng.uniform(-1, 2)))
    def build(ap):
        kw = {}
        if use_psf:
            kw["psf"] = ap.image.PSF_Image(data=scenes._psf_moffat(2.5, 1.5 + 0.1 * pw, pw), pixelscale=ps)
        tar = ap.image.Target_Image(data=np.zeros((H, W)), pixelscale=ps, zeropoint=22.5, **kw)
        return ap.models.AstroPhot_Model(name=f"f{k}", model_type=f"{kind} galaxy model", target=tar, psf_mode="full" if use_psf else "none", parameters=dict(pars))
    mr, mo = build(ref), build(ours)
    a = _datas(mr())[0]
    scene, _ = lower(mo)
    xo = mo.parameters.vector_values().numpy()
    b = orc.sample(scene, xo, as_rep=False)[0]
    e = np.abs(a - b).max() / np.abs(a).max()
    Jr = _datas(mr.jacobian())[0]; Jo = orc.jacobian(scene, xo, as_rep=False)[0]
    sc_ = np.maximum(np.abs(Jr).reshape(-1, Jr.shape[-1]).max(axis=0), 1e-300)
    ej = (np.abs(Jo - Jr).reshape(-1, Jr.shape[-1]) / sc_).max()
    worst = max(worst, e, ej)
    flag = "" if max(e, ej) < 1e-9 else "   <-- CHECK"
    print(f"{k:2d} {kind:11s} ps={ps} {W}x{H} psf={pw if use_psf else 0:2d} img {e:.1e} jac {ej:.1e} {({k_: (round(v,3) if not isinstance(v, list) else [round(x,3) for x in v]) for k_, v in pars.items()}) if flag else ''}{flag}", flush=True)
print("worst", worst)
