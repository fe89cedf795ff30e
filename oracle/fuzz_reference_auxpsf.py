"""Random galaxy models whose PSF is an auxiliary PSF *model* (model_object.py:133-147,307-310) of every PSF family
(sersic, exponential, gaussian, moffat, moffat2d, spline), PSF parameters free or locked, bilinear or no sub-pixel shift,
sampled and differentiated by the REFERENCE and by the oracle (through astrophot_b200.lowering).  Build container only.
python oracle/fuzz_reference_auxpsf.py        Recorded: 24 scenes, worst relative difference 1.2e-14 (images and Jacobians)."""
import os, sys
HERE = os.path.dirname(os.path.abspath(__file__)); sys.path[:0] = [HERE, os.path.dirname(HERE), os.path.join(os.path.dirname(HERE), "tests")]
from make_golden import import_reference, _datas
import numpy as np, torch
ref = import_reference()
import astrophot_b200 as ours, astrophot_oracle as orc
from astrophot_b200.lowering import lower
ours.AP_config.ap_device = "cpu"
rng = np.random.default_rng(77)
worst = 0
KINDS = ["sersic", "exponential", "gaussian", "moffat", "moffat2d", "spline"]
GAL = ["sersic", "exponential", "gaussian", "moffat"]
for k in range(24):
    kind = KINDS[k % len(KINDS)]
    gal = GAL[(k // 6) % 4]
    ps = float(rng.choice([1.0, 0.7]))
    H, W = int(rng.integers(36, 56)), int(rng.integers(36, 56))
    shift = "bilinear" if rng.integers(0, 2) else "none"
    lock_psf = bool(rng.integers(0, 3) == 0)
    pw = int(rng.choice([7, 9, 11]))
    if kind == "sersic": pp = dict(n=float(rng.uniform(0.6, 2.5)), Re=float(rng.uniform(1.2, 2.5) * ps))
    elif kind == "exponential": pp = dict(Re=float(rng.uniform(1.2, 2.5) * ps))
    elif kind == "gaussian": pp = dict(sigma=float(rng.uniform(0.9, 2.0) * ps))
    elif kind == "moffat": pp = dict(n=float(rng.uniform(1.5, 4.0)), Rd=float(rng.uniform(1.5, 3) * ps))
    elif kind == "moffat2d": pp = dict(n=float(rng.uniform(1.5, 4.0)), Rd=float(rng.uniform(1.5, 3) * ps), q=float(rng.uniform(0.5, 0.95)), PA=float(rng.uniform(0.1, 3.0)))
    else:
        prof = [0.0, 1.0 * ps, 2.0 * ps, 3.5 * ps, 6.0 * ps]
        pp = {"I(R)": {"value": [float(1.0 - 0.4 * r / ps + rng.uniform(-0.05, 0.05)) for r in prof], "prof": prof}}
    if lock_psf:
        pp = {n_: ({**v, "locked": True} if isinstance(v, dict) else {"value": v, "locked": True}) for n_, v in pp.items()}
    pars = {"center": [float(rng.uniform(0.4, 0.6) * W * ps), float(rng.uniform(0.4, 0.6) * H * ps)], "q": float(rng.uniform(0.3, 0.9)), "PA": float(rng.uniform(0, np.pi))}
    if gal == "sersic": pars.update(n=float(rng.uniform(0.7, 4.0)), Re=float(rng.uniform(3, 8) * ps), Ie=float(rng.uniform(-1, 1)))
    elif gal == "exponential": pars.update(Re=float(rng.uniform(3, 8) * ps), Ie=float(rng.uniform(-1, 1)))
    elif gal == "gaussian": pars.update(sigma=float(rng.uniform(2, 6) * ps), flux=float(rng.uniform(0, 2)))
    else: pars.update(n=float(rng.uniform(1.2, 3.0)), Rd=float(rng.uniform(2, 6) * ps), I0=float(rng.uniform(-1, 1)))

    def build(ap):
        ptar = ap.image.PSF_Image(data=np.zeros((pw, pw)), pixelscale=ps)
        pm = ap.models.AstroPhot_Model(name=f"ap{k}", model_type=f"{kind} psf model", target=ptar,
                                       parameters={n_: (dict(v) if isinstance(v, dict) else v) for n_, v in pp.items()})
        tar = ap.image.Target_Image(data=np.zeros((H, W)), pixelscale=ps, zeropoint=22.5)
        return ap.models.AstroPhot_Model(name=f"ag{k}", model_type=f"{gal} galaxy model", target=tar, psf_mode="full", psf=pm,
                                         psf_subpixel_shift=shift, parameters=dict(pars))

    mr, mo = build(ref), build(ours)
    a = _datas(mr())[0]
    scene, _ = lower(mo)
    xo = mo.parameters.vector_values().numpy()
    xr = mr.parameters.vector_values().detach().cpu().numpy()
    assert list(mo.parameters.vector_names()) == list(mr.parameters.vector_names()) and np.array_equal(xo, xr)
    b = orc.sample(scene, xo, as_rep=False)[0]
    e = np.abs(a - b).max() / np.abs(a).max()
    Jr = _datas(mr.jacobian())[0]; Jo = orc.jacobian(scene, xo, as_rep=False)[0]
    sc_ = np.maximum(np.abs(Jr).reshape(-1, Jr.shape[-1]).max(axis=0), 1e-300)
    ej = (np.abs(Jo - Jr).reshape(-1, Jr.shape[-1]) / sc_).max()
    worst = max(worst, e, ej)
    flag = "" if max(e, ej) < 1e-9 else "   <-- CHECK"
    print(f"{k:2d} {gal:11s} psf={kind:11s} {pw:2d} shift={shift:8s} {W}x{H} ps={ps} locked={int(lock_psf)} P={len(xo)} img {e:.1e} jac {ej:.1e}{flag}", flush=True)
print("worst", worst)
