"""Random point sources drawn from a PSF *model* (point_source.py:122-140), alone and next to a sky in a group, sampled
and differentiated by the REFERENCE and by the oracle (through astrophot_b200.lowering, APB_F_AMP sources): every PSF
family (sersic, exponential, gaussian, moffat, moffat2d, spline), normalised or not, PSF parameters free or locked,
square and sheared pixels.  Build container only.   python oracle/fuzz_reference_points.py
Recorded: 36 scenes, worst relative difference 1.2e-14 (images and Jacobians); it also found the parameter order of
the moffat2d psf model (q, PA last), since fixed."""
import os, sys
HERE = os.path.dirname(os.path.abspath(__file__)); sys.path[:0] = [HERE, os.path.dirname(HERE), os.path.join(os.path.dirname(HERE), "tests")]
from make_golden import import_reference, _datas
import numpy as np, torch
ref = import_reference()
import astrophot_b200 as ours, astrophot_oracle as orc
from astrophot_b200.lowering import lower
ours.AP_config.ap_device = "cpu"
rng = np.random.default_rng(321)
worst = 0
KINDS = ["sersic", "exponential", "gaussian", "moffat", "moffat2d", "spline"]
for k in range(36):
    kind = KINDS[k % len(KINDS)]
    sheared = bool(rng.integers(0, 2))
    ps = np.array([[0.8, 0.1], [-0.05, 0.9]]) if sheared else float(rng.choice([1.0, 0.6]))
    scale = 0.85 if sheared else ps
    H, W = int(rng.integers(28, 50)), int(rng.integers(28, 50))
    normalize = bool(rng.integers(0, 2))
    in_group = bool(rng.integers(0, 2))
    lock_psf = bool(rng.integers(0, 3) == 0)
    pw = int(rng.choice([9, 13, 17]))
    if kind == "sersic": pp = dict(n=float(rng.uniform(0.6, 3.0)), Re=float(rng.uniform(1.5, 4) * scale))
    elif kind == "exponential": pp = dict(Re=float(rng.uniform(1.5, 4) * scale))
    elif kind == "gaussian": pp = dict(sigma=float(rng.uniform(1.0, 3) * scale))
    elif kind == "moffat": pp = dict(n=float(rng.uniform(1.5, 4.0)), Rd=float(rng.uniform(1.5, 4) * scale))
    elif kind == "moffat2d": pp = dict(n=float(rng.uniform(1.5, 4.0)), Rd=float(rng.uniform(1.5, 4) * scale), q=float(rng.uniform(0.5, 0.95)), PA=float(rng.uniform(0.1, 3.0)))
    else:
        prof = [0.0, 1.0 * scale, 2.5 * scale, 5.0 * scale, 9.0 * scale]
        pp = {"I(R)": {"value": [float(1.0 - 0.35 * r / scale + rng.uniform(-0.05, 0.05)) for r in prof], "prof": prof}}
    if lock_psf:
        pp = {n_: ({**v, "locked": True} if isinstance(v, dict) else {"value": v, "locked": True}) for n_, v in pp.items()}
    frac = rng.uniform(0.35, 0.65, size=2)
    flux = float(rng.uniform(0.5, 2.5))

    def build(ap):
        ptar = ap.image.PSF_Image(data=np.zeros((pw, pw)), pixelscale=ps)
        pm = ap.models.AstroPhot_Model(name=f"fp{k}", model_type=f"{kind} psf model", target=ptar, normalize_psf=normalize,
                                       parameters={n_: (dict(v) if isinstance(v, dict) else v) for n_, v in pp.items()})
        tar = ap.image.Target_Image(data=np.zeros((H, W)), pixelscale=ps, zeropoint=22.5)
        cen = (tar.window.pixel_to_plane(torch.tensor([frac[0] * W, frac[1] * H], dtype=torch.float64))).detach().cpu().numpy()
        star = ap.models.AstroPhot_Model(name=f"fs{k}", model_type="point model", target=tar, psf=pm,
                                         parameters={"center": [float(cen[0]), float(cen[1])], "flux": flux})
        if not in_group:
            return star
        sky = ap.models.AstroPhot_Model(name=f"fk{k}", model_type="flat sky model", target=tar, parameters={"F": -1.0})
        sky.initialize()
        return ap.models.AstroPhot_Model(name=f"fg{k}", model_type="group model", models=[star, sky], target=tar)

    mr, mo = build(ref), build(ours)
    a = _datas(mr())[0]
    scene, _ = lower(mo)
    xo = mo.parameters.vector_values().numpy()
    xr = mr.parameters.vector_values().detach().cpu().numpy()
    assert list(mo.parameters.vector_names()) == list(mr.parameters.vector_names()) and np.allclose(xo, xr, rtol=1e-13, atol=0)
    xo = xr          # (each package maps the pixel centre to the plane itself: equal to the last bit or two)
    b = orc.sample(scene, xo, as_rep=False)[0]
    e = np.abs(a - b).max() / np.abs(a).max()
    Jr = _datas(mr.jacobian())[0]; Jo = orc.jacobian(scene, xo, as_rep=False)[0]
    sc_ = np.maximum(np.abs(Jr).reshape(-1, Jr.shape[-1]).max(axis=0), 1e-300)
    ej = (np.abs(Jo - Jr).reshape(-1, Jr.shape[-1]) / sc_).max()
    worst = max(worst, e, ej)
    flag = "" if max(e, ej) < 1e-9 else "   <-- CHECK"
    print(f"{k:2d} {kind:11s} sheared={int(sheared)} {W}x{H} norm={int(normalize)} group={int(in_group)} locked={int(lock_psf)} P={len(xo)} img {e:.1e} jac {ej:.1e}{flag}", flush=True)
print("worst", worst)
