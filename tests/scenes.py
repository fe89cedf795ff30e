"""Seeded synthetic scenes shared by the golden generator (which builds them with
the *reference* package) and the tests (which build them with astrophot_b200).

``build(ap, name, data=None)`` only uses API common to both packages, so the
same code constructs the reference model and ours.  ``data`` maps image index
-> dict(data=, variance=) arrays (from the golden fixture); when absent the
targets are zero images (enough for sample()/jacobian() parity).
"""
import numpy as np


def _psf_moffat(n, Rd, width):
    x = np.arange(width) - (width - 1) / 2
    X, Y = np.meshgrid(x, x, indexing="xy")
    # 5x5 sub-sampled moffat, normalised
    sub = (np.arange(5) - 2) / 5.0
    z = np.zeros((width, width))
    for a in sub:
        for b in sub:
            z += 1.0 / (1.0 + ((X + b) ** 2 + (Y + a) ** 2) / Rd**2) ** n
    return z / z.sum()


def _psf_gauss(sigma, width):
    x = np.arange(width) - (width - 1) / 2
    X, Y = np.meshgrid(x, x, indexing="xy")
    z = np.exp(-0.5 * (X**2 + Y**2) / sigma**2)
    return z / z.sum()


def _target(ap, shape, data, idx=0, pixelscale=1.0, psf=None, origin=None, mask=None, **kw):
    d = None if data is None else data.get(idx)
    arr = np.zeros(shape) if d is None else d["data"]
    var = None if d is None else d.get("variance")
    kwargs = dict(data=arr, pixelscale=pixelscale, zeropoint=22.5)
    if var is not None:
        kwargs["variance"] = var
    if psf is not None:
        kwargs["psf"] = psf
    if origin is not None:
        kwargs["origin"] = origin
    if mask is not None:
        kwargs["mask"] = mask
    kwargs.update(kw)
    return ap.image.Target_Image(**kwargs)


def build(ap, name, data=None):
    """Returns (model, meta) for scene ``name``."""
    M = ap.models.AstroPhot_Model
    if name == "c1_sersic":
        tar = _target(ap, (100, 100), data)
        m = M(name="c1", model_type="sersic galaxy model", target=tar,
              parameters={"center": [50.3, 49.6], "q": 0.6, "PA": 1.0, "n": 2.0, "Re": 10.0, "Ie": 1.0})
        return m, {}
    if name == "sersic_sheared":
        S = np.array([[0.8, 0.1], [-0.05, 0.9]])
        tar = _target(ap, (72, 80), data, pixelscale=S, origin=[3.0, -2.0])
        m = M(name="shr", model_type="sersic galaxy model", target=tar,
              parameters={"center": [33.1, 24.7], "q": 0.45, "PA": 2.1, "n": 3.1, "Re": 7.0, "Ie": 0.7})
        return m, {}
    if name == "sersic_nointegrate":
        tar = _target(ap, (64, 64), data)
        m = M(name="noint", model_type="sersic galaxy model", target=tar, integrate_mode="none",
              parameters={"center": [31.2, 33.9], "q": 0.7, "PA": 0.4, "n": 1.5, "Re": 8.0, "Ie": 0.5})
        return m, {}
    if name == "sersic_trapezoid":
        tar = _target(ap, (56, 52), data, pixelscale=0.9)
        m = M(name="trz", model_type="sersic galaxy model", target=tar, sampling_mode="trapezoid",
              parameters={"center": [22.4, 26.1], "q": 0.55, "PA": 1.3, "n": 1.8, "Re": 7.0, "Ie": 0.8})
        return m, {}
    if name == "sersic_quad5":
        tar = _target(ap, (50, 50), data, pixelscale=0.8)
        m = M(name="q5", model_type="sersic galaxy model", target=tar, sampling_mode="quad:5",
              parameters={"center": [20.1, 19.3], "q": 0.5, "PA": 0.9, "n": 1.0, "Re": 8.0, "Ie": 1.0})
        return m, {}
    if name in ("exponential", "gaussian", "moffat", "spline"):
        tar = _target(ap, (64, 60), data, pixelscale=0.9)
        common = {"center": [26.3, 30.2], "q": 0.65, "PA": 1.9}
        if name == "exponential":
            pars = dict(common, Re=6.0, Ie=0.8)
        elif name == "gaussian":
            pars = dict(common, sigma=5.0, flux=3.0)
        elif name == "moffat":
            pars = dict(common, n=2.2, Rd=4.0, I0=1.2)
        else:
            prof = [0.0, 1.5, 3.0, 5.0, 8.0, 12.0, 18.0, 26.0]
            val = [1.6, 1.45, 1.25, 0.95, 0.5, 0.0, -0.7, -1.6]
            pars = dict(common)
            pars["I(R)"] = {"value": val, "prof": prof}
        m = M(name=f"k_{name}", model_type=f"{name} galaxy model", target=tar, parameters=pars)
        return m, {}
    if name in ("psf_sersic", "psf_sersic_noshift", "psf_sersic_lanczos3"):
        psf = ap.image.PSF_Image(data=_psf_moffat(2.5, 2.0, 11), pixelscale=1.0)
        tar = _target(ap, (64, 64), data, psf=psf)
        m = M(name=name, model_type="sersic galaxy model", target=tar, psf_mode="full",
              psf_subpixel_shift={"psf_sersic": "bilinear", "psf_sersic_noshift": "none",
                                  "psf_sersic_lanczos3": "lanczos:3"}[name],
              parameters={"center": [30.8, 33.3], "q": 0.55, "PA": 2.4, "n": 2.5, "Re": 6.0, "Ie": 1.0})
        return m, {}
    if name in ("sersic_modelmask", "psf_sersic_modelmask"):
        # the model's OWN mask (model_object.py:370-371; not the target's): the model contributes nothing there
        import torch
        psf = ap.image.PSF_Image(data=_psf_moffat(2.5, 2.0, 11), pixelscale=1.0)
        tar = _target(ap, (60, 64), data, psf=psf)
        mk = np.zeros((60, 64), dtype=bool)
        mk[20:27, 30:41] = True
        mk[np.random.default_rng(77).integers(0, 60, 40), np.random.default_rng(78).integers(0, 64, 40)] = True
        m = M(name=name, model_type="sersic galaxy model", target=tar, mask=torch.as_tensor(mk),
              psf_mode="full" if name.startswith("psf") else "none",
              parameters={"center": [30.8, 31.3], "q": 0.55, "PA": 2.4, "n": 2.1, "Re": 6.0, "Ie": 1.0})
        return m, {}
    if name in ("aux_psf_moffat", "aux_psf_gauss_noshift"):
        # PSF *model* as the auxiliary PSF of a galaxy model: its parameters are fitted with the galaxy's
        # (model_object.py:133-147,307-310; BASELINE config[1] variant B)
        if name == "aux_psf_moffat":
            ptar = ap.image.PSF_Image(data=np.zeros((13, 13)), pixelscale=1.0)
            pm = M(name="auxm", model_type="moffat psf model", target=ptar, parameters={"n": 2.5, "Rd": 2.2})
        else:
            ptar = ap.image.PSF_Image(data=np.zeros((11, 11)), pixelscale=1.0)
            pm = M(name="auxg", model_type="gaussian psf model", target=ptar, parameters={"sigma": 1.4})
        tar = _target(ap, (60, 64), data)
        m = M(name=name, model_type="sersic galaxy model", target=tar, psf_mode="full", psf=pm,
              psf_subpixel_shift="bilinear" if name == "aux_psf_moffat" else "none",
              parameters={"center": [31.7, 28.4], "q": 0.6, "PA": 0.8, "n": 2.2, "Re": 6.5, "Ie": 0.9})
        return m, {}
    if name == "point":
        psf = ap.image.PSF_Image(data=_psf_moffat(2.5, 2.0, 15), pixelscale=1.0)
        tar = _target(ap, (40, 44), data, psf=psf)
        m = M(name="pt", model_type="point model", target=tar,
              parameters={"center": [20.7, 18.4], "flux": 1.0})
        return m, {}
    if name == "lanczos_group":
        # lanczos:2 / lanczos:3 sub-pixel shifts (_model_methods.py:209-227): a PSF-convolved galaxy (the wider shifted
        # stamp wraps around the padded image in the reference's FFT convolution) and two point sources, one at an edge
        psf = ap.image.PSF_Image(data=_psf_moffat(2.5, 1.8, 9), pixelscale=1.0)
        tar = _target(ap, (48, 52), data, psf=psf)
        models = [
            M(name="lz_gal", model_type="sersic galaxy model", target=tar, psf_mode="full", psf_subpixel_shift="lanczos:2",
              parameters={"center": [25.7, 22.2], "q": 0.6, "PA": 0.7, "n": 1.8, "Re": 5.0, "Ie": 0.8}),
            M(name="lz_pt", model_type="point model", target=tar, psf_subpixel_shift="lanczos:3",
              parameters={"center": [12.3, 30.8], "flux": 1.2}),
            M(name="lz_pte", model_type="point model", target=tar, psf_subpixel_shift="lanczos:3", window=[[38, 52], [0, 14]],
              parameters={"center": [49.6, 3.4], "flux": 1.5}),
        ]
        g = M(name="lzg", model_type="group model", models=models, target=tar, psf_mode="full")
        return g, {}
    if name == "point_edge":
        psf = ap.image.PSF_Image(data=_psf_moffat(2.5, 2.0, 15), pixelscale=1.0)
        tar = _target(ap, (40, 44), data, psf=psf)
        m = M(name="pte", model_type="point model", target=tar, window=[[0, 12], [25, 40]],
              parameters={"center": [3.49, 37.51], "flux": 1.3})
        return m, {}
    if name == "group":
        rng = np.random.default_rng(7)
        psf = ap.image.PSF_Image(data=_psf_moffat(2.5, 1.8, 9), pixelscale=1.0)
        tar = _target(ap, (96, 96), data, psf=psf)
        models = []
        cen = [(28.3, 30.6), (60.2, 40.1), (45.7, 70.4)]
        for k, (cx, cy) in enumerate(cen):
            models.append(M(name=f"gal{k}", model_type="sersic galaxy model", target=tar, psf_mode="full",
                            window=[[int(cx) - 16, int(cx) + 16], [int(cy) - 16, int(cy) + 16]],
                            parameters={"center": [cx, cy], "q": 0.5 + 0.1 * k, "PA": 0.7 * (k + 1),
                                        "n": 1.5 + 0.8 * k, "Re": 4.0 + k, "Ie": 0.6 + 0.1 * k}))
        for k in range(4):
            cx, cy = rng.uniform(15, 80, size=2)
            models.append(M(name=f"star{k}", model_type="point model", target=tar,
                            window=[[int(cx) - 7, int(cx) + 8], [int(cy) - 7, int(cy) + 8]],
                            parameters={"center": [cx, cy], "flux": 1.0 + 0.2 * k}))
        sky = M(name="sky", model_type="flat sky model", target=tar, parameters={"F": -1.0})
        sky.initialize()
        models.append(sky)
        g = M(name="grp", model_type="group model", models=models, target=tar, psf_mode="full")
        return g, {}
    if name == "group_nosky":
        # windows do not tile the image -> fit_mask matters
        psf = ap.image.PSF_Image(data=_psf_gauss(1.3, 7), pixelscale=1.0)
        tar = _target(ap, (70, 80), data, psf=psf)
        m1 = M(name="ga", model_type="sersic galaxy model", target=tar, psf_mode="full",
               window=[[5, 45], [8, 44]],
               parameters={"center": [25.4, 26.3], "q": 0.6, "PA": 1.1, "n": 2.0, "Re": 5.0, "Ie": 0.9})
        m2 = M(name="gb", model_type="exponential galaxy model", target=tar,
               window=[[35, 75], [30, 66]],
               parameters={"center": [55.2, 47.9], "q": 0.8, "PA": 2.5, "Re": 6.0, "Ie": 0.5})
        g = M(name="grp2", model_type="group model", models=[m1, m2], target=tar, psf_mode="full")
        return g, {}
    if name == "group_meanref":
        # compact sources with the mean-of-window integration reference inside a much larger group window: the far
        # field of each profile is below the rounding of that mean (the device skips it, the reference sums it)
        tar = _target(ap, (120, 128), data)
        m1 = M(name="mg", model_type="gaussian galaxy model", target=tar, window=[[20, 52], [30, 62]],
               parameters={"center": [36.3, 45.8], "q": 0.8, "PA": 0.6, "sigma": 1.6, "flux": 2.0})
        m2 = M(name="me", model_type="exponential galaxy model", target=tar, window=[[70, 110], [60, 100]],
               parameters={"center": [90.4, 80.1], "q": 0.7, "PA": 2.0, "Re": 1.2, "Ie": 1.5})
        prof = [0.0, 0.8, 1.6, 2.5, 3.5, 5.0]
        val = [2.0, 1.7, 1.1, 0.2, -1.0, -3.0]
        m3 = M(name="ms", model_type="spline galaxy model", target=tar, window=[[40, 80], [8, 48]],
               parameters={"center": [60.2, 27.7], "q": 0.9, "PA": 1.1, "I(R)": {"value": val, "prof": prof}})
        g = M(name="grp3", model_type="group model", models=[m1, m2, m3], target=tar)
        return g, {}
    if name == "plane_sky_group":
        # galaxy on a tilted background: `plane sky model` (planesky_model.py), natural flux units, slopes fitted
        psf = ap.image.PSF_Image(data=_psf_gauss(1.2, 7), pixelscale=0.8)
        tar = _target(ap, (72, 64), data, pixelscale=0.8, psf=psf)
        m1 = M(name="pg", model_type="sersic galaxy model", target=tar, psf_mode="full", window=[[12, 52], [16, 60]],
               parameters={"center": [25.3, 30.6], "q": 0.7, "PA": 0.5, "n": 1.7, "Re": 5.0, "Ie": 0.9})
        sky = M(name="psky", model_type="plane sky model", target=tar,
                parameters={"center": [25.6, 28.8], "F": 0.35, "delta": [0.004, -0.0025]})
        g = M(name="grp_plane", model_type="group model", models=[m1, sky], target=tar, psf_mode="full")
        return g, {}
    if name == "masked_locked_edge":
        # target mask (a block + scattered pixels), locked parameters, model window cut by the image edge so that the
        # PSF border reaches outside the image
        rng = np.random.default_rng(21)
        mask = np.zeros((56, 60), dtype=bool)
        mask[20:26, 30:41] = True
        mask[rng.integers(0, 56, 40), rng.integers(0, 60, 40)] = True
        psf = ap.image.PSF_Image(data=_psf_moffat(2.5, 1.6, 9), pixelscale=1.0)
        tar = _target(ap, (56, 60), data, psf=psf, mask=mask)
        m = M(name="mle", model_type="sersic galaxy model", target=tar, psf_mode="full", window=[[-6, 34], [18, 62]],
              parameters={"center": [9.4, 39.2], "q": 0.65, "PA": 1.2, "n": {"value": 2.0, "locked": True},
                          "Re": 6.0, "Ie": 0.8})
        return m, {}
    if name == "psf_sheared_novar":
        # general 2x2 pixelscale with a PSF (sub-pixel shift through S^-1), target without variance (weights = 1)
        S = np.array([[0.7, 0.08], [-0.05, 0.75]])
        psf = ap.image.PSF_Image(data=_psf_moffat(2.5, 1.7, 9), pixelscale=S)
        if data is not None:
            data = {k: {"data": v["data"]} for k, v in data.items()}
        tar = _target(ap, (60, 66), data, pixelscale=S, origin=[1.0, -3.0], psf=psf)
        m = M(name="psn", model_type="sersic galaxy model", target=tar, psf_mode="full",
              parameters={"center": [24.3, 17.9], "q": 0.6, "PA": 0.3, "n": 1.4, "Re": 5.5, "Ie": 1.1})
        return m, {}
    if name == "sersic_knobs":
        # non-default integration knobs: 3x3 re-gridding, two levels, 2-point Gauss-Legendre, tighter tolerance
        tar = _target(ap, (64, 64), data)
        m = M(name="knb", model_type="sersic galaxy model", target=tar, integrate_gridding=3, integrate_max_depth=2,
              integrate_quad_level=2, sampling_tolerance=3e-3,
              parameters={"center": [30.6, 33.2], "q": 0.5, "PA": 2.0, "n": 3.5, "Re": 6.0, "Ie": 0.7})
        return m, {}
    if name == "group_edge":
        # a group whose sub-model windows stick out of the image on two sides (group window = their union)
        psf = ap.image.PSF_Image(data=_psf_gauss(1.1, 7), pixelscale=1.0)
        tar = _target(ap, (64, 72), data, psf=psf)
        m1 = M(name="ge1", model_type="sersic galaxy model", target=tar, psf_mode="full", window=[[-10, 30], [20, 70]],
               parameters={"center": [8.3, 50.6], "q": 0.7, "PA": 0.4, "n": 2.2, "Re": 5.0, "Ie": 0.9})
        m2 = M(name="ge2", model_type="exponential galaxy model", target=tar, window=[[40, 80], [-6, 30]],
               parameters={"center": [60.4, 10.7], "q": 0.6, "PA": 2.2, "Re": 4.0, "Ie": 0.7})
        g = M(name="grp_edge", model_type="group model", models=[m1, m2], target=tar, psf_mode="full")
        return g, {}
    if name == "joint":
        tars, models = [], []
        for b in range(3):
            psf = ap.image.PSF_Image(data=_psf_gauss(1.2 + 0.1 * b, 9), pixelscale=1.0)
            tars.append(_target(ap, (48, 48), data, idx=b, psf=psf))
        tlist = ap.image.Target_Image_List(tars)
        for b in range(3):
            pars = {"center": [23.6, 24.3], "q": 0.6, "PA": 1.0, "n": 2.0, "Re": 6.0, "Ie": 0.3 + 0.1 * b}
            m = M(name=f"band{b}", model_type="sersic galaxy model", target=tars[b], psf_mode="full", parameters=pars)
            if b > 0:
                for p in ("center", "q", "PA", "n", "Re"):
                    m[p].value = models[0][p]
            models.append(m)
        g = M(name="joint", model_type="group model", models=models, target=tlist, psf_mode="full")
        return g, {}
    if name == "crowded":
        # scale model of BASELINE config[2]: overlapping PSF-convolved Sersics + point sources + sky
        rng = np.random.default_rng(3)
        size = 192
        psf = ap.image.PSF_Image(data=_psf_moffat(2.5, 2.0, 15), pixelscale=1.0)
        tar = _target(ap, (size, size), data, psf=psf)
        models = []
        for k in range(8):
            cx, cy = rng.uniform(30, size - 30, size=2)
            models.append(M(name=f"g{k}", model_type="sersic galaxy model", target=tar, psf_mode="full",
                            window=[[int(cx) - 24, int(cx) + 24], [int(cy) - 24, int(cy) + 24]],
                            parameters={"center": [cx, cy], "q": rng.uniform(0.4, 0.9), "PA": rng.uniform(0, np.pi),
                                        "n": rng.uniform(1, 4), "Re": rng.uniform(3, 8), "Ie": rng.uniform(0, 1)}))
        for k in range(14):
            cx, cy = rng.uniform(12, size - 12, size=2)
            models.append(M(name=f"p{k}", model_type="point model", target=tar,
                            window=[[int(cx) - 8, int(cx) + 9], [int(cy) - 8, int(cy) + 9]],
                            parameters={"center": [cx, cy], "flux": rng.uniform(1, 2)}))
        sky = M(name="csky", model_type="flat sky model", target=tar, parameters={"F": -2.0})
        sky.initialize()
        models.append(sky)
        g = M(name="crowd", model_type="group model", models=models, target=tar, psf_mode="full")
        return g, {}
    if name in ("psf_sersic_up2", "psf_sersic_up3_direct", "psf_sersic_up2_chunked"):
        # super-sampled PSF (model_object.py:313-314,348-349): PSF pixels 1/2 (1/3) of the image's; the model is sampled,
        # integrated and convolved on the fine grid, then block-summed back
        up = 3 if name == "psf_sersic_up3_direct" else 2
        # (_chunked: the Jacobian is taken chunk by chunk, _model_methods.py:349-395, each chunk on its own fine grid)
        extra = {"image_chunksize": 20} if name.endswith("_chunked") else {}
        psf = ap.image.PSF_Image(data=_psf_moffat(2.5, 1.8 * up, 6 * up + 1), pixelscale=1.0 / up)
        tar = _target(ap, (44, 48), data, psf=psf)
        m = M(name=name, model_type="sersic galaxy model", target=tar, psf_mode="full",
              psf_convolve_mode="fft" if up == 2 else "direct", **extra,
              parameters={"center": [22.8, 20.3], "q": 0.55, "PA": 2.4, "n": 2.5, "Re": 5.0, "Ie": 1.0})
        return m, {}
    if name == "group_up2":
        # a group on a target with a 2x super-sampled PSF (sheared pixels): galaxy with its own window, a point source,
        # another cut by the image edge (point_source.py:145-175 with psf_upscale), an unconvolved galaxy and the sky
        S = np.array([[0.8, 0.06], [-0.04, 0.85]])
        psf = ap.image.PSF_Image(data=_psf_moffat(2.5, 3.4, 13), pixelscale=S / 2)
        tar = _target(ap, (60, 64), data, pixelscale=S, origin=[2.0, -1.0], psf=psf)
        models = [
            M(name="u_gal", model_type="sersic galaxy model", target=tar, psf_mode="full", window=[[8, 44], [10, 50]],
              parameters={"center": [22.3, 24.1], "q": 0.6, "PA": 0.7, "n": 1.8, "Re": 4.0, "Ie": 0.9}),
            M(name="u_pt", model_type="point model", target=tar, window=[[34, 49], [30, 45]],
              parameters={"center": [35.7, 31.2], "flux": 1.2}),
            M(name="u_pte", model_type="point model", target=tar, window=[[50, 64], [0, 12]],
              parameters={"center": [50.1, 2.2], "flux": 1.5}),
            M(name="u_exp", model_type="exponential galaxy model", target=tar, window=[[30, 64], [36, 60]],
              parameters={"center": [40.2, 41.5], "q": 0.8, "PA": 2.0, "Re": 3.0, "Ie": 0.6}),
        ]
        sky = M(name="u_sky", model_type="flat sky model", target=tar, parameters={"F": -1.2})
        sky.initialize()
        models.append(sky)
        g = M(name="grp_up2", model_type="group model", models=models, target=tar, psf_mode="full")
        return g, {}
    if name == "point_psf_model_up2":
        # point source drawn from a PSF model whose grid is 2x finer than the image (point_source.py:122-143,181)
        ptar = ap.image.PSF_Image(data=np.zeros((25, 25)), pixelscale=0.4)
        pm = M(name="ppu_psf", model_type="moffat psf model", target=ptar, parameters={"n": 2.2, "Rd": 2.0})
        tar = _target(ap, (36, 40), data, pixelscale=0.8)
        m = M(name="ppu", model_type="point model", target=tar, psf=pm,
              parameters={"center": [15.3, 13.9], "flux": 1.2})
        return m, {}
    if name == "aux_psf_up2":
        # galaxy convolved with a PSF *model* sampled on a grid 2x finer than the image; PSF width fitted with the galaxy
        ptar = ap.image.PSF_Image(data=np.zeros((15, 15)), pixelscale=0.5)
        pm = M(name="auxu", model_type="gaussian psf model", target=ptar, parameters={"sigma": 1.1})
        tar = _target(ap, (44, 48), data)
        m = M(name=name, model_type="sersic galaxy model", target=tar, psf_mode="full", psf=pm,
              parameters={"center": [23.7, 20.4], "q": 0.6, "PA": 0.8, "n": 2.2, "Re": 5.5, "Ie": 0.9})
        return m, {}
    if name == "moffat_psf_model":
        # a PSF model fitted to a star cut-out held as a PSF_Image (no variance: unit weights)
        ptar = ap.image.PSF_Image(data=np.zeros((25, 25)) if data is None else data[0]["data"], pixelscale=1.0)
        m = M(name="mpsf", model_type="moffat psf model", target=ptar, parameters={"n": 2.5, "Rd": 3.0})
        return m, {}
    if name == "gaussian_psf_model":
        ptar = ap.image.PSF_Image(data=np.zeros((21, 21)), pixelscale=1.0)
        m = M(name="gpsf", model_type="gaussian psf model", target=ptar, parameters={"sigma": 1.5})
        return m, {}
    if name == "point_psf_model":
        # point source drawn from a PSF *model* (point_source.py:122-140): PSF parameters fitted with centre and flux
        ptar = ap.image.PSF_Image(data=np.zeros((15, 15)), pixelscale=0.8)
        pm = M(name="ppm_psf", model_type="moffat psf model", target=ptar, parameters={"n": 2.2, "Rd": 2.4})
        tar = _target(ap, (40, 44), data, pixelscale=0.8)
        m = M(name="ppm", model_type="point model", target=tar, psf=pm,
              parameters={"center": [17.3, 14.9], "flux": 1.2})
        return m, {}
    if name == "point_psf_model_group":
        # two stars sharing one (normalised) Moffat PSF model, a third with an un-normalised Gaussian PSF model on its
        # own small window, a galaxy and the sky; the shared PSF parameters are fitted with everything else
        ptar = ap.image.PSF_Image(data=np.zeros((17, 17)), pixelscale=1.0)
        pm = M(name="shared_psf", model_type="moffat psf model", target=ptar, parameters={"n": 2.6, "Rd": 2.1})
        ptar2 = ap.image.PSF_Image(data=np.zeros((13, 13)), pixelscale=1.0)
        pg = M(name="gauss_psf", model_type="gaussian psf model", target=ptar2, normalize_psf=False,
               parameters={"sigma": 1.6, "flux": 0.0})
        tar = _target(ap, (56, 60), data)
        models = [
            M(name="starA", model_type="point model", target=tar, psf=pm,
              parameters={"center": [18.4, 20.7], "flux": 2.0}),
            M(name="starB", model_type="point model", target=tar, psf=pm,
              parameters={"center": [41.6, 33.2], "flux": 1.6}),
            M(name="starC", model_type="point model", target=tar, psf=pg, window=[[22, 41], [38, 55]],
              parameters={"center": [31.3, 46.8], "flux": 1.4}),
            M(name="ppg_gal", model_type="sersic galaxy model", target=tar,
              parameters={"center": [30.2, 14.6], "q": 0.6, "PA": 1.1, "n": 1.7, "Re": 5.0, "Ie": 0.2}),
        ]
        sky = M(name="ppg_sky", model_type="flat sky model", target=tar, parameters={"F": -1.5})
        sky.initialize()
        models.append(sky)
        g = M(name="ppg", model_type="group model", models=models, target=tar)
        return g, {}
    raise KeyError(name)


SAMPLE_SCENES = ["c1_sersic", "sersic_sheared", "sersic_nointegrate", "sersic_quad5", "exponential", "gaussian",
                 "moffat", "spline", "psf_sersic", "psf_sersic_noshift", "point", "point_edge", "group",
                 "group_nosky", "joint", "moffat_psf_model", "gaussian_psf_model", "crowded", "aux_psf_moffat",
                 "aux_psf_gauss_noshift", "sersic_trapezoid", "group_meanref", "plane_sky_group", "masked_locked_edge", "psf_sheared_novar", "sersic_knobs",
                 "point_psf_model", "point_psf_model_group", "group_edge", "sersic_modelmask", "psf_sersic_modelmask",
                 "psf_sersic_lanczos3", "lanczos_group", "psf_sersic_up2", "group_up2", "point_psf_model_up2", "aux_psf_up2"]
# psf_upscale that is not a power of two: the reference forms 1 / psf_upscale in float32 (model_object.py:313-314 ->
# window_object.py:233-239), which perturbs the fine pixel scale by 3e-8 and floors the fine window one pixel short, so
# that the last image row and column of the model stay empty.  astrophot_b200 (and the oracle) use the exact grid; the
# golden pins the interior at 1e-6.
ODD_UPSCALE_SCENES = ["psf_sersic_up3_direct"]
CPU_ONLY_SCENES = ["psf_sersic_up2_chunked"]       # oracle vs reference only (added after the round's GPU budget was spent)
# scenes with an LM golden (noise seed, start perturbation)
LM_SCENES = {"c1_sersic": 1, "psf_sersic": 4, "group": 6, "joint": 7, "group_nosky": 8, "crowded": 9, "aux_psf_moffat": 12,
             "plane_sky_group": 13, "masked_locked_edge": 14, "psf_sheared_novar": 15,
             "moffat_psf_model": 16, "point_psf_model": 17, "point_psf_model_group": 18,
             "psf_sersic_modelmask": 19, "psf_sersic_lanczos3": 20, "lanczos_group": 21,
             "psf_sersic_up2": 22, "group_up2": 23, "point_psf_model_up2": 24, "aux_psf_up2": 25}
CPU_LM_SCENES = {"psf_sersic_up2_chunked": 26}     # (the reference's own LM cannot fit group_edge: its Group_Model.fit_mask mis-sizes windows that stick out)
ALL_LM_SCENES = {**LM_SCENES, **CPU_LM_SCENES}
NOISE_SCALE = {"moffat_psf_model": 0.02}      # noise of make_data relative to the default recipe


# total_flux / total_flux_uncertainty / total_magnitude(_uncertainty) fixture (oracle/make_flux_golden.py)
FLUX_SCENES = ["c1_sersic", "psf_sersic", "point", "group", "plane_sky_group", "moffat", "spline"]   # (the reference cannot do image lists)
ITER_SCENES = ("group", "group_nosky")     # also fitted with fit.Iter in the goldens
LM_KWARGS_SCENES = ("psf_sersic",)         # also fitted with non-default LM knobs
LM_KWARGS = dict(acceleration=0.7, curvature_limit=0.6, Lup=7.0, Ldn=5.0, L0=3.0, max_step_iter=8)


def make_data(truth_images, seed, scale=1.0):
    """Noisy data + variance from noiseless truth (tests/utils.py:73 recipe)."""
    rng = np.random.default_rng(seed)
    out = {}
    for i, t in enumerate(truth_images):
        var = (0.1 * scale) ** 2 + scale * t / 100.0
        out[i] = {"data": t + rng.normal(size=t.shape) * np.sqrt(var), "variance": var}
    return out


def perturb(x_rep, seed, scale=0.05):
    rng = np.random.default_rng(1000 + seed)
    return np.asarray(x_rep, dtype=np.float64) + scale * rng.normal(size=len(x_rep))


# ---------------------------------------------------------------------------
# random scenes with super-sampled PSFs (oracle/fuzz_reference_upscale.py: reference vs oracle; test_cuda_parity: CUDA vs oracle)
# ---------------------------------------------------------------------------
def upscale_fuzz_builders(n=12, seed=4242):
    """[(description, build(ap) -> model)]: groups of PSF-convolved galaxies of mixed families on their own windows, point
    sources (PSF image, or a PSF model on the finer grid), unconvolved models and a sky on a target whose PSF has pixels
    1/2 or 1/4 of the image's (model_object.py:312-315,348-349, point_source.py:123-127,147-149,181)."""
    rng = np.random.default_rng(seed)
    out = []
    for k in range(n):
        up = int(rng.choice([2, 4]))
        sheared = bool(rng.integers(0, 2))
        S = np.array([[0.8, 0.07], [-0.05, 0.9]]) if sheared else np.eye(2) * float(rng.choice([1.0, 0.5]))
        H, W = int(rng.integers(40, 60)), int(rng.integers(40, 60))
        pw = int(rng.choice([5, 7])) * up + 1
        pw += 1 - pw % 2
        psf_model_points = bool(rng.integers(0, 2))
        specs = []
        for g in range(int(rng.integers(1, 4))):
            fam = str(rng.choice(["sersic", "exponential", "gaussian", "moffat"]))
            cx, cy = rng.uniform(8, W - 8), rng.uniform(8, H - 8)
            half = int(rng.integers(8, 16))
            win = [[max(0, int(cx) - half), min(W, int(cx) + half)], [max(0, int(cy) - half), min(H, int(cy) + half)]]
            c = S @ np.array([cx, cy])
            pars = {"center": [float(c[0]), float(c[1])], "q": float(rng.uniform(0.3, 0.9)), "PA": float(rng.uniform(0, np.pi))}
            sc_ = float(np.sqrt(abs(np.linalg.det(S))))
            if fam == "sersic":
                pars.update(n=float(rng.uniform(0.7, 4.0)), Re=float(rng.uniform(2, 5) * sc_), Ie=float(rng.uniform(-0.5, 1)))
            elif fam == "exponential":
                pars.update(Re=float(rng.uniform(2, 5) * sc_), Ie=float(rng.uniform(-0.5, 1)))
            elif fam == "gaussian":
                pars.update(sigma=float(rng.uniform(1.5, 3) * sc_), flux=float(rng.uniform(0.5, 2)))
            else:
                pars.update(n=float(rng.uniform(1.2, 3.0)), Rd=float(rng.uniform(2, 4) * sc_), I0=float(rng.uniform(-0.5, 1)))
            kw = dict(sampling_mode=str(rng.choice(["midpoint", "simpsons", "quad:3", "trapezoid"])),
                      integrate_mode=str(rng.choice(["threshold", "threshold", "none"])),
                      sampling_tolerance=float(rng.choice([1e-2, 1e-3])), integrate_max_depth=int(rng.choice([2, 3])),
                      psf_mode=str(rng.choice(["full", "full", "none"])), psf_subpixel_shift=str(rng.choice(["bilinear", "none"])),
                      psf_convolve_mode=str(rng.choice(["fft", "direct"])))
            specs.append((f"{fam} galaxy model", win, pars, kw))
        pts = []
        for s in range(int(rng.integers(1, 3))):
            cx, cy = rng.uniform(2, W - 2), rng.uniform(2, H - 2)
            c = S @ np.array([cx, cy])
            pts.append(([[max(0, int(cx) - 6), min(W, int(cx) + 7)], [max(0, int(cy) - 6), min(H, int(cy) + 7)]],
                        {"center": [float(c[0]), float(c[1])], "flux": float(rng.uniform(0.5, 2))},
                        str(rng.choice(["bilinear", "lanczos:3"]))))

        def build(ap, k=k, up=up, S=S, H=H, W=W, pw=pw, psf_model_points=psf_model_points, specs=specs, pts=pts):
            psf = ap.image.PSF_Image(data=_psf_moffat(2.5, 0.9 * up + 0.05 * pw, pw), pixelscale=S / up)
            tar = ap.image.Target_Image(data=np.zeros((H, W)), pixelscale=S, zeropoint=22.5, psf=psf)
            M = ap.models.AstroPhot_Model
            models = [M(name=f"u{k}m{i}", model_type=mt, target=tar, window=win, parameters=dict(pars), **kw)
                      for i, (mt, win, pars, kw) in enumerate(specs)]
            pm = None
            if psf_model_points:
                ptar = ap.image.PSF_Image(data=np.zeros((6 * up + 1, 6 * up + 1)), pixelscale=S / up)
                pm = M(name=f"u{k}pm", model_type="moffat psf model", target=ptar, normalize_psf=False,
                       parameters={"n": 2.4, "Rd": 1.6 * float(np.sqrt(abs(np.linalg.det(S)))), "I0": -0.5})
            for i, (win, pars, shift) in enumerate(pts):
                if pm is not None and i == 0:
                    models.append(M(name=f"u{k}p{i}", model_type="point model", target=tar, window=win, parameters=dict(pars), psf=pm))
                else:
                    models.append(M(name=f"u{k}p{i}", model_type="point model", target=tar, window=win, parameters=dict(pars),
                                    psf_subpixel_shift=shift))
            sky = M(name=f"u{k}sky", model_type="flat sky model", target=tar, parameters={"F": -1.5})
            sky.initialize()
            return M(name=f"u{k}", model_type="group model", models=models + [sky], target=tar, psf_mode="full")

        desc = (f"{W}x{H} up={up} sheared={int(sheared)} psf={pw} models={len(specs)}+{len(pts)} "
                f"psf_model_point={int(psf_model_points)}")
        out.append((desc, up, build))
    return out


# ---------------------------------------------------------------------------
# model.initialize(): models built WITHOUT parameter values on noisy data (oracle/make_init_golden.py)
# ---------------------------------------------------------------------------
INIT_SCENES = ["init_sersic", "init_sersic_fixed_shape", "init_exponential", "init_gaussian", "init_moffat", "init_spline",
               "init_point", "init_flat_sky", "init_plane_sky", "init_moffat_psf", "init_gaussian_psf", "init_masked_sheared",
               "init_group"]


def init_data(name, load_golden):
    """Noisy image for an initialisation scene: the reference's own model image of a golden scene + seeded noise."""
    src = {"init_sersic": "c1_sersic", "init_sersic_fixed_shape": "c1_sersic", "init_exponential": "exponential",
           "init_gaussian": "gaussian", "init_moffat": "moffat", "init_spline": "spline", "init_point": "point",
           "init_flat_sky": "group", "init_plane_sky": "plane_sky_group", "init_moffat_psf": "moffat_psf_model",
           "init_gaussian_psf": "gaussian_psf_model", "init_masked_sheared": "sersic_sheared", "init_group": "group_nosky"}[name]
    truth = load_golden(src)["img0"]
    scale = 0.02 if "psf" in name else 1.0
    if name in ("init_flat_sky", "init_group"):
        truth = truth + 0.3          # a sky level to find
    return make_data([truth], 700 + INIT_SCENES.index(name), scale=scale)[0]["data"]


def build_init(ap, name, dat):
    """Model of an initialisation scene with its parameters left open."""
    M = ap.models.AstroPhot_Model
    if name in ("init_sersic", "init_sersic_fixed_shape"):
        tar = ap.image.Target_Image(data=dat, pixelscale=1.0, zeropoint=22.5)
        pars = {} if name == "init_sersic" else {"center": [50.3, 49.6], "q": 0.6, "PA": 1.0}
        return M(name=name, model_type="sersic galaxy model", target=tar, parameters=pars)
    if name in ("init_exponential", "init_gaussian", "init_moffat", "init_spline"):
        tar = ap.image.Target_Image(data=dat, pixelscale=0.9, zeropoint=22.5)
        return M(name=name, model_type=f"{name[5:]} galaxy model", target=tar)
    if name == "init_point":
        psf = ap.image.PSF_Image(data=_psf_moffat(2.5, 2.0, 15), pixelscale=1.0)
        tar = ap.image.Target_Image(data=dat, pixelscale=1.0, zeropoint=22.5, psf=psf)
        return M(name=name, model_type="point model", target=tar)
    if name in ("init_flat_sky", "init_plane_sky"):
        tar = ap.image.Target_Image(data=dat, pixelscale=1.0, zeropoint=22.5)
        return M(name=name, model_type=f"{name[5:-4]} sky model", target=tar, window=[[5, 80], [10, 90]])
    if name in ("init_moffat_psf", "init_gaussian_psf"):
        ptar = ap.image.PSF_Image(data=dat, pixelscale=1.0)
        return M(name=name, model_type=f"{name[5:-4]} psf model", target=ptar)
    if name == "init_masked_sheared":
        S = np.array([[0.8, 0.1], [-0.05, 0.9]])
        mask = np.zeros(dat.shape, dtype=bool)
        mask[10:18, 50:60] = True
        tar = ap.image.Target_Image(data=dat, pixelscale=S, origin=[3.0, -2.0], zeropoint=22.5, mask=mask)
        return M(name=name, model_type="sersic galaxy model", target=tar, window=[[8, 72], [4, 68]])
    if name == "init_group":
        tar = ap.image.Target_Image(data=dat, pixelscale=1.0, zeropoint=22.5)
        sky = M(name="ig_sky", model_type="flat sky model", target=tar)
        gals = [M(name=f"ig_gal{k}", model_type="sersic galaxy model", target=tar,
                  window=[[int(cx) - 16, int(cx) + 16], [int(cy) - 16, int(cy) + 16]])
                for k, (cx, cy) in enumerate([(28.3, 30.6), (60.2, 40.1), (45.7, 70.4)])]
        return M(name=name, model_type="group model", models=[sky] + gals, target=tar)
    raise KeyError(name)


def segmentation_inputs(load_golden):
    """(segmentation map, image) for the ap.utils.initialize tests: the crowded golden scene + seeded noise, segments =
    connected regions above a threshold (scipy.ndimage.label), ids shuffled so that they are not 1..K in scan order."""
    from scipy import ndimage
    img = make_data([load_golden("crowded")["img0"]], 321)[0]["data"]
    lab, n = ndimage.label(ndimage.gaussian_filter(img, 1.5) > 0.25)
    perm = np.random.default_rng(5).permutation(n) + 3
    seg = np.where(lab > 0, perm[np.maximum(lab, 1) - 1], 0)
    return seg, img
