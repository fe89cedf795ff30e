"""CPU-side checks: the C-ABI library loads and exports every symbol the header
declares (no compute without a GPU), host classes behave like the reference's
(window arithmetic, parameter representation round trips), and the product
path refuses to run without the native library / a CUDA device."""
import os
import re

import numpy as np
import pytest
import torch

import astrophot_b200 as ap
from astrophot_b200 import cabi
from conftest import ROOT


def test_library_exports_every_declared_symbol():
    hdr = open(os.path.join(ROOT, "include", "astrophot_b200.h")).read()
    declared = set(re.findall(r"\b(apb_[a-z0-9_]+)\s*\(", hdr))
    declared -= {"apb_plan"}
    assert declared == set(cabi.EXPORTS)
    L = cabi.load_library()
    for name in declared:
        assert hasattr(L, name), name
    assert L.apb_version() >= 100


def test_struct_sizes_match_header():
    import ctypes as C
    assert C.sizeof(cabi.apb_param_t) == 24
    assert C.sizeof(cabi.apb_image_t) == 8 + 8 * 8 + 3 * 8 + 8
    assert C.sizeof(cabi.apb_psf_t) == 24
    assert C.sizeof(cabi.apb_source_t) == 4 * 40 + 8 * 24 + 8 + 8 * 20 + 4 * 12 + 16 + 8 + 16
    assert C.sizeof(cabi.apb_owner_t) == 4 * (6 + 24)
    assert C.sizeof(cabi.apb_opts_t) == 24
    assert C.sizeof(cabi.apb_stats_t) == 8 * (1 + 5 + 2 + 2 + 2 + 10)


@pytest.mark.skipif(torch.cuda.is_available(), reason="checks the no-GPU failure mode")
def test_no_cpu_fallback():
    tar = ap.image.Target_Image(data=np.zeros((16, 16)), pixelscale=1.0)
    m = ap.models.AstroPhot_Model(name="nofb", model_type="sersic galaxy model", target=tar,
                                  parameters={"center": [8, 8], "q": 0.5, "PA": 0.1, "n": 1, "Re": 2, "Ie": 0})
    with pytest.raises(ap.errors.NativeLibraryError):
        m()


def test_missing_library_is_loud(tmp_path):
    with pytest.raises(ap.errors.NativeLibraryError):
        cabi.load_library(str(tmp_path / "nope.so"))


def test_window_arithmetic():
    W = ap.image.Window
    a = W(origin=(0, 0), pixel_shape=(100, 80), pixelscale=0.5)
    b = W(origin=(10, 5), pixel_shape=(100, 100), pixelscale=0.5)
    u, i = a | b, a & b
    np.testing.assert_allclose(u.origin.numpy(), [0, 0])
    np.testing.assert_allclose(u.pixel_shape.numpy(), [120, 110])
    np.testing.assert_allclose(i.origin.numpy(), [10, 5])
    np.testing.assert_allclose(i.pixel_shape.numpy(), [80, 70])
    rows, cols = a.get_self_indices(i)
    assert (rows.start, rows.stop, cols.start, cols.stop) == (10, 80, 20, 100)
    np.testing.assert_allclose(a.pixel_to_plane(torch.tensor([0.0, 0.0])).numpy(), [0.25, 0.25])
    np.testing.assert_allclose(a.plane_to_pixel(a.pixel_to_plane(torch.tensor([3.0, 7.0]))).numpy(), [3.0, 7.0])
    c = a.copy().pad_pixel((3,))
    np.testing.assert_allclose(c.pixel_shape.numpy(), [106, 86])
    np.testing.assert_allclose(c.origin.numpy(), [-1.5, -1.5])
    assert a.copy() == a and (a != b)


def test_parameter_representation_round_trip():
    P = ap.param.Parameter_Node
    for kw, val in [({"limits": (0, 1)}, 0.6), ({"limits": (0, None)}, 10.0), ({"limits": (None, 4.0)}, -3.0),
                    ({"limits": (0, np.pi), "cyclic": True}, 1.0), ({}, 5.0)]:
        p = P("p", value=val, **kw)
        r = p.vector_representation()
        np.testing.assert_allclose(p.vector_transform_rep_to_val(r).numpy(), [val], rtol=1e-13)
    with pytest.raises(ap.errors.InvalidParameter):
        P("q", value=2.0, limits=(0, 1))
    c = P("c", value=4.0, limits=(0, np.pi), cyclic=True)
    np.testing.assert_allclose(c.value.numpy(), 4.0 - np.pi)


def test_linked_parameters_share_a_slot():
    import scenes
    from astrophot_b200.lowering import lower

    model, _ = scenes.build(ap, "joint")
    assert len(model.parameters.vector_values()) == 9
    scene, info = lower(model)
    assert scene.n_par == 9 and len(scene.images) == 3
    for e in range(6):   # cx cy q PA n Re shared by the three bands
        assert len({s.slot[e] for s in scene.sources}) == 1
    assert len({s.slot[6] for s in scene.sources}) == 3


def test_group_overrides_psf_mode_like_reference():
    import scenes

    model, _ = scenes.build(ap, "group")
    assert all(m.psf_mode in ("full", "none") for m in model.models.values())
    assert model.models["gal0"].psf_mode == "full" and model.models["sky"].psf_mode == "none"
    tar = model.target
    g2 = ap.models.AstroPhot_Model(name="g2", model_type="group model", target=tar,
                                   models=[ap.models.AstroPhot_Model(name="x1", model_type="sersic galaxy model",
                                                                     target=tar, psf_mode="full",
                                                                     parameters={"center": [8, 8], "q": 0.5, "PA": 0.1,
                                                                                 "n": 1, "Re": 2, "Ie": 0})])
    assert g2.models["x1"].psf_mode == "none"   # group default wins (group_model_object.py:85-89)


def test_fit_mask_of_group():
    import scenes

    model, _ = scenes.build(ap, "group_nosky")
    fm = model.fit_mask()
    assert fm.shape == (58, 70) and 0 < int(fm.sum()) < fm.numel()


def test_jacobian_chunks_of_large_windows():
    """A window larger than image_chunksize is cut into the reference's Jacobian chunks (_model_methods.py:349-395):
    one piece per chunk with the chunk as output and Jacobian working window, the forward working window untouched."""
    from astrophot_b200.lowering import lower
    ap.AP_config.ap_device = "cpu"
    tar = ap.image.Target_Image(data=np.zeros((1300, 2100)), pixelscale=1.0)
    m = ap.models.AstroPhot_Model(name="big", model_type="sersic galaxy model", target=tar,
                                  parameters={"center": [1000.0, 600.0], "q": 0.6, "PA": 1.0, "n": 2.0, "Re": 50.0, "Ie": 1.0})
    scene, info = lower(m)
    assert info.chunked and len(scene.sources) == 3 * 2           # ceil(2100/1000) x ceil(1300/1000) chunks of 700 x 650
    assert sorted(s.out for s in scene.sources) == sorted((700 * a, 650 * b, 700, 650) for a in range(3) for b in range(2))
    assert all(s.jac == s.out and s.fwd == (0, 0, 2100, 1300) for s in scene.sources)
    assert len({tuple(s.slot) for s in scene.sources}) == 1       # the pieces share the model's parameters
    whole, info2 = lower(m, chunk_jacobian=False)
    assert not info2.chunked and len(whole.sources) == 1 and whole.sources[0].out == (0, 0, 2100, 1300)
    m.image_chunksize = 4000                                      # a user knob of the reference (model_object.py:109)
    assert len(lower(m)[0].sources) == 1
    # sky and point models have no integration threshold: never cut
    sky = ap.models.AstroPhot_Model(name="sk", model_type="flat sky model", target=tar, parameters={"F": -1.0})
    sky.initialize()
    assert len(lower(sky)[0].sources) == 1


def test_auxiliary_psf_model_survives_tiling_and_sharding():
    from astrophot_b200.lowering import lower, shard_scene, tile_scene
    import scenes
    ap.AP_config.ap_device = "cpu"
    model, _ = scenes.build(ap, "aux_psf_moffat")
    scene, _ = lower(model)
    assert [im.aux for im in scene.images] == [False, True] and scene.psfs[0].source == 0
    cut = tile_scene(scene, 2, 2)
    assert [im.aux for im in cut.images] == [False] * 4 + [True] and cut.owners is None
    for rank in range(2):
        part = shard_scene(cut, rank, 2)
        assert [im.aux for im in part.images] == [False, False, True]
        ps = part.psfs[0]
        assert part.sources[ps.source].image == 2 and all(s.psf == 0 for s in part.sources if s.image < 2)


def test_point_source_from_a_psf_model_lowers_to_an_amplitude_source(monkeypatch):
    """point_source.py:122-140: the PSF model's profile at the point source's centre times 10^flux (FLAG_AMP)."""
    from astrophot_b200 import scene as sc
    from astrophot_b200.lowering import lower, tile_scene
    import scenes
    ap.AP_config.ap_device = "cpu"
    model, _ = scenes.build(ap, "point_psf_model_group")
    scene, _ = lower(model)
    assert not scene.psfs and all(s.psf < 0 for s in scene.sources)         # no stamp, no shift, no convolution
    a, b, c = scene.sources[:3]
    for s in (a, b):
        assert s.kind == sc.KIND_MOFFAT and s.flags == sc.FLAG_RADIAL | sc.FLAG_NORMALIZE | sc.FLAG_AMP
        assert s.n_elem == len(sc.ELEMS[sc.KIND_MOFFAT]) + 1 and s.slot[-1] >= 0          # flux: the last element
        assert s.sampling_mode == sc.SAMPLE_SIMPSONS and s.tolerance == 1e-3                # the PSF model's knobs
        assert s.out == s.fwd == s.jac
    assert a.slot[4:6] == b.slot[4:6] and min(a.slot[4:6]) >= 0                          # one shared PSF model
    assert a.slot[:2] != b.slot[:2] and a.slot[-1] != b.slot[-1]                         # own centre and flux
    assert a.slot[2] == a.slot[3] == -1 and (a.cval[2], a.cval[3]) == (1.0, 0.0)         # radial: q = 1, PA = 0
    assert c.kind == sc.KIND_GAUSSIAN and c.flags == sc.FLAG_RADIAL | sc.FLAG_AMP and c.fwd != c.out
    # shared PSF parameters: dense solve, no owner layout; a normalised piece may not be cut
    with pytest.raises(ap.errors.SpecificationConflict, match="tile"):
        tile_scene(scene, 2, 2)
    # a normalised PSF model is normalised over the window it is sampled on: inside a group that is the group's
    tar = model.target
    pm = model.models["starA"].psf
    small = ap.models.AstroPhot_Model(name="starS", model_type="point model", target=tar, psf=pm,
                                      window=[[10, 30], [12, 32]], parameters={"center": [18.4, 20.7], "flux": 2.0})
    sky = ap.models.AstroPhot_Model(name="skyS", model_type="flat sky model", target=tar, parameters={"F": -1.0})
    sky.initialize()
    g = ap.models.AstroPhot_Model(name="gS", model_type="group model", models=[small, sky], target=tar)
    with pytest.raises(ap.errors.SpecificationConflict, match="normalised PSF model"):
        lower(g)
    assert lower(small)[0].sources[0].out == (0, 0, 20, 20)          # on its own (scene image = its window) it is fine


def test_iter_lm_visits_the_chunks_like_the_reference():
    """Iter_LM._sweep against the reference's selection rules (fit/iterative.py:225-275), restated here: integer
    chunks deal the identities out from the front / by random.sample of the remaining ones, explicit chunks go in
    order / by random.choice without replacement -- same RNG calls, so the same sweep under the same seed."""
    import random
    import scenes
    from astrophot_b200.fit import Iter_LM
    ap.AP_config.ap_device = "cpu"
    model, _ = scenes.build(ap, "group_nosky")
    ids = list(model.parameters.vector_identities())
    assert len(ids) == 13

    def expected(chunks, method, seed):
        random.seed(seed)
        out, left = [], list(ids)
        if isinstance(chunks, int):
            while left:
                take = random.sample(left, min(len(left), chunks)) if method == "random" else left[:chunks]
                out.append(sorted(ids.index(p) for p in take))
                left = [p for p in left if p not in take]
        else:
            choices = list(range(len(chunks)))
            while choices:
                k = random.choice(choices) if method == "random" else choices[0]
                choices.remove(k)
                out.append(sorted(ids.index(p) for p in chunks[k]))
        return out

    explicit = (tuple(ids[0:5]), tuple(ids[5:6]), tuple(ids[6:13]))
    for chunks in (5, 13, 50, explicit):
        for method in ("sequential", "random"):
            opt = Iter_LM(model, initial_state=np.zeros(13), chunks=chunks, method=method)
            for seed in (0, 7):
                random.seed(seed)
                got = [sorted(np.flatnonzero(m.numpy()).tolist()) for m in opt._sweep(ids)]
                assert got == expected(chunks, method, seed), (chunks, method, seed)
                assert sorted(sum(got, [])) == list(range(13))      # every parameter exactly once per sweep


def test_iter_lm_control_flow_on_the_oracle(monkeypatch):
    """Iter_LM end to end without a GPU: the per-chunk LM is replaced by the CPU oracle's LM on the scene lowered
    under the chunk's Param_Mask; the sweep history must reproduce the reference's (golden `iterlm_*`)."""
    import astrophot_oracle as orc
    import scenes
    from astrophot_b200 import fit as fitmod
    from astrophot_b200.lowering import lower
    from conftest import load_golden, golden_data

    class OracleLM:
        def __init__(self, model, ndf=None, max_iter=100, relative_tolerance=1e-5, **kw):
            self.model, self.ndf, self.max_iter, self.rtol = model, ndf, max_iter, relative_tolerance

        def fit(self):
            scene, _ = lower(self.model, for_fit=True)
            x0 = self.model.parameters.vector_representation().numpy()
            self.r = orc.lm_fit(scene, x0, max_iter=self.max_iter, relative_tolerance=self.rtol, ndf=self.ndf)
            loss = np.array(self.r["loss_history"])
            ok = np.isfinite(loss)
            best = np.array(self.r["lambda_history"])[ok][np.argmin(loss[ok])]
            self.model.parameters.vector_set_representation(torch.as_tensor(best))
            return self

        def res_loss(self):
            loss = np.array(self.r["loss_history"])
            return float(np.min(loss[np.isfinite(loss)]))

    ap.AP_config.ap_device = "cpu"
    monkeypatch.setattr(fitmod, "LM", OracleLM)
    for name in scenes.ITER_SCENES:
        fix = load_golden(name)
        model, _ = scenes.build(ap, name, data=golden_data(fix))
        model.parameters.vector_set_representation(torch.as_tensor(fix["x0"]))
        res = fitmod.Iter_LM(model, initial_state=fix["x0"], chunks=8, method="sequential", max_iter=2,
                             LM_kwargs={"max_iter": 3, "relative_tolerance": 0.0}).fit()
        np.testing.assert_allclose(res.loss_history, fix["iterlm_loss_history"], rtol=1e-8)
        np.testing.assert_allclose(np.array(res.lambda_history), fix["iterlm_lambda_history"], rtol=1e-7, atol=1e-7)


def test_parameter_report_and_state_round_trip():
    """The text a user reads after a fit (``print(model.parameters)``) and the plain-dict state of the parameter graph
    (reference: `parameter.py:598-647,677-742`, `param/base.py:134-161`)."""
    from astrophot_b200.param import Parameter_Node
    a = Parameter_Node("a", value=[1.0, 2.0], uncertainty=[0.1, 0.2], limits=(0, None), units="arcsec")
    b = Parameter_Node("b", value=3.0, locked=True, prof=[0.0, 1.0], cyclic=True, limits=(0, 3.5))
    ptr = Parameter_Node("ptr", value=a)
    g = Parameter_Node("g", link=(a, b, ptr))
    assert str(a) == "a: [1.0, 2.0] +- [0.1, 0.2] [arcsec], limits: (0.0, None)"
    assert str(b) == "b: 3.0 [none], limits: (0.0, 3.5), cyclic, locked"
    assert str(g) == "g:\n" + str(a) + "\n" + str(b)            # every leaf once, pointers are not repeated
    lines = repr(g).split("\n")
    assert lines[0].startswith("g (id-") and lines[0].endswith(", branch node):")
    assert lines[1].startswith("  a (id-") and " points to: a (id-" in lines[3]
    assert b.print_params(include_id=False).endswith("prof: [0.0, 1.0]")
    g2 = Parameter_Node("other", state=Parameter_Node("g", link=(a, b)).get_state())
    assert g2.name == "g" and list(g2.nodes) == ["a", "b"]
    assert str(g2["a"]) == str(a) and str(g2["b"]) == str(b) and g2["b"].locked and g2["a"].identity == a.identity
    np.testing.assert_array_equal(g2["b"].prof.numpy(), [0.0, 1.0])
    assert g.get_state()["nodes"][2]["value"] == f"NODE:{a.identity}"


@pytest.mark.parametrize("name,ext", [("psf_sersic", ".yaml"), ("group", ".yaml"), ("aux_psf_moffat", ".json"), ("joint", ".yaml"),
                                      ("crowded", ".json"), ("spline", ".yaml"), ("masked_locked_edge", ".yaml"),
                                      ("point_psf_model_group", ".yaml"), ("moffat_psf_model", ".yaml")])
def test_model_save_load_round_trip(tmp_path, name, ext):
    """model.save() / AstroPhot_Model(filename=...) (reference: `core_model.py:374-451`, `_model_methods.py:437-482`,
    `group_model_object.py:325-365`): the reloaded model lowers to the very same scene tables -- windows, knobs, PSFs,
    parameter order, limits, locks -- and parameters shared between sub-models (joint fits, one PSF model for several
    stars) are still shared afterwards."""
    import scenes
    from astrophot_b200.lowering import lower
    ap.AP_config.ap_device = "cpu"
    model, _ = scenes.build(ap, name)
    path = str(tmp_path / f"saved{ext}")
    model.save(path)
    again = ap.models.AstroPhot_Model(name=f"again {name}", filename=path, target=model.target)
    assert type(again) is type(model) and again.name == f"again {name}"
    a, b = lower(model)[0], lower(again)[0]
    assert len(a.sources) == len(b.sources) and len(a.psfs) == len(b.psfs)
    for x, y in zip(a.sources, b.sources):
        assert {k: v for k, v in x.__dict__.items() if k != "name"} == {k: v for k, v in y.__dict__.items() if k != "name"}
    np.testing.assert_array_equal(a.transform, b.transform)
    np.testing.assert_array_equal(a.lo, b.lo)
    np.testing.assert_array_equal(a.hi, b.hi)
    moved = model.parameters.vector_values() * 1.01
    model.parameters.vector_set_values(moved)
    again.parameters.vector_set_values(moved)
    for x, y in zip(lower(model)[0].sources, lower(again)[0].sources):
        assert x.cval == y.cval
    with pytest.raises(ValueError):
        model.save(str(tmp_path / "saved.hdf5"))


def test_parameter_order_of_every_model_family():
    """The column order of the Jacobian is the models' ``parameter_order`` (reference: each model's ``_parameter_order``;
    the moffat2d psf model appends q, PA after the Moffat parameters, `moffat_model.py:134`)."""
    want = {
        "sersic galaxy model": ("center", "q", "PA", "n", "Re", "Ie"),
        "exponential galaxy model": ("center", "q", "PA", "Re", "Ie"),
        "gaussian galaxy model": ("center", "q", "PA", "sigma", "flux"),
        "moffat galaxy model": ("center", "q", "PA", "n", "Rd", "I0"),
        "spline galaxy model": ("center", "q", "PA", "I(R)"),
        "point model": ("center", "flux"),
        "flat sky model": ("center", "F"),
        "plane sky model": ("center", "F", "delta"),
        "sersic psf model": ("center", "n", "Re", "Ie"),
        "exponential psf model": ("center", "Re", "Ie"),
        "gaussian psf model": ("center", "sigma", "flux"),
        "moffat psf model": ("center", "n", "Rd", "I0"),
        "moffat2d psf model": ("center", "n", "Rd", "I0", "q", "PA"),
        "spline psf model": ("center", "I(R)"),
    }
    have = {m.model_type: tuple(m._parameter_order) for m in ap.models.AstroPhot_Model.List_Models(usable=True)
            if m.model_type != "group model"}
    assert have == want


def test_supersampled_psf_lowering():
    """psf_upscale (model_object.py:312-315, point_source.py:123-127,147-149): lowered as `upscale` on the sources that have
    a PSF, windows left in image pixels; the oracle's fine-grid sampling conserves what the image-pixel sampling gives for
    a smooth source."""
    import scenes
    from astrophot_b200 import scene as sc
    from astrophot_b200.lowering import lower, tile_scene
    ap.AP_config.ap_device = "cpu"
    for desc, up, build in scenes.upscale_fuzz_builders(4):
        scene, _ = lower(build(ap))
        for s in scene.sources:
            has_psf = s.psf >= 0 or bool(s.flags & sc.FLAG_AMP)
            assert s.upscale == (up if has_psf else 1), (desc, s.name)
            im = scene.images[s.image]
            assert s.out[0] >= 0 and s.out[0] + s.out[2] <= im.W and s.out[1] + s.out[3] <= im.H
        cut = tile_scene(scene, 2, 2)
        assert {s.upscale for s in cut.sources} == {s.upscale for s in scene.sources}
    model, _ = scenes.build(ap, "psf_sersic_up2")
    assert [s.upscale for s in lower(model)[0].sources] == [2]
    tar = ap.image.Target_Image(data=np.zeros((20, 20)), pixelscale=0.5,
                                psf=ap.image.PSF_Image(data=np.ones((5, 5)), pixelscale=1.0))
    m = ap.models.AstroPhot_Model(name="coarse", model_type="point model", target=tar, parameters={"center": [5, 5], "flux": 1})
    with pytest.raises(ap.errors.SpecificationConflict):
        lower(m)


def test_balanced_tile_cuts_partition_the_image():
    """lowering._balanced_cuts / tile_scene(balance=True): whatever the cost profile, the cuts are strictly increasing, start
    at 0 and end at the image edge, and the tiles own every pixel once; an empty cost profile gives the even cuts."""
    from astrophot_b200.lowering import _balanced_cuts, lower, tile_scene
    import scenes
    rng = np.random.default_rng(8)
    for trial in range(200):
        L = int(rng.integers(4, 300))
        n = int(rng.integers(1, min(L, 9) + 1))
        kind = trial % 4
        cost = (np.zeros(L), rng.uniform(0, 1, L), (rng.uniform(0, 1, L) > 0.97) * rng.uniform(0, 100, L),
                np.concatenate([np.zeros(L - 1), [5.0]]))[kind]
        cuts = _balanced_cuts(cost, n)
        assert cuts[0] == 0 and cuts[-1] == L and len(cuts) == n + 1
        assert all(b > a for a, b in zip(cuts, cuts[1:])), (trial, cuts)
        if kind == 0:
            assert cuts == [round(k * L / n) for k in range(n + 1)]
        elif kind == 1 and L >= 20 * n:       # a smooth profile: every run within one element of the even share of the cost
            tot = cost.sum()
            for a, b in zip(cuts, cuts[1:]):
                assert abs(cost[a:b].sum() - tot / n) <= 2 * cost.max() + 1e-12
    ap.AP_config.ap_device = "cpu"
    model, _ = scenes.build(ap, "crowded")
    scene, _ = lower(model)
    im0 = scene.images[0]
    for ny, nx in ((1, 2), (2, 2), (2, 4), (3, 5)):
        cut = tile_scene(scene, ny, nx)
        seen = np.zeros((im0.H, im0.W), dtype=int)
        for im in cut.images:
            x0, y0 = (np.round(np.asarray(im0.rij) - np.asarray(im.rij))).astype(int)
            seen[y0:y0 + im.H, x0:x0 + im.W] += 1
        assert len(cut.images) == ny * nx and np.all(seen == 1)
        # every piece of a source sits inside its tile
        for s in cut.sources:
            im = cut.images[s.image]
            assert s.out[0] >= 0 and s.out[1] >= 0 and s.out[0] + s.out[2] <= im.W and s.out[1] + s.out[3] <= im.H
