"""Lower a model tree to the flat tables of the C ABI.

The reference evaluates a group by walking Python objects, one torch op at a
time, and scatters per-model Jacobians into a dense (H, W, P) tensor
(`group_model_object.py:183-283`).  Here the walk happens once: the tree
becomes a source table (kind, windows in image pixels, parameter slots,
sampling knobs), an image table (geometry + data/weight/mask pointers), a PSF
table and a parameter table (transform + limits).  Everything the kernels need
about windows — the quirk that a sub-model inside a group is *sampled* over
the group's window but *differentiated* over its own
(`group_model_object.py:211-227` vs `_model_methods.py:294-299`) — is encoded
as three rectangles per source: ``out``, ``fwd`` and ``jac``.
"""
import numpy as np
import torch

from . import AP_config
from . import scene as sc
from .errors import SpecificationConflict, InvalidParameter
from .image import (Image_List, Jacobian_Image, Jacobian_Image_List, Model_Image, Model_Image_List,
                    PSF_Image, Window, Window_List)
from .param import Parameter_Node

__all__ = ["lower", "shard_scene", "tile_scene", "wrap_model_images", "wrap_jacobian_images", "LoweringInfo"]


class LoweringInfo:
    """Host-side bookkeeping that goes with a Scene."""

    def __init__(self):
        self.windows = []       # Window of each scene image (the evaluated region)
        self.targets = []       # Target image (full) of each scene image
        self.is_list = False
        self.identities = None
        self.components = []    # component model per source
        self.chunked = False    # some model was cut into Jacobian chunks (window > image_chunksize)


def _resolve_leaf(node):
    """Follow pointer nodes to the leaf that owns the tensor."""
    seen = 0
    while isinstance(node._value, Parameter_Node):
        node = node._value
        seen += 1
        if seen > 64:
            raise InvalidParameter("pointer chain too long")
    if callable(node._value) and not isinstance(node._value, torch.Tensor):
        raise SpecificationConflict(
            f"parameter '{node.name}' is function-valued; only tensor and pointer values can be lowered to the device")
    return node


def _components(model):
    from .models import Group_Model

    if isinstance(model, Group_Model):
        out = []
        for sub in model.models.values():
            out.extend(_components(sub))
        return out
    return [model]


def _rect(region: Window, win: Window):
    """Pixel rectangle (x0, y0, w, h) of ``win`` inside ``region``."""
    rows, cols = region.get_self_indices(win)
    return (cols.start, rows.start, max(cols.stop - cols.start, 0), max(rows.stop - rows.start, 0))


def _rect_unclipped(region: Window, win: Window):
    """Pixel rectangle of ``win`` in ``region``'s pixel grid WITHOUT clipping to it (may start below 0 / reach past
    the region).  The reference hands an explicitly requested window to the sampling code as it is
    (`model_object.py:296-299`: ``working_window = window.copy()``) while the output image is the target cut to it:
    a model window that sticks out of the target is sampled on all of it."""
    lo = np.round(region._plane2pix(win._origin)[:, 0] + 0.5).astype(np.int64)
    hi = np.round(region._plane2pix(win._origin + win._end)[:, 0] + 0.5).astype(np.int64)
    return (int(lo[0]), int(lo[1]), int(max(hi[0] - lo[0], 0)), int(max(hi[1] - lo[1], 0)))


_SAMPLING = {"midpoint": sc.SAMPLE_MIDPOINT, "simpsons": sc.SAMPLE_SIMPSONS, "trapezoid": sc.SAMPLE_TRAPEZOID}


def _shift_code(name):
    if name == "none":
        return sc.SHIFT_NONE
    if name == "bilinear":
        return sc.SHIFT_BILINEAR
    if isinstance(name, str) and name.startswith("lanczos"):
        # "lanczos:k" (_model_methods.py:209-227): code 10 + k
        try:
            order = int(name[name.find(":") + 1:])
        except ValueError:
            raise SpecificationConflict(f"unrecognized subpixel shift method: {name}")
        if not 1 <= order <= 8:
            raise SpecificationConflict(f"lanczos order out of range (1..8): {name}")
        return sc.SHIFT_LANCZOS + order
    raise SpecificationConflict(f"unrecognized subpixel shift method: {name}")


def lower(model, window=None, for_fit=False, chunk_jacobian=True):
    """Build (Scene, LoweringInfo) for ``model`` evaluated on ``window``
    (default: the model's own window).  ``for_fit`` ORs the model's
    ``fit_mask()`` into each image mask the way ``LM.__init__`` does
    (`fit/lm.py:204-222`).  ``chunk_jacobian=False`` leaves models whose window exceeds ``image_chunksize`` in one
    piece: right for forward-only plans (the reference never chunks the forward model), see below."""
    from .models import AstroPhot_Model, Group_Model, PSF_Model, Point_Source, Component_Model

    info = LoweringInfo()
    target = model.target
    info.is_list = isinstance(target, Image_List)
    is_group = isinstance(model, Group_Model)

    # ---- regions (one scene image per target image)
    mwin = model.window
    if window is not None:
        mwin = mwin & window
    if info.is_list:
        targets = list(target.image_list)
        regions = list(mwin.window_list)
    else:
        targets = [target]
        regions = [mwin]
    asked = [r.copy() for r in regions]     # the requested windows as they are (may stick out of their targets)
    explicit = window is not None
    images = []
    fmask = None
    if for_fit:
        fmask = model.fit_mask()
        fmask = list(fmask) if isinstance(fmask, (tuple, list)) else [fmask]
        if sum(int(torch.sum(f)) for f in fmask) == 0:
            fmask = None
    for k, (tar, reg) in enumerate(zip(targets, regions)):
        sub = tar[reg]
        w = sub.window
        mask = sub.mask if getattr(sub, "has_mask", False) else None
        if fmask is not None:
            fm = fmask[k]
            if window is not None:   # fit_mask() is on the model window; cut to the requested region
                rows, cols = model.window.window_list[k].get_self_indices(w) if info.is_list else model.window.get_self_indices(w)
                fm = fm[rows, cols]
            mask = fm if mask is None else (mask | fm)
        images.append(sc.SceneImage(
            H=int(w._shape[1]), W=int(w._shape[0]), S=w._S.copy(), rij=w._rij.copy(), rxy=w._rxy.copy(),
            data=sub.data.contiguous(),
            weight=sub.weight.contiguous() if getattr(sub, "has_weight", False) else None,
            mask=None if mask is None else mask.contiguous()))
        info.windows.append(w)
        info.targets.append(tar)
    ident_to_img = {t.identity: i for i, t in enumerate(targets)}

    # ---- parameter table
    leaves = list(model.parameters.flat(include_locked=False, include_links=False).values())
    slot_of = {}
    transform, lo, hi = [], [], []
    n = 0
    for leaf in leaves:
        m = leaf.mask.reshape(-1).numpy()
        l0, l1 = leaf.limits
        l0 = None if l0 is None else np.broadcast_to(l0.numpy(), leaf.shape).reshape(-1)
        l1 = None if l1 is None else np.broadcast_to(l1.numpy(), leaf.shape).reshape(-1)
        for e in range(leaf.size):
            if not m[e]:
                continue
            slot_of[(leaf.identity, e)] = n
            a = np.nan if l0 is None else float(l0[e])
            b = np.nan if l1 is None else float(l1[e])
            if leaf.cyclic:
                t = sc.TR_CYCLIC
            elif l0 is None and l1 is None:
                t = sc.TR_NONE
            elif l1 is None:
                t = sc.TR_LOWER
            elif l0 is None:
                t = sc.TR_UPPER
            else:
                t = sc.TR_BOTH
            transform.append(t)
            lo.append(a)
            hi.append(b)
            n += 1
    info.identities = list(model.parameters.vector_identities())

    psfs, psf_index = [], {}

    def build_source(comp, ii, region, out, fwd, jac, host=None):
        """``host``: the point source that draws ``comp`` (a PSF model) at its own centre, times 10^flux
        (point_source.py:122-140)."""
        # the model's own mask (model_object.py:370-371): zeroes the model (and, through autodiff, its Jacobian) on the
        # image it was sampled on, so it must have that image's shape -- the model's window, which inside a group must
        # then equal the window the group samples it on
        mk, mk_origin = (comp if host is None else host).mask, (0, 0)
        if host is not None and comp.mask is not None:
            raise SpecificationConflict(f"{comp.name}: a mask on a PSF model that a point source is drawn from is not supported")
        if mk is not None:
            mk = np.ascontiguousarray(torch.as_tensor(mk).detach().cpu().numpy() != 0)
            for rect in (out, fwd):
                if mk.shape == (rect[3], rect[2]):
                    mk_origin = (int(rect[0]), int(rect[1]))
                    break
            else:
                raise SpecificationConflict(
                    f"{comp.name}: the model mask has shape {mk.shape}, the image the model is sampled on {(out[3], out[2])}")
        # elements
        names = list(sc.ELEMS[comp._kind])
        nodes = []
        if isinstance(comp, PSF_Model):
            have = set(comp._parameter_order)
        for nm in names:
            if nm in ("cx", "cy"):
                nodes.append(((comp if host is None else host).parameters["center"], 0 if nm == "cx" else 1))
            elif nm in ("dx", "dy"):          # plane sky slopes: the two elements of `delta`
                nodes.append((comp.parameters["delta"], 0 if nm == "dx" else 1))
            elif isinstance(comp, PSF_Model) and nm in ("q", "PA") and nm not in have:
                nodes.append((None, 1.0 if nm == "q" else 0.0))
            else:
                nodes.append((comp.parameters[nm], 0))
        prof = []
        if comp._kind == sc.KIND_SPLINE:
            pn = _resolve_leaf(comp.parameters["I(R)"])
            if pn.prof is None or pn.value is None:
                raise InvalidParameter(f"{comp.name}: spline model needs I(R) values and prof radii")
            prof = [float(v) for v in pn.prof.reshape(-1)]
            if len(prof) > sc.MAX_PROF:
                raise SpecificationConflict(f"spline with more than {sc.MAX_PROF} nodes")
            if not getattr(comp, "extend_profile", True):
                raise SpecificationConflict("extend_profile=False is not supported by astrophot_b200")
            for k in range(len(prof)):
                nodes.append((comp.parameters["I(R)"], k))
        if host is not None:
            nodes.append((host.parameters["flux"], 0))      # FLAG_AMP: the last element
            if len(nodes) > sc.MAX_ELEM:
                raise SpecificationConflict(f"{host.name}: too many elements for one source")
        slot, cval = [], []
        for node, e in nodes:
            if node is None:
                slot.append(-1)
                cval.append(float(e))
                continue
            leaf = _resolve_leaf(node)
            if leaf.value is None:
                if leaf.name == "center" and comp._kind == sc.KIND_FLAT_SKY:   # flat sky: centre is irrelevant
                    slot.append(-1)
                    cval.append(0.0)
                    continue
                raise InvalidParameter(f"{comp.name}: parameter '{leaf.name}' has no value")
            slot.append(slot_of.get((leaf.identity, e), -1))
            cval.append(float(leaf.value.reshape(-1)[e]))

        # sampling knobs
        sm = comp.sampling_mode
        quad_init = 3
        if sm in _SAMPLING:
            smode = _SAMPLING[sm]
        elif isinstance(sm, str) and "quad" in sm:
            smode = sc.SAMPLE_QUAD
            quad_init = int(sm[sm.find(":") + 1:])
        else:
            raise SpecificationConflict(
                f"{comp.name} has unknown sampling mode: {sm}. Should be one of: midpoint, simpsons, quad:level, trapezoid")
        im = comp.integrate_mode
        if im == "none":
            imode = sc.INTEGRATE_NONE
        elif im == "threshold":
            imode = sc.INTEGRATE_THRESHOLD
        else:
            raise SpecificationConflict(
                f"{comp.name} has unknown integration mode: {im}. Should be one of: none, threshold")
        if comp._kind in (sc.KIND_FLAT_SKY, sc.KIND_POINT, sc.KIND_PLANE_SKY):
            imode = sc.INTEGRATE_NONE

        # psf
        pidx = -1
        up = 1
        pmode = comp.psf_mode
        if pmode not in ("none", "full"):
            raise SpecificationConflict(f"unknown psf_mode: {pmode}")
        if comp.psf_convolve_mode not in ("fft", "direct"):
            raise SpecificationConflict(f"unrecognized psf_convolve_mode: {comp.psf_convolve_mode}")
        if pmode == "full":
            psf = comp.psf
            if psf is None:
                raise SpecificationConflict(f"{comp.name}: psf_mode='full' but no PSF on model or target")
            pwin = psf.window if isinstance(psf, PSF_Image) else psf.target.window
            # super-sampled PSF (model_object.py:312-315,348-349; point_source.py:147-149,181): the source is sampled on
            # pixels 1 / up of the image's and block-summed back
            up = int(np.round(float(region.pixel_length) / float(pwin.pixel_length)))
            if up < 1:
                raise SpecificationConflict(f"{comp.name}: the PSF's pixels are larger than the image's")
            if up > 16:
                raise SpecificationConflict(f"{comp.name}: psf_upscale = {up} > 16")
            if id(psf) not in psf_index:
                if isinstance(psf, PSF_Image):
                    psfs.append(sc.ScenePSF(data=psf.data.contiguous()))
                else:
                    psrc, shape = aux_psf_source(psf)
                    psfs.append(sc.ScenePSF(data=None, source=psrc, shape=shape))
                psf_index[id(psf)] = len(psfs) - 1
            pidx = psf_index[id(psf)]
        if pidx >= 0 and comp._kind != sc.KIND_POINT and _shift_code(comp.psf_subpixel_shift) > sc.SHIFT_LANCZOS + 1 \
                and not (tuple(out) == tuple(fwd) == tuple(jac)):
            # the lanczos:k stamp is 2 (k - 1) pixels wider than the PSF border: in the reference's circular FFT
            # convolution its outer taps wrap around the padded WORKING image -- the group's window in a forward pass,
            # the model's own in the Jacobian.  Reproduced when the two coincide.
            raise SpecificationConflict(
                f"{comp.name}: psf_subpixel_shift='{comp.psf_subpixel_shift}' on a PSF-convolved model needs the model's "
                "window to be the window it is sampled on (stand-alone, or the group's window)")
        if comp._kind == sc.KIND_POINT and comp.psf_subpixel_shift == "none":
            # the reference shifts a point source's PSF image unconditionally (point_source.py:157-162 -> _shift_psf),
            # which has no "none" method: same error here
            raise SpecificationConflict("unrecognized subpixel shift method: none")
        flags = comp._flags
        if isinstance(comp, PSF_Model) and comp.normalize_psf:
            flags |= sc.FLAG_NORMALIZE
        if host is not None:
            flags |= sc.FLAG_AMP
            if (flags & sc.FLAG_NORMALIZE) and not (tuple(out) == tuple(fwd) == tuple(jac)):
                # the reference normalises the PSF model over the whole working window (inside a group: the GROUP
                # window in the forward pass, the point source's own window in the Jacobian, psf_model_object.py:255)
                raise SpecificationConflict(
                    f"{host.name}: a point source with a normalised PSF model must have the window it is sampled on "
                    "(inside a group: the group's) as its own window; set normalize_psf=False on the PSF model or "
                    "give the point source the group's window")
        return sc.SceneSource(
            kind=comp._kind, image=ii, out=out, fwd=fwd, jac=jac, slot=slot, cval=cval, flags=flags, prof=prof,
            sampling_mode=smode, quad_init=quad_init, integrate_mode=imode,
            quad_level=int(comp.integrate_quad_level), gridding=int(comp.integrate_gridding),
            max_depth=int(comp.integrate_max_depth), tolerance=float(comp.sampling_tolerance),
            softening=float(comp.softening), ref_mode=comp._ref_mode, psf=pidx,
            psf_shift=_shift_code(comp.psf_subpixel_shift),
            conv_mode=sc.CONV_DIRECT if comp.psf_convolve_mode == "direct" else sc.CONV_AUTO, name=comp.name,
            mask=mk, mask_origin=mk_origin, upscale=up)

    # ---- sources
    sources = []
    aux_images = []          # grids of auxiliary PSF models (PSF_Image targets), appended after the target images
    n_real = len(images)

    def aux_psf_source(pm):
        """PSF *model* used as the PSF of a host model (model_object.py:133-147): it is sampled on its own
        PSF_Image grid on every pass (model_object.py:307-310), like a stand-alone PSF model, and its free
        parameters are parameters of the host."""
        if not isinstance(pm, PSF_Model) or getattr(pm, "_kind", None) is None:
            raise SpecificationConflict(
                f"auxiliary PSF model type '{pm.model_type}' is outside the hot-path scope of astrophot_b200 (SURVEY.md §8f)")
        if pm.mask is not None:
            raise SpecificationConflict(f"{pm.name}: a mask on an auxiliary PSF model is not supported")
        w = pm.window
        ii = n_real + len(aux_images)
        aux_images.append(sc.SceneImage(H=int(w._shape[1]), W=int(w._shape[0]), S=w._S.copy(), rij=w._rij.copy(),
                                        rxy=w._rxy.copy(), aux=True))
        rect = (0, 0, int(w._shape[0]), int(w._shape[1]))
        src = build_source(pm, ii, w, rect, rect, rect)
        sources.append(src)
        info.components.append(pm)
        return len(sources) - 1, (int(w._shape[1]), int(w._shape[0]))

    def point_from_psf_model(comp, ii, region, out, fwd, jac):
        """Point source whose PSF is a PSF *model* (point_source.py:122-140): the reference samples the PSF model on
        the working window shifted by -centre -- the model's own sampling mode and integration knobs, normalised over
        that window when ``normalize_psf`` -- and multiplies by 10^flux.  No PSF stamp, no shift, no convolution: it is
        lowered as a source of the PSF model's kind on the target's grid whose centre elements are the point source's
        and whose last element is the flux (FLAG_AMP)."""
        pm = comp.psf
        if not isinstance(pm, PSF_Model) or getattr(pm, "_kind", None) is None:
            raise SpecificationConflict(
                f"PSF model type '{pm.model_type}' is outside the hot-path scope of astrophot_b200 (SURVEY.md §8f)")
        # super-sampled PSF model (point_source.py:123-127,181): sampled on pixels 1 / up of the image's, block-summed back
        up = int(np.round(float(region.pixel_length) / float(pm.target.window.pixel_length)))
        if not 1 <= up <= 16:
            raise SpecificationConflict(f"{comp.name}: psf_upscale = {up} outside 1..16")
        if getattr(pm, "model_integrated", False) is not False:
            raise SpecificationConflict("model_integrated PSF models are not supported by astrophot_b200")
        src = build_source(pm, ii, region, out, fwd, jac, host=comp)
        src.upscale = up
        return src

    for comp in _components(model):
        if isinstance(comp, Component_Model) and comp._kind is None:
            continue     # the abstract bases ("model", "galaxy model", ...) evaluate to zero (model_object.py:235-256)
        if not isinstance(comp, Component_Model):
            raise SpecificationConflict(
                f"model type '{comp.model_type}' is outside the hot-path scope of astrophot_b200 (SURVEY.md §2)")
        ii = ident_to_img.get(comp.target.identity)
        if ii is None:
            raise SpecificationConflict(f"{comp.name}: target not part of the evaluated model's target")
        region = info.windows[ii]
        cwin = comp.window
        out = _rect(region, cwin)
        if out[2] <= 0 or out[3] <= 0:
            continue
        # Working windows.  The Jacobian is always taken on (own window & requested window) as it is
        # (_model_methods.py:294-299); the forward model on the window it is asked for as it is when one is passed
        # explicitly (``model(window=...)``), else on the target cut to the model's window (model_object.py:291-296)
        # -- which is what the reference's LM does: its forward is ``partial(model, as_representation=True)``
        # without a window (fit/lm.py:173), only its Jacobian sees the uncut window.  Inside a group both sub-model
        # passes get the group's window (group_model_object.py:211-227) / its overlap with the sub-model's own
        # (group_model_object.py:258-266).
        jac = _rect_unclipped(region, cwin & asked[ii])
        if explicit:
            fwd = _rect_unclipped(region, asked[ii]) if is_group else jac
        else:
            fwd = (0, 0, images[ii].W, images[ii].H) if is_group else out
        if comp._kind == sc.KIND_POINT and isinstance(comp.psf, AstroPhot_Model):
            src = point_from_psf_model(comp, ii, region, out, fwd, jac)
        else:
            src = build_source(comp, ii, region, out, fwd, jac)
        # Windows larger than image_chunksize pixels: the reference evaluates the Jacobian chunk by chunk
        # (_model_methods.py:349-395), each chunk sampled on its OWN sub-window -- so the integration threshold
        # (total_flux / numel, mean reference) of the derivative pass is the chunk's, not the window's.  Reproduced by
        # cutting the source into one piece per chunk: output window = the chunk, Jacobian working window = the chunk,
        # forward working window unchanged (the forward model is never chunked).  The chunk grid is the reference's,
        # including its rounding (a remainder smaller than half a chunk is left uncovered by its Jacobian).
        csize = int(getattr(comp, "image_chunksize", 1000))
        threshold_kind = comp._kind not in (sc.KIND_FLAT_SKY, sc.KIND_PLANE_SKY, sc.KIND_POINT)
        if chunk_jacobian and threshold_kind and src.integrate_mode == sc.INTEGRATE_THRESHOLD and max(jac[2], jac[3]) > csize:
            info.chunked = True
            ncx, ncy = int(np.ceil(jac[2] / csize)), int(np.ceil(jac[3] / csize))
            cw, ch = int(np.round(jac[2] / ncx)), int(np.round(jac[3] / ncy))
            covered = np.zeros((out[3], out[2]), dtype=bool)
            for nx_ in range(ncx):
                for ny_ in range(ncy):
                    x0c, x1c = jac[0] + cw * nx_, jac[0] + min(jac[2], cw * (nx_ + 1))
                    y0c, y1c = jac[1] + ch * ny_, jac[1] + min(jac[3], ch * (ny_ + 1))
                    ox0, oy0 = max(out[0], x0c), max(out[1], y0c)
                    ox1, oy1 = min(out[0] + out[2], x1c), min(out[1] + out[3], y1c)
                    if ox1 <= ox0 or oy1 <= oy0:
                        continue
                    piece = sc.SceneSource(**{**src.__dict__})
                    piece.out = (ox0, oy0, ox1 - ox0, oy1 - oy0)
                    piece.jac = (x0c, y0c, x1c - x0c, y1c - y0c)
                    sources.append(piece)
                    info.components.append(comp)
                    covered[oy0 - out[1]:oy1 - out[1], ox0 - out[0]:ox1 - out[0]] = True
            if not covered.all():
                raise SpecificationConflict(
                    f"{comp.name}: window {jac[2]}x{jac[3]} does not divide into image_chunksize={csize} chunks the way the "
                    "reference rounds them (its Jacobian would skip the last pixels); choose another image_chunksize")
            continue
        sources.append(src)
        info.components.append(comp)

    images = images + aux_images
    scene = sc.Scene(images=images, sources=sources, psfs=psfs,
                     transform=np.array(transform, dtype=np.int32),
                     lo=np.array(lo, dtype=np.float64), hi=np.array(hi, dtype=np.float64),
                     identities=info.identities)
    return scene, info


def wrap_model_images(model, info, outs):
    ims = [Model_Image(data=o, window=w.copy(), zeropoint=t.zeropoint, target_identity=t.identity)
           for o, w, t in zip(outs, info.windows, info.targets)]
    return Model_Image_List(ims) if info.is_list else ims[0]


def wrap_jacobian_images(model, info, outs):
    ims = [Jacobian_Image(parameters=list(info.identities), data=o, window=w.copy(), zeropoint=t.zeropoint,
                          target_identity=t.identity)
           for o, w, t in zip(outs, info.windows, info.targets)]
    return Jacobian_Image_List(ims) if info.is_list else ims[0]


def shard_scene(scene, rank, world):
    """Keep the images (bands / tiles) owned by ``rank`` (round robin) and the
    sources on them; the parameter table stays global so that every rank's
    J^T W J lands in the same (P, P) layout and a plain sum all-reduce merges
    them."""
    real = [i for i, im in enumerate(scene.images) if not im.aux]
    if world > len(real):
        # decided from the whole scene, so every rank raises the same error (a rank without images would otherwise stop
        # alone and leave the others waiting in their first collective)
        raise SpecificationConflict(
            f"a fit over {world} ranks needs at least {world} images (bands or tiles) to deal out, this one has "
            f"{len(real)}: use fewer ranks, cut the image into more tiles (LM(tiles=...)), or run replicas")
    keep = [i for k, i in enumerate(real) if k % world == rank] + [i for i, im in enumerate(scene.images) if im.aux]
    remap = {old: new for new, old in enumerate(keep)}
    srcs, new_index = [], {}
    for k, s in enumerate(scene.sources):
        if s.image in remap:
            s2 = sc.SceneSource(**{**s.__dict__})
            s2.image = remap[s.image]
            new_index[k] = len(srcs)
            srcs.append(s2)
    return sc.Scene(images=[scene.images[i] for i in keep], sources=srcs, psfs=_remap_psfs(scene.psfs, new_index),
                    transform=scene.transform, lo=scene.lo, hi=scene.hi, identities=scene.identities,
                    owners=scene.owners)


def _remap_psfs(psfs, new_index):
    """PSFs produced by an auxiliary PSF-model source point at it by index into the source list."""
    out = []
    for ps in psfs:
        if ps.source >= 0:
            ps = sc.ScenePSF(data=None, source=new_index[ps.source], shape=ps.shape)
        out.append(ps)
    return out


def _balanced_cuts(cost, n):
    """Cut positions 0 = c_0 < ... < c_n = len(cost) that split the 1-D cost profile into ``n`` runs of (nearly) equal
    cost; equal lengths where there is no cost to go by."""
    L = len(cost)
    total = float(cost.sum())
    if n <= 1 or total <= 0.0:
        return [round(k * L / n) for k in range(n + 1)]
    cum = np.concatenate([[0.0], np.cumsum(cost)])
    cuts = [0]
    for k in range(1, n):
        c = int(np.searchsorted(cum, total * k / n))
        cuts.append(min(max(c, cuts[-1] + 1), L - (n - k)))
    cuts.append(L)
    return cuts


def _tile_cost(im, srcs):
    """Per-pixel cost estimate of the sources of one image (H x W): what a tile pays for is, above all, the sub-pixel
    integration and the convolution of the models it evaluates; point sources and sky cost about a tenth per pixel."""
    cost = np.zeros((im.H, im.W))
    for s in srcs:
        x0, y0, w, h = s.out
        xa, ya, xb, yb = max(x0, 0), max(y0, 0), min(x0 + w, im.W), min(y0 + h, im.H)
        if xb <= xa or yb <= ya or (xb - xa) * (yb - ya) > 0.5 * im.H * im.W:
            continue                    # (a sky: the same everywhere)
        heavy = s.kind not in (sc.KIND_POINT, sc.KIND_FLAT_SKY) and s.integrate_mode != sc.INTEGRATE_NONE
        cost[ya:yb, xa:xb] += 1.0 if heavy else 0.1
    return cost


def tile_scene(scene, ny, nx, balance=True):
    """Cut every image of the scene into ``ny`` x ``nx`` tiles (one big image, as in a
    crowded field or a mosaic, partitioned by image tile).  Every tile becomes an image of its own (a
    view of data / weight / mask with the pixel origin moved), and a source is handed to every tile its
    output window intersects with that window clipped to the tile; its working windows (``fwd``, ``jac``:
    integration threshold, mean reference, curvature edge) are only translated, so the clipped source
    computes the same pixel values as the whole one.  Each pixel is owned by exactly one tile:
    chi^2, J^T W r and J^T W J of the tiles add up to those of the whole image, and ``shard_scene`` deals the
    tiles to the ranks.  A source near a tile border is evaluated (PSF border included) by each tile it
    touches -- redundant work instead of a halo exchange.  ``Scene.owners`` records the uncut models (the
    block-sparse J^T W J is laid out on them, identically on every rank)."""
    if ny * nx <= 1:
        return scene
    has_aux = any(im.aux for im in scene.images) or any(s.flags & sc.FLAG_AMP for s in scene.sources)
    # (the parameters of an auxiliary PSF model, or of the PSF model point sources are drawn from, are shared by every
    # source using it: no owner layout, dense solve)
    owners = None if has_aux else [(s.image, tuple(s.out), [sl for sl in s.slot if sl >= 0]) for s in scene.sources]
    images, sources, origin = [], [], []
    first_tile = []
    for im in scene.images:
        first_tile.append(len(images))
        if im.aux:                      # grid of an auxiliary PSF model: never cut
            images.append(im)
            origin.append((0, 0, im.W, im.H))
            continue
        # cuts: rows first, then every strip on its own, so that the tiles carry about the same estimated work
        # (``balance``; equal areas otherwise).  Where the cuts fall changes no pixel value.
        ii = len(first_tile) - 1
        cost = _tile_cost(im, [s for s in scene.sources if s.image == ii]) if balance else np.zeros((im.H, im.W))
        ys = _balanced_cuts(cost.sum(axis=1), ny)
        for a in range(ny):
            xs = _balanced_cuts(cost[ys[a]:ys[a + 1]].sum(axis=0), nx)
            for b in range(nx):
                y0, y1, x0, x1 = ys[a], ys[a + 1], xs[b], xs[b + 1]
                if y1 <= y0 or x1 <= x0:
                    continue
                cut = lambda t: None if t is None else t[y0:y1, x0:x1].contiguous()
                images.append(sc.SceneImage(H=y1 - y0, W=x1 - x0, S=np.array(im.S, dtype=np.float64).copy(),
                                            rij=np.asarray(im.rij, dtype=np.float64) - np.array([x0, y0], dtype=np.float64),
                                            rxy=np.array(im.rxy, dtype=np.float64).copy(),
                                            data=cut(im.data), weight=cut(im.weight), mask=cut(im.mask)))
                origin.append((x0, y0, x1 - x0, y1 - y0))
    n_tiles_of = first_tile[1:] + [len(images)]
    new_index = {}
    for k, s in enumerate(scene.sources):
        for t in range(first_tile[s.image], n_tiles_of[s.image]):
            tx, ty, tw, th = origin[t]
            ox, oy, ow, oh = s.out
            ix0, iy0 = max(ox, tx), max(oy, ty)
            ix1, iy1 = min(ox + ow, tx + tw), min(oy + oh, ty + th)
            if ix1 <= ix0 or iy1 <= iy0:
                continue
            if (s.flags & sc.FLAG_NORMALIZE) and (ix0, iy0, ix1, iy1) != (ox, oy, ox + ow, oy + oh):
                # the stamp of a normalised source must cover the window it is normalised over
                raise SpecificationConflict(
                    f"{s.name}: a point source drawn from a normalised PSF model cannot be cut by a tile border")
            s2 = sc.SceneSource(**{**s.__dict__})
            s2.image = t
            s2.owner = k
            s2.out = (ix0 - tx, iy0 - ty, ix1 - ix0, iy1 - iy0)
            s2.fwd = (s.fwd[0] - tx, s.fwd[1] - ty, s.fwd[2], s.fwd[3])
            s2.jac = (s.jac[0] - tx, s.jac[1] - ty, s.jac[2], s.jac[3])
            s2.mask_origin = (s.mask_origin[0] - tx, s.mask_origin[1] - ty)
            new_index[k] = len(sources)
            sources.append(s2)
    # the tiles come first, aux images last (as lower() lays them out)
    order = [i for i, im in enumerate(images) if not im.aux] + [i for i, im in enumerate(images) if im.aux]
    pos = {old: new for new, old in enumerate(order)}
    images = [images[i] for i in order]
    for s2 in sources:
        s2.image = pos[s2.image]
    return sc.Scene(images=images, sources=sources, psfs=_remap_psfs(scene.psfs, new_index), transform=scene.transform, lo=scene.lo,
                    hi=scene.hi, identities=scene.identities, owners=owners)
