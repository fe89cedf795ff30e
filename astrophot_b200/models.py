"""Model types of the forward-model-and-fit hot path.

Same factory and method surface as the reference (``AstroPhot_Model(model_type=
"sersic galaxy model", target=..., window=..., parameters={...}, psf_mode=...)``,
``model()``, ``.sample()``, ``.jacobian()``, ``.fit_mask()``; reference:
`models/core_model.py:26-502`, `model_object.py:24-444`,
`group_model_object.py:27-365`, `psf_model_object.py:20-330`,
`point_source.py:17-189`) but none of its arithmetic: every class here is a
*description* (kind, parameter DAG, window, sampling knobs).  ``sample`` and
``jacobian`` lower the description to flat tables (``lowering.py``) and hand
them to the sm_100a library through the C ABI (``cabi.py``).  There is no
torch/CPU evaluation path in this package.

In scope: sersic / exponential / gaussian / moffat / spline
galaxy, point, flat sky, group; sersic / exponential / gaussian / moffat /
moffat2d / spline psf models.  ``initialize()`` fills the parameters the user left
open from the data with the reference's recipes (``init_heuristics.py``, host numpy).
"""
from collections import OrderedDict
from copy import deepcopy
from typing import Optional

import numpy as np
import torch

from . import AP_config
from . import scene as sc
from .errors import (InvalidTarget, InvalidWindow, NameNotAllowed, SpecificationConflict,
                     UnrecognizedModel, InvalidParameter)
from .image import (Image, Image_List, Model_Image, PSF_Image, Target_Image, Target_Image_List,
                    Window, Window_List, Jacobian_Image)
from .param import Parameter_Node, Param_Unlock, Param_SoftLimits

__all__ = [
    "AstroPhot_Model", "Component_Model", "Galaxy_Model", "Sersic_Galaxy", "Exponential_Galaxy",
    "Gaussian_Galaxy", "Moffat_Galaxy", "Spline_Galaxy", "Point_Source", "Sky_Model", "Flat_Sky",
    "Group_Model", "PSF_Model", "Sersic_PSF", "Exponential_PSF", "Gaussian_PSF", "Moffat_PSF",
    "Moffat2D_PSF", "Spline_PSF",
]


def _all_subclasses(cls):
    out = []
    for sub in cls.__subclasses__():
        out.append(sub)
        out.extend(_all_subclasses(sub))
    return out


class AstroPhot_Model:
    """Base of every model; also the factory: ``AstroPhot_Model(model_type=...)``
    returns an instance of the class whose ``model_type`` string matches."""

    model_type = "model"
    parameter_specs = {}
    _parameter_order = ()
    usable = False
    model_names = []
    special_kwargs = ["parameters", "filename", "model_type", "usable", "models", "psf"]

    def __new__(cls, *, filename=None, model_type=None, **kwargs):
        if filename is not None:
            # a saved model names its own type (reference: `core_model.py:99-106`)
            model_type = AstroPhot_Model.load(filename)["model_type"]
        if model_type is not None:
            for M in _all_subclasses(AstroPhot_Model):
                if M.model_type == model_type:
                    return super().__new__(M)
            raise UnrecognizedModel(f"Unknown AstroPhot model type: {model_type}")
        return super().__new__(cls)

    def __init__(self, *, name=None, target=None, window=None, locked=False, **kwargs):
        self._window = None
        self._target = None
        self.name = name
        self.parameters = Parameter_Node(self.name)
        self.target = target
        self.window = window
        self._locked = locked
        self.mask = kwargs.get("mask", None)
        for key, val in kwargs.items():
            if key in self.special_kwargs or key == "mask":
                continue
            setattr(self, key, val)

    # -- naming ----------------------------------------------------------
    @property
    def name(self):
        return self._name

    @name.setter
    def name(self, name):
        if name is None:
            i = 0
            while f"{self.model_type} [{i}]" in AstroPhot_Model.model_names:
                i += 1
            name = f"{self.model_type} [{i}]"
        if ":" in name or "|" in name:
            raise NameNotAllowed(
                "characters '|' and ':' are reserved for internal model operations please do not include these in a model name")
        self._name = name
        AstroPhot_Model.model_names.append(name)

    def __del__(self):
        try:
            AstroPhot_Model.model_names.remove(self._name)
        except Exception:
            pass

    # -- parameters ------------------------------------------------------------
    @classmethod
    def build_parameter_specs(cls, user_specs=None):
        specs = {}
        for base in reversed(cls.__mro__):
            specs.update(deepcopy(base.__dict__.get("parameter_specs", {})))
        if isinstance(user_specs, dict):
            for p, spec in user_specs.items():
                if isinstance(spec, Parameter_Node):
                    specs[p] = spec
                elif isinstance(spec, dict):
                    specs.setdefault(p, {}).update(spec)
                else:
                    specs.setdefault(p, {})["value"] = spec
        return specs

    def build_parameters(self):
        for p in self._parameter_order:
            if p in self.parameters:
                continue
            spec = self.parameter_specs[p]
            if isinstance(spec, Parameter_Node):
                self.parameters.link(spec)
            elif isinstance(spec, dict):
                self.parameters.link(Parameter_Node(p, **spec))
            else:
                raise ValueError(f"unrecognized parameter specification for {p}")

    @property
    def parameter_order(self):
        return tuple(P.name for P in self.parameters)

    def __getitem__(self, key):
        return self.parameters[key]

    def __contains__(self, key):
        return key in self.parameters

    # -- target / window ---------------------------------------------------
    @property
    def target(self):
        return self._target

    @target.setter
    def target(self, tar):
        if not (tar is None or isinstance(tar, Target_Image)):
            raise InvalidTarget("AstroPhot_Model target must be a Target_Image instance.")
        self._target = tar

    @property
    def window(self):
        if self._window is None:
            if self.target is None:
                raise ValueError("This model has no target or window, these must be provided by the user")
            return self.target.window.copy()
        return self._window

    @window.setter
    def window(self, window):
        self.set_window(window)

    def set_window(self, window):
        if window is None:
            self._window = None
        elif isinstance(window, Window):
            self._window = window
        elif len(window) == 2:
            self._window = self.target.window.copy().crop_to_pixel(window)
        else:
            raise InvalidWindow(f"Unrecognized window format: {str(window)}")

    @property
    def locked(self):
        return self._locked

    @locked.setter
    def locked(self, val):
        self._locked = val

    @property
    def is_initialized(self):
        return all((not P.leaf) or (P.value is not None) for P in self.parameters)

    default_uncertainty = 1e-2   # relative start uncertainty when none is given (reference: `core_model.py:93`)

    def initialize(self, target=None, parameters=None, **kwargs):
        """When this returns every parameter should have a value (reference: `core_model.py:180-186`).  The
        subclasses fill in what the user left open from the data (``init_heuristics.py``); nothing here."""

    def _init_target(self, target):
        """The image to take start values from (reference: ``select_target``, `_shared_methods.py:33-49`)."""
        if target is None:
            return self.target
        if isinstance(target, Target_Image_List) and not isinstance(self.target, Image_List):
            for sub in target:
                if sub.identity == self.target.identity:
                    return sub
            raise RuntimeError(f"{self.name} could not find matching target to initialize with")
        return target

    def make_model_image(self, window=None):
        window = self.window if window is None else self.window & window
        return self.target[window].model_image()

    def fit_mask(self):
        return torch.zeros_like(self.target[self.window].mask)

    # -- evaluation through the native library -------------------------------
    def _plan(self, window=None):
        from .lowering import lower
        from .cabi import Plan

        scene, info = lower(self, window=window)
        return (Plan(scene) if scene.sources else None), scene, info

    def __call__(self, image=None, parameters=None, window=None, as_representation=False, **kwargs):
        if isinstance(parameters, (torch.Tensor, np.ndarray, list, tuple)):
            if as_representation:
                self.parameters.vector_set_representation(parameters)
            else:
                self.parameters.vector_set_values(parameters)
        return self.sample(image=image, window=window, **kwargs)

    def sample(self, image=None, window=None, parameters=None):
        """Model flux on ``window`` (default: the model's window); added into
        ``image`` when one is given (reference: `model_object.py:258-375`,
        `group_model_object.py:183-231`)."""
        from .lowering import wrap_model_images

        plan, scene, info = self._plan(window)
        x = torch.as_tensor(self.parameters.vector_values().numpy(), dtype=torch.float64,
                            device=AP_config.ap_device)
        if scene.sources:
            outs = plan.sample(x, as_rep=False)
        else:       # nothing but abstract base models: zero flux, nothing to launch
            outs = [torch.zeros((im.H, im.W), dtype=torch.float64, device=AP_config.ap_device) for im in scene.images]
        result = wrap_model_images(self, info, outs)
        if image is None:
            return result
        image += result
        return image

    def jacobian(self, parameters=None, as_representation=False, window=None, pass_jacobian=None, **kwargs):
        """d(model)/d(parameters) as a ``Jacobian_Image`` (columns ordered as
        ``parameters.vector_identities()``).  Materialising J is for small
        problems and tests; ``fit.LM`` never does (reference:
        `_model_methods.py:260-347`, `group_model_object.py:233-283`)."""
        from .lowering import wrap_jacobian_images

        if parameters is not None:
            if as_representation:
                self.parameters.vector_set_representation(parameters)
            else:
                self.parameters.vector_set_values(parameters)
        plan, scene, info = self._plan(window)
        vec = self.parameters.vector_representation() if as_representation else self.parameters.vector_values()
        x = torch.as_tensor(vec.numpy(), dtype=torch.float64, device=AP_config.ap_device)
        if scene.sources:
            outs = plan.jacobian(x, as_rep=as_representation)
        else:
            outs = [torch.zeros((im.H, im.W, scene.n_par), dtype=torch.float64, device=AP_config.ap_device)
                    for im in scene.images]
        result = wrap_jacobian_images(self, info, outs)
        if pass_jacobian is not None:
            pass_jacobian += result
            return pass_jacobian
        return result

    def total_flux(self, parameters=None, window=None):
        """Sum of the model image (reference: `core_model.py:265-268`)."""
        img = self(parameters=parameters, window=window)
        datas = [i.data for i in img.image_list] if isinstance(img, Image_List) else [img.data]
        return sum(torch.sum(d.to(torch.float64)) for d in datas).cpu()

    def total_flux_uncertainty(self, parameters=None, window=None):
        """Linear propagation of the parameter uncertainties (natural units) through the total flux: the column sums
        of the Jacobian are d(total flux)/d(parameter) (reference: `core_model.py:270-276`, which builds the dense
        Jacobian as well)."""
        if parameters is not None:
            self.parameters.vector_set_values(parameters)
        jac = self.jacobian(window=window)
        datas = [j.data for j in jac.image_list] if isinstance(jac, Image_List) else [jac.data]
        dF = sum(torch.sum(d.to(torch.float64).reshape(-1, d.shape[-1]), dim=0) for d in datas).cpu()
        unc = self.parameters.vector_uncertainty().to(torch.float64)
        return torch.sqrt(torch.sum((dF * unc) ** 2))

    def total_magnitude(self, parameters=None, window=None):
        """-2.5 log10(total flux) + zeropoint (reference: `core_model.py:278-282`)."""
        F = self.total_flux(parameters=parameters, window=window)
        return -2.5 * torch.log10(F) + self.target.zeropoint

    def total_magnitude_uncertainty(self, parameters=None, window=None):
        """|2.5 dF / (F ln 10)| (reference: `core_model.py:284-290`)."""
        F = self.total_flux(parameters=parameters, window=window)
        dF = self.total_flux_uncertainty(parameters=parameters, window=window)
        return torch.abs(2.5 * dF / (F * np.log(10)))

    # -- saved state (reference: `core_model.py:374-451`); the target image is never part of it ------------------
    def get_state(self, *args, **kwargs):
        return {"name": self.name, "model_type": self.model_type}

    def save(self, filename="AstroPhot.yaml"):
        """Write ``get_state()`` as YAML or JSON."""
        state = self.get_state()
        if isinstance(filename, str) and filename.endswith(".yaml"):
            import yaml
            with open(filename, "w") as f:
                yaml.dump(state, f, indent=2, sort_keys=False)      # (the order of a group's models is part of the state)
        elif isinstance(filename, str) and filename.endswith(".json"):
            import json
            with open(filename, "w") as f:
                json.dump(state, f, indent=2)
        else:
            raise ValueError(f"Unrecognized filename format: {filename}, must be one of: .json, .yaml "
                             "(hdf5 needs h5py, which astrophot_b200 does not depend on)")

    @classmethod
    def load(cls, filename="AstroPhot.yaml"):
        """The state dictionary of a saved model: a dict (returned as it is), an open text stream, a .yaml or .json file."""
        import io
        if isinstance(filename, dict):
            return filename
        if isinstance(filename, io.TextIOBase):
            import yaml
            return yaml.load(filename, Loader=yaml.FullLoader)
        if isinstance(filename, str) and filename.endswith(".yaml"):
            import yaml
            with open(filename, "r") as f:
                return yaml.load(f, Loader=yaml.FullLoader)
        if isinstance(filename, str) and filename.endswith(".json"):
            import json
            with open(filename, "r") as f:
                return json.load(f)
        raise ValueError(f"Unrecognized filename format: {filename}, must be one of: .json, .yaml or python dictionary.")

    @classmethod
    def List_Models(cls, usable=None):
        models = _all_subclasses(cls)
        if cls is not AstroPhot_Model and not issubclass(cls, PSF_Model):
            # PSF models share the component machinery here but are a separate family in the reference
            # (psf_model_object.py: PSF_Model(AstroPhot_Model)): Component_Model.List_Models() does not list them
            models = [m for m in models if not issubclass(m, PSF_Model)]
        if usable is not None:
            models = [m for m in models if m.usable is usable]
        return models

    @classmethod
    def List_Model_Names(cls, usable=None):
        return sorted((m.model_type for m in cls.List_Models(usable)), key=lambda n: n[::-1])

    def __eq__(self, other):
        return self is other

    __hash__ = object.__hash__

    def __str__(self):
        return str(self.parameters)


def _sersic_b(n):
    """b(n) of the Sérsic profile (reference: `utils/conversions/functions.py:7-22`)."""
    return (2 * n - 1 / 3 + 4 / (405 * n) + 46 / (25515 * n**2) + 131 / (1148175 * n**3)
            - 2194697 / (30690717750 * n**4))


def _pval(model, name):
    return model.parameters[name].value.to(torch.float64).reshape(-1)[0]


def _moffat_total_flux(I0, n, Rd, q):
    """Reference: `utils/conversions/functions.py:227-237`."""
    return I0 * np.pi * Rd**2 * q / (n - 1)


class Component_Model(AstroPhot_Model):
    """A single light source (reference: `model_object.py:24-111` for the
    attribute list and defaults)."""

    model_type = AstroPhot_Model.model_type
    parameter_specs = {"center": {"units": "arcsec", "uncertainty": [0.1, 0.1]}}
    _parameter_order = ("center",)

    psf_mode = "none"                 # none, full
    psf_convolve_mode = "fft"         # fft, direct  (same result; the library picks by PSF size)
    psf_subpixel_shift = "bilinear"   # bilinear, lanczos:k, none
    sampling_mode = "midpoint"        # midpoint, simpsons, quad:N, trapezoid
    sampling_tolerance = 1e-2
    integrate_mode = "threshold"      # none, threshold
    integrate_max_depth = 3
    integrate_gridding = 5
    integrate_quad_level = 3
    jacobian_chunksize = 10           # kept for API compatibility; unused (J is never chunked here)
    image_chunksize = 1000
    softening = 1e-3

    track_attrs = ["psf_mode", "psf_convolve_mode", "psf_subpixel_shift", "sampling_mode", "sampling_tolerance",
                   "integrate_mode", "integrate_max_depth", "integrate_gridding", "integrate_quad_level",
                   "jacobian_chunksize", "image_chunksize", "softening"]
    _kind = None
    _ref_mode = sc.REF_MEAN
    _flags = 0

    def __init__(self, *, name=None, **kwargs):
        self._psf = None
        super().__init__(name=name, **kwargs)
        # as in the reference (core_model.py:128-134 sets user attributes before the model's own parameters exist):
        # the parameters of an auxiliary PSF model come first in the parameter vector
        if "filename" in kwargs:
            self.load(kwargs["filename"], new_name=name)
            return
        if "psf" in kwargs:
            self.psf = kwargs["psf"]
        self.parameter_specs = self.build_parameter_specs(kwargs.get("parameters", None))
        self.build_parameters()
        if isinstance(kwargs.get("parameters", None), torch.Tensor):
            self.parameters.value = kwargs["parameters"]

    def get_state(self, save_params=True):
        """Name, type, window, parameters, the knobs that differ from the class defaults and the model's own PSF
        (reference: `model_object.py:408-428`)."""
        state = super().get_state()
        state["window"] = self.window.get_state()
        if save_params:
            state["parameters"] = self.parameters.get_state()
        state["target_identity"] = getattr(self, "_target_identity", None)
        if isinstance(self._psf, PSF_Image):
            state["psf"] = {"type": "PSF_Image", "data": self._psf.data.detach().cpu().tolist(),
                            "window": self._psf.window.get_state()}
        elif isinstance(self._psf, AstroPhot_Model):
            state["psf"] = self._psf.get_state()
            # (the grid a PSF model is sampled on is its target's; zeros stand in for the pixel data)
            state["psf"]["target_window"] = self._psf.target.window.get_state()
        for key in self.track_attrs:
            if getattr(self, key) != getattr(self.__class__, key):
                state[key] = getattr(self, key)
        return state

    def load(self, filename="AstroPhot.yaml", new_name=None):
        """Take window, knobs, parameters and PSF from a saved state; the target stays the one this model was given
        (reference: `_model_methods.py:437-482`)."""
        from .image import Window

        state = AstroPhot_Model.load(filename)
        self.name = state["name"] if new_name is None else new_name
        self.window = Window(state=state["window"])
        self._target_identity = state.get("target_identity", None)
        self.target = self.target
        for key in self.track_attrs:
            if key in state:
                setattr(self, key, state[key])
        if isinstance(state["parameters"], Parameter_Node):
            self.parameters = state["parameters"]
        else:
            self.parameters = Parameter_Node(self.name, state=state["parameters"]).relink()
        if "psf" in state:
            if state["psf"].get("type", "AstroPhot_Model") == "PSF_Image":
                self._psf = PSF_Image(data=np.array(state["psf"]["data"]), window=Window(state=state["psf"]["window"]))
            else:
                sub = dict(state["psf"])
                sub["parameters"] = self.parameters[sub["name"]]
                grid = Window(state=sub["target_window"])
                ptar = PSF_Image(data=np.zeros((int(grid.pixel_shape[1]), int(grid.pixel_shape[0]))), window=grid)
                self.set_aux_psf(AstroPhot_Model(name=sub["name"], filename=sub, target=ptar), add_parameters=False)
        return state

    @property
    def psf(self):
        if self._psf is None:
            try:
                return self.target.psf
            except AttributeError:
                return None
        return self._psf

    @psf.setter
    def psf(self, val):
        if val is None:
            self._psf = None
        elif isinstance(val, PSF_Image):
            self._psf = val
        elif isinstance(val, AstroPhot_Model):
            self.set_aux_psf(val)
        else:
            self._psf = PSF_Image(data=val, pixelscale=self.target.pixelscale)
            AP_config.ap_logger.warning(
                "Setting PSF with pixel matrix, assuming target pixelscale is the same as PSF pixelscale. To remove "
                "this warning, set PSFs as an ap.image.PSF_Image or ap.models.AstroPhot_Model object instead.")

    def set_aux_psf(self, aux_psf, add_parameters=True):
        self._psf = aux_psf
        if add_parameters:
            self.parameters.link(aux_psf.parameters)

    _init_family = None      # key of init_heuristics.PROFILE_FITS for the parametric profile families

    @torch.no_grad()
    def initialize(self, target=None, parameters=None, **kwargs):
        """Fill in the parameters the user left open from the data, in the reference's order: centre
        (`model_object.py:176-233`), shape (`galaxy_model_object.py:56-113`), then the family's own parameters."""
        target = self._init_target(target)
        self._init_center(target)
        self._init_shape(target)
        self._init_own(target)

    def _init_shape(self, target):
        pass

    def _init_own(self, target):
        if self._init_family is not None:
            self._init_profile(target, self._init_family)

    def _init_center(self, target):
        """The window's centre refined by an iterated centre-of-light search, when no centre was given."""
        from . import init_heuristics as ih

        c = self.parameters["center"]
        if c.value is not None:
            return
        with Param_Unlock(c), Param_SoftLimits(c):
            c.value = self.window.center
        if c.locked or target is None:
            return
        area = target[self.window]
        pc = area.plane_to_pixel(c.value)
        dat = area.data.detach().cpu().numpy()
        com = ih.light_centroid((float(pc[1]), float(pc[0])), dat)
        if np.any(np.array(com) < 0) or np.any(np.array(com) >= np.array(dat.shape)):
            AP_config.ap_logger.warning("center of mass failed, using center of window")
            return
        c.value = area.pixel_to_plane(torch.tensor([com[1], com[0]], dtype=torch.float64))

    # -- pieces shared by the initialize() methods of the profile families ------------------------
    def _init_radii(self, area):
        """Radius of every pixel of ``area`` in the model's own metric (rotation, axis ratio, softening)."""
        coords = area.get_coordinate_meshgrid()
        cen = self.parameters["center"].value.to(torch.float64)
        X, Y = coords[0] - cen[0], coords[1] - cen[1]
        X, Y = self._init_transform(X, Y, area)
        return torch.sqrt(X**2 + Y**2 + self.softening**2).detach().cpu().numpy()

    def _init_transform(self, X, Y, area):
        return X, Y

    def _init_profile(self, target, family):
        """Start values of a parametric profile: Nelder-Mead fit to the binned radial profile
        (reference: `_shared_methods.py:125-164`)."""
        from . import init_heuristics as ih

        names, prof, x0_of = ih.PROFILE_FITS[family]
        pars = [self.parameters[k] for k in names]
        if all(P.value is not None for P in pars):
            return
        area = target[self.window]
        mask = area.mask.detach().cpu().numpy().astype(bool) if area.has_mask else None
        R, I, _ = ih.radial_profile(area.data.detach().cpu().numpy().astype(np.float64), mask, self._init_radii(area),
                                    float(area.pixel_area))
        x0 = list(x0_of(R, I))
        for k, P in enumerate(pars):
            if P.value is not None:
                x0[k] = float(P.value)
        x, ok, spread = ih.fit_profile(R, I, prof, x0)
        for k, P in enumerate(pars):
            with Param_Unlock(P), Param_SoftLimits(P):
                if P.value is None:
                    P.value = x[k] if ok else x0[k]
                if P.uncertainty is None:
                    P.uncertainty = spread[k]

    @property
    def target(self):
        return self._target

    @target.setter
    def target(self, tar):
        if not (tar is None or isinstance(tar, Target_Image)):
            raise InvalidTarget("AstroPhot_Model target must be a Target_Image instance.")
        ident = getattr(self, "_target_identity", None)
        if isinstance(tar, Target_Image_List) and ident is not None:
            for sub in tar:
                if sub.identity == ident:
                    tar = sub
                    break
            else:
                raise InvalidTarget(
                    f"Could not find target in Target_Image_List with matching identity to {self.name}: {ident}")
        self._target = tar
        if tar is not None and not isinstance(tar, Image_List):
            self._target_identity = tar.identity


class Galaxy_Model(Component_Model):
    """Adds axis ratio ``q`` and position angle ``PA`` (reference:
    `galaxy_model_object.py:44-53`)."""

    model_type = f"galaxy {Component_Model.model_type}"
    parameter_specs = {
        "q": {"units": "b/a", "limits": (0, 1), "uncertainty": 0.03},
        "PA": {"units": "radians", "limits": (0, np.pi), "cyclic": True, "uncertainty": 0.06},
    }
    _parameter_order = Component_Model._parameter_order + ("q", "PA")

    def _init_transform(self, X, Y, area):
        th = -(self.parameters["PA"].value.to(torch.float64) - area.north)
        s, c = torch.sin(th), torch.cos(th)
        return c * X - s * Y, (s * X + c * Y) / self.parameters["q"].value.to(torch.float64)

    def _init_shape(self, target):
        """PA from the second angular moment of the light, q from the m=2 Fourier amplitude of trial isophotes
        (reference: `galaxy_model_object.py:56-113`)."""
        from . import init_heuristics as ih

        PA, q = self.parameters["PA"], self.parameters["q"]
        if PA.value is not None and q.value is not None:
            return
        area = target[self.window]
        dat = area.data.detach().cpu().numpy().astype(np.float64).copy()
        mask = area.mask.detach().cpu().numpy().astype(bool) if area.has_mask else None
        if mask is not None:
            dat[mask] = np.median(dat[~mask])
        edge = ih.edge_pixels(dat)
        edge_average = np.nanmedian(edge)
        edge_scatter = ih.half_spread(edge[np.isfinite(edge)])
        cen = self.parameters["center"].value.to(torch.float64)
        pc = area.plane_to_pixel(cen)
        if PA.value is None:
            coords = area.get_coordinate_meshgrid()
            X, Y = (coords[0] - cen[0]).cpu().numpy(), (coords[1] - cen[1]).cpu().numpy()
            w = dat - edge_average
            ang = ih.moment_position_angle(w, X, Y) if mask is None else ih.moment_position_angle(w[~mask], X[~mask], Y[~mask])
            with Param_Unlock(PA), Param_SoftLimits(PA):
                PA.value = (ang + area.north) % np.pi
                if PA.uncertainty is None:
                    PA.uncertainty = (5 * np.pi / 180) * torch.ones_like(PA.value)
        if q.value is None:
            q_samples = np.linspace(0.2, 0.9, 15)
            # (quirk 1 of init_heuristics: the centre goes in as (pixel-y, pixel-x) and is used as (x, y))
            best = ih.axis_ratio_scan(area.data.detach().cpu().numpy().astype(np.float64) - edge_average,
                                      float(pc[1]), float(pc[0]), 3 * edge_scatter,
                                      float(PA.value) - target.north, q_samples)
            with Param_Unlock(q), Param_SoftLimits(q):
                q.value = best
                if q.uncertainty is None:
                    q.uncertainty = q.value * self.default_uncertainty


class Sersic_Galaxy(Galaxy_Model):
    """I(R) = Ie exp(-b_n ((R/Re)^(1/n) - 1)) (reference: `sersic_model.py:41-91`)."""

    model_type = f"sersic {Galaxy_Model.model_type}"
    parameter_specs = {
        "n": {"units": "none", "limits": (0.36, 8), "uncertainty": 0.05},
        "Re": {"units": "arcsec", "limits": (0, None)},
        "Ie": {"units": "log10(flux/arcsec^2)"},
    }
    _parameter_order = Galaxy_Model._parameter_order + ("n", "Re", "Ie")
    usable = True
    _init_family = "sersic"
    _kind = sc.KIND_SERSIC
    _ref_mode = sc.REF_SERSIC_FLUX    # total_flux / numel, sersic_model.py:87-89

    def total_flux(self, parameters=None, window=None):
        """Analytic flux to infinity, not the image sum (reference: `sersic_model.py:79-85`,
        `utils/conversions/functions.py:168-190`)."""
        if isinstance(parameters, (torch.Tensor, np.ndarray, list, tuple)):
            self.parameters.vector_set_values(parameters)
        n, Re, q, Ie = (_pval(self, k) for k in ("n", "Re", "q", "Ie"))
        bn = _sersic_b(n)
        return 2 * np.pi * 10**Ie * Re**2 * q * n * (torch.exp(bn) * bn ** (-2 * n)) * torch.exp(torch.lgamma(2 * n))


class Exponential_Galaxy(Galaxy_Model):
    model_type = f"exponential {Galaxy_Model.model_type}"
    parameter_specs = {
        "Re": {"units": "arcsec", "limits": (0, None)},
        "Ie": {"units": "log10(flux/arcsec^2)"},
    }
    _parameter_order = Galaxy_Model._parameter_order + ("Re", "Ie")
    usable = True
    _init_family = "exponential"
    _kind = sc.KIND_EXPONENTIAL


class Gaussian_Galaxy(Galaxy_Model):
    model_type = f"gaussian {Galaxy_Model.model_type}"
    parameter_specs = {
        "sigma": {"units": "arcsec", "limits": (0, None)},
        "flux": {"units": "log10(flux)"},
    }
    _parameter_order = Galaxy_Model._parameter_order + ("sigma", "flux")
    usable = True
    _init_family = "gaussian"
    _kind = sc.KIND_GAUSSIAN


class Moffat_Galaxy(Galaxy_Model):
    model_type = f"moffat {Galaxy_Model.model_type}"
    parameter_specs = {
        "n": {"units": "none", "limits": (0.1, 10), "uncertainty": 0.05},
        "Rd": {"units": "arcsec", "limits": (0, None)},
        "I0": {"units": "log10(flux/arcsec^2)"},
    }
    _parameter_order = Galaxy_Model._parameter_order + ("n", "Rd", "I0")
    usable = True
    _init_family = "moffat"
    _kind = sc.KIND_MOFFAT

    def total_flux(self, parameters=None, window=None):
        """Analytic flux to infinity (reference: `moffat_model.py:60-66`)."""
        if isinstance(parameters, (torch.Tensor, np.ndarray, list, tuple)):
            self.parameters.vector_set_values(parameters)
        return _moffat_total_flux(10 ** _pval(self, "I0"), _pval(self, "n"), _pval(self, "Rd"), _pval(self, "q"))


class Spline_Galaxy(Galaxy_Model):
    """log10 brightness is a cubic spline through ``I(R)`` at radii ``I(R).prof``
    (reference: `spline_model.py:27-53`)."""

    model_type = f"spline {Galaxy_Model.model_type}"
    parameter_specs = {"I(R)": {"units": "log10(flux/arcsec^2)"}}
    _parameter_order = Galaxy_Model._parameter_order + ("I(R)",)
    usable = True
    extend_profile = True
    _kind = sc.KIND_SPLINE

    def _init_own(self, target):
        """Node radii growing by 20 % per node out to the window's half-diagonal, node values from the binned radial
        profile (reference: `_shared_methods.py:426-453`)."""
        from . import init_heuristics as ih

        P = self.parameters["I(R)"]
        if P.value is not None and P.prof is not None:
            return
        if P.prof is None:
            half = self.window.shape.to(torch.float64) / 2
            step = 2 * float(target.pixel_length)
            prof = [0.0, step]
            while prof[-1] < float(torch.max(half)):
                prof.append(prof[-1] + max(step, prof[-1] * 0.2))
            prof = prof[:-2] + [float(torch.sqrt(torch.sum(half**2)))]
            P.prof = prof
        pr = P.prof.detach().cpu().numpy().astype(np.float64)
        area = target[self.window]
        mask = area.mask.detach().cpu().numpy().astype(bool) if area.has_mask else None
        bins = [pr[0]] + list((pr[:-1] + pr[1:]) / 2) + [pr[-1] * 100]
        _, I, S = ih.radial_profile(area.data.detach().cpu().numpy().astype(np.float64), mask, self._init_radii(area),
                                    float(area.pixel_area), rad_bins=bins)
        with Param_Unlock(P), Param_SoftLimits(P):
            P.value = I
            P.uncertainty = S


class Point_Source(Component_Model):
    """Delta function times the PSF (reference: `point_source.py:17-189`)."""

    model_type = f"point {Component_Model.model_type}"
    parameter_specs = {"flux": {"units": "log10(flux)"}}
    _parameter_order = Component_Model._parameter_order + ("flux",)
    usable = True
    _kind = sc.KIND_POINT

    def __init__(self, *args, **kwargs):
        super().__init__(*args, **kwargs)
        if self.psf is None:
            raise ValueError("Point_Source needs psf information")

    def _init_own(self, target):
        """Flux: the light in the window above the median of its edge pixels (reference: `point_source.py:40-65`)."""
        from . import init_heuristics as ih

        F = self.parameters["flux"]
        if F.value is not None:
            return
        area = target[self.window]
        dat = area.data.detach().cpu().numpy().astype(np.float64)
        with Param_Unlock(F), Param_SoftLimits(F):
            F.value = np.log10(np.abs(np.sum(dat - np.median(ih.edge_pixels(dat)))))
            F.uncertainty = torch.std(area.data.to(torch.float64).cpu()) / (np.log(10) * 10 ** F.value)

    @property
    def psf_mode(self):
        return "full"

    @psf_mode.setter
    def psf_mode(self, value):
        pass


class Sky_Model(Component_Model):
    """Sky is never convolved nor sub-pixel integrated (reference:
    `sky_model_object.py:21-35`)."""

    model_type = f"sky {Component_Model.model_type}"
    parameter_specs = {"center": {"units": "arcsec", "locked": True, "uncertainty": 0.0}}

    @property
    def psf_mode(self):
        return "none"

    @psf_mode.setter
    def psf_mode(self, val):
        pass

    @property
    def integrate_mode(self):
        return "none"

    @integrate_mode.setter
    def integrate_mode(self, val):
        pass


class Flat_Sky(Sky_Model):
    model_type = f"flat {Sky_Model.model_type}"
    parameter_specs = {"F": {"units": "log10(flux/arcsec^2)"}}
    _parameter_order = Sky_Model._parameter_order + ("F",)
    usable = True
    _kind = sc.KIND_FLAT_SKY

    def _init_own(self, target):
        """F: median of the window per unit area; uncertainty: the 1-sigma range of the pixels over sqrt(window size)
        (reference: `flatsky_model.py:29-52`)."""
        from . import init_heuristics as ih

        F = self.parameters["F"]
        if F.value is not None and F.uncertainty is not None:
            return
        dat = target[self.window].data.detach().cpu().numpy().astype(np.float64)
        area = float(target.pixel_area)
        with Param_Unlock(F), Param_SoftLimits(F):
            if F.value is None:
                # (torch.median: the LOWER of the two middle values of an even count, as in the reference)
                F.value = np.log10(np.abs(float(torch.median(torch.as_tensor(dat)))) / area)
            if F.uncertainty is None:
                spread = ih.half_spread(dat, 31.731 / 2, 100 - 31.731 / 2) / area
                F.uncertainty = (spread / np.sqrt(np.prod(self.window.shape.detach().cpu().numpy()))) / (
                    10 ** float(F.value) * np.log(10))


class Plane_Sky(Sky_Model):
    """Sky brightness plane  I = pixel_area F + X dx + Y dy  about the (locked) centre, natural flux units
    (reference: `planesky_model.py:13-74`)."""

    model_type = f"plane {Sky_Model.model_type}"
    parameter_specs = {"F": {"units": "flux/arcsec^2"}, "delta": {"units": "flux/arcsec"}}
    _parameter_order = Sky_Model._parameter_order + ("F", "delta")
    usable = True
    _kind = sc.KIND_PLANE_SKY
    _flags = sc.FLAG_RADIAL        # no rotation / axis ratio elements

    def _init_own(self, target):
        """F: median of the window per unit area; no slope (reference: `planesky_model.py:37-63`)."""
        from . import init_heuristics as ih

        F, delta = self.parameters["F"], self.parameters["delta"]
        dat = None
        if F.value is None or F.uncertainty is None:
            dat = target[self.window].data.detach().cpu().numpy().astype(np.float64)
        with Param_Unlock(F), Param_SoftLimits(F):
            if F.value is None:
                F.value = np.median(dat) / float(target.pixel_area)
            if F.uncertainty is None:
                F.uncertainty = ih.half_spread(dat, 31.731 / 2, 100 - 31.731 / 2) / np.sqrt(
                    np.prod(self.window.shape.detach().cpu().numpy()))
        with Param_Unlock(delta), Param_SoftLimits(delta):
            if delta.value is None:
                delta.value = [0.0, 0.0]
                delta.uncertainty = [self.default_uncertainty, self.default_uncertainty]


# ---------------------------------------------------------------------------
# PSF models: target is a PSF_Image, centre locked at (0, 0), normalised
# ---------------------------------------------------------------------------
class PSF_Model(Component_Model):
    model_type = f"psf {AstroPhot_Model.model_type}"
    parameter_specs = {"center": {"units": "arcsec", "value": (0.0, 0.0), "uncertainty": (0.1, 0.1), "locked": True}}
    _parameter_order = ("center",)
    model_integrated = False
    normalize_psf = True
    track_attrs = Component_Model.track_attrs + ["normalize_psf", "model_integrated"]      # saved with the model
    sampling_mode = "simpsons"
    sampling_tolerance = 1e-3
    integrate_mode = "threshold"
    _flags = sc.FLAG_RADIAL

    @property
    def psf_mode(self):
        return "none"

    @psf_mode.setter
    def psf_mode(self, val):
        pass

    @property
    def target(self):
        return self._target

    @target.setter
    def target(self, tar):
        if not (tar is None or isinstance(tar, PSF_Image)):
            raise InvalidTarget("PSF_Model target must be a PSF_Image instance.")
        self._target = tar

    def make_model_image(self, window=None):
        window = self.window if window is None else self.window & window
        return self.target[window].model_image()

    def fit_mask(self):
        return torch.zeros_like(self.target[self.window].mask)


class Sersic_PSF(PSF_Model):
    model_type = f"sersic {PSF_Model.model_type}"
    parameter_specs = {
        "n": {"units": "none", "limits": (0.36, 8), "uncertainty": 0.05},
        "Re": {"units": "arcsec", "limits": (0, None)},
        "Ie": {"units": "log10(flux/arcsec^2)", "value": 0.0, "locked": True},
    }
    _parameter_order = PSF_Model._parameter_order + ("n", "Re", "Ie")
    usable = True
    _init_family = "sersic"
    _kind = sc.KIND_SERSIC


class Exponential_PSF(PSF_Model):
    model_type = f"exponential {PSF_Model.model_type}"
    parameter_specs = {
        "Re": {"units": "arcsec", "limits": (0, None)},
        "Ie": {"units": "log10(flux/arcsec^2)", "value": 0.0, "locked": True},
    }
    _parameter_order = PSF_Model._parameter_order + ("Re", "Ie")
    usable = True
    _init_family = "exponential"
    _kind = sc.KIND_EXPONENTIAL


class Gaussian_PSF(PSF_Model):
    model_type = f"gaussian {PSF_Model.model_type}"
    parameter_specs = {
        "sigma": {"units": "arcsec", "limits": (0, None)},
        "flux": {"units": "log10(flux)", "value": 0.0, "locked": True},
    }
    _parameter_order = PSF_Model._parameter_order + ("sigma", "flux")
    usable = True
    _init_family = "gaussian"
    _kind = sc.KIND_GAUSSIAN


class Moffat_PSF(PSF_Model):
    model_type = f"moffat {PSF_Model.model_type}"
    parameter_specs = {
        "n": {"units": "none", "limits": (0.1, 10), "uncertainty": 0.05},
        "Rd": {"units": "arcsec", "limits": (0, None)},
        "I0": {"units": "log10(flux/arcsec^2)", "value": 0.0, "locked": True},
    }
    _parameter_order = PSF_Model._parameter_order + ("n", "Rd", "I0")
    usable = True
    _init_family = "moffat"
    _kind = sc.KIND_MOFFAT

    def total_flux(self, parameters=None, window=None):
        """Analytic flux to infinity with q = 1 (reference: `moffat_model.py:111-117`)."""
        if isinstance(parameters, (torch.Tensor, np.ndarray, list, tuple)):
            self.parameters.vector_set_values(parameters)
        return _moffat_total_flux(10 ** _pval(self, "I0"), _pval(self, "n"), _pval(self, "Rd"), 1.0)


class Moffat2D_PSF(PSF_Model):
    total_flux = Moffat_PSF.total_flux       # the reference's Moffat2D_PSF inherits it (q = 1 there too)

    def _init_shape(self, target):
        """Fixed start shape (reference: `moffat_model.py:139-148`)."""
        for name, start in (("q", 0.9), ("PA", 0.1)):
            P = self.parameters[name]
            with Param_Unlock(P), Param_SoftLimits(P):
                if P.value is None:
                    P.value = start

    def _init_transform(self, X, Y, area):
        return Galaxy_Model._init_transform(self, X, Y, area)
    model_type = f"moffat2d {PSF_Model.model_type}"
    parameter_specs = {
        "q": {"units": "b/a", "limits": (0, 1), "uncertainty": 0.03},
        "PA": {"units": "radians", "limits": (0, np.pi), "cyclic": True, "uncertainty": 0.06},
        "n": {"units": "none", "limits": (0.1, 10), "uncertainty": 0.05},
        "Rd": {"units": "arcsec", "limits": (0, None)},
        "I0": {"units": "log10(flux/arcsec^2)", "value": 0.0, "locked": True},
    }
    _parameter_order = PSF_Model._parameter_order + ("n", "Rd", "I0", "q", "PA")   # (moffat_model.py:134: q, PA come last)
    usable = True
    _init_family = "moffat"
    _kind = sc.KIND_MOFFAT
    _flags = 0


class Spline_PSF(PSF_Model):
    model_type = f"spline {PSF_Model.model_type}"
    parameter_specs = {"I(R)": {"units": "log10(flux/arcsec^2)"}}
    _parameter_order = PSF_Model._parameter_order + ("I(R)",)
    usable = True
    extend_profile = True
    _kind = sc.KIND_SPLINE

    def _init_own(self, target):
        """Node radii growing by 20 % per node out to the window's half-diagonal, node values from the binned radial
        profile (reference: `_shared_methods.py:426-453`)."""
        from . import init_heuristics as ih

        P = self.parameters["I(R)"]
        if P.value is not None and P.prof is not None:
            return
        if P.prof is None:
            half = self.window.shape.to(torch.float64) / 2
            step = 2 * float(target.pixel_length)
            prof = [0.0, step]
            while prof[-1] < float(torch.max(half)):
                prof.append(prof[-1] + max(step, prof[-1] * 0.2))
            prof = prof[:-2] + [float(torch.sqrt(torch.sum(half**2)))]
            P.prof = prof
        pr = P.prof.detach().cpu().numpy().astype(np.float64)
        area = target[self.window]
        mask = area.mask.detach().cpu().numpy().astype(bool) if area.has_mask else None
        bins = [pr[0]] + list((pr[:-1] + pr[1:]) / 2) + [pr[-1] * 100]
        _, I, S = ih.radial_profile(area.data.detach().cpu().numpy().astype(np.float64), mask, self._init_radii(area),
                                    float(area.pixel_area), rad_bins=bins)
        with Param_Unlock(P), Param_SoftLimits(P):
            P.value = I
            P.uncertainty = S


# ---------------------------------------------------------------------------
# Group
# ---------------------------------------------------------------------------
class Group_Model(AstroPhot_Model):
    """Sum of sub-models (reference: `group_model_object.py:27-365`).  On the
    device a group is just a longer source table; nothing is evaluated per
    sub-model in Python."""

    model_type = f"group {AstroPhot_Model.model_type}"
    usable = True

    def __init__(self, *, name=None, models=None, **kwargs):
        if "model" in kwargs:
            AP_config.ap_logger.warning("kwarg `model` is not used in Group_Model, did you mean `models` instead?")
        self.models = OrderedDict()
        self._psf_mode = "none"
        super().__init__(name=name, models=models, **kwargs)
        if models is not None:
            self.add_model(models)
        if "filename" in kwargs:
            self.load(kwargs["filename"], new_name=name)
        self.update_window()
        if "psf_mode" in kwargs:
            self.psf_mode = kwargs["psf_mode"]

    def add_model(self, model, _update=True):
        if isinstance(model, (tuple, list)):
            # one window update for the whole list: the reference re-unions every window per added
            # model (group_model_object.py:70-90), which is quadratic in the number of sources
            for mod in model:
                self.add_model(mod, _update=False)
            self.update_window()
            return
        if model.name in self.models:
            if model is self.models[model.name]:
                return
            raise KeyError(
                f"{self.name} already has model with name {model.name}, every model must have a unique name.")
        self.models[model.name] = model
        self.parameters.link(model.parameters)
        # the group's psf_mode overrides the sub-model's; its target is handed to sub-models that have none
        # (group_model_object.py:85-89,306-323)
        model.psf_mode = self.psf_mode
        self._offer_target(model, self.target)
        if _update:
            self.update_window()

    def update_window(self, include_locked=False):
        if isinstance(self.target, Image_List):
            wins = [None] * len(self.target.image_list)
            for model in self.models.values():
                if model.locked and not include_locked:
                    continue
                if isinstance(model.target, Image_List):
                    pairs = zip(model.target, model.window)
                else:
                    pairs = [(model.target, model.window)]
                for tar, win in pairs:
                    i = self.target.index(tar)
                    if wins[i] is None:
                        wins[i] = win.copy()
                    else:
                        wins[i] |= win
            self._window = Window_List(wins) if all(w is not None for w in wins) and wins else None
        else:
            new = None
            for model in self.models.values():
                if model.locked and not include_locked:
                    continue
                if new is None:
                    new = model.window.copy()
                else:
                    new |= model.window
            self._window = new

    @property
    def target(self):
        return self._target

    @target.setter
    def target(self, tar):
        if not (tar is None or isinstance(tar, Target_Image)):
            raise InvalidTarget("Group_Model target must be a Target_Image instance.")
        self._target = tar
        for model in getattr(self, "models", {}).values():
            self._offer_target(model, tar)

    @staticmethod
    def _offer_target(model, tar):
        """group_model_object.py:312-323: a sub-model without a target takes the group's; one that has its own keeps it
        (a mismatch is only reported); with an image list every sub-model keeps the member it was built on."""
        if tar is None or isinstance(tar, Image_List):
            return
        if model.target is None:
            model.target = tar
        elif isinstance(model.target, Image_List) or model.target.identity != tar.identity:
            AP_config.ap_logger.warning(
                f"Group_Model target does not match model {model.name} target. This may cause issues. Use the same "
                "Target_Image object for all relevant models.")

    @property
    def psf_mode(self):
        return self._psf_mode

    @psf_mode.setter
    def psf_mode(self, value):
        self._psf_mode = value
        for model in getattr(self, "models", {}).values():
            model.psf_mode = value

    def get_state(self, save_params=True):
        """The group's parameter graph once, and every sub-model's state without its parameters (reference:
        `group_model_object.py:325-337`)."""
        state = AstroPhot_Model.get_state(self)
        if save_params:
            state["parameters"] = self.parameters.get_state()
        state["models"] = {m.name: m.get_state(save_params=False) for m in self.models.values()}
        if self.psf_mode != "none":
            state["psf_mode"] = self.psf_mode
        return state

    def load(self, filename="AstroPhot.yaml", new_name=None):
        """Rebuild the parameter graph and hand every sub-model its branch of it; sub-models the group does not hold
        yet are created on the group's target (reference: `group_model_object.py:339-365`)."""
        state = AstroPhot_Model.load(filename)
        self.name = state["name"] if new_name is None else new_name
        if isinstance(state["parameters"], Parameter_Node):
            self.parameters = state["parameters"]
        else:
            self.parameters = Parameter_Node(self.name, state=state["parameters"]).relink()
        for name, sub in state["models"].items():
            sub = dict(sub)
            sub["parameters"] = self.parameters[name]
            if name in self.models:
                self.models[name].load(sub)
            else:
                self.add_model(AstroPhot_Model(name=name, filename=sub, target=self.target))
        self.update_window()
        if "psf_mode" in state:
            self.psf_mode = state["psf_mode"]
        return state

    @torch.no_grad()
    def initialize(self, target=None, parameters=None, **kwargs):
        """Initialise the sub-models in order, each on what the earlier ones leave of the target
        (reference: `group_model_object.py:130-150`)."""
        target = self._init_target(target)
        left = target.copy()
        for model in self.models.values():
            model.initialize(target=left)
            left -= model()

    def fit_mask(self):
        """True where no sub-model has anything to say (reference:
        `group_model_object.py:152-181`)."""
        if isinstance(self.target, Image_List):
            masks = [torch.ones_like(m) for m in self.target[self.window].mask]
            for model in self.models.values():
                sub = model.fit_mask()
                if isinstance(model.target, Image_List):
                    trip = zip(model.target, model.window, sub)
                else:
                    trip = [(model.target, model.window, sub)]
                for tar, win, sm in trip:
                    i = self.target.index(tar)
                    gw = self.window.window_list[i]
                    masks[i][gw.get_self_indices(win)] &= sm[win.get_self_indices(gw)]
            return tuple(masks)
        mask = torch.ones_like(self.target[self.window].mask)
        for model in self.models.values():
            mask[self.window.get_self_indices(model.window)] &= model.fit_mask()[model.window.get_self_indices(self.window)]
        return mask

    def make_model_image(self, window=None):
        window = self.window if window is None else self.window & window
        return self.target[window].model_image()

    def __iter__(self):
        return iter(self.models.values())
