"""Image containers and window arithmetic (host glue around the hot path).

Keeps the public surface of the reference's ``astrophot/image`` package
(`window_object.py:11-560`, `wcs.py:432-703`, `image_object.py:18-480`,
`target_image.py:17-706`, `psf_image.py:17-158`, `jacobian_image.py:13-175`)
for the constructors and methods the forward-model-and-fit path touches.
Independent implementation: window geometry is plain float64 numpy on the
host (six affine coefficients + an integer shape), pixel data are torch
tensors on ``AP_config.ap_device``.  FITS / astropy-WCS I/O is out of scope
(SURVEY.md §2 row 25).

Conventions (same as the reference): pixel coordinate (i, j) has i along x
(columns) and j along y (rows); ``data[j, i]``; integer coordinates are pixel
centres; plane = S @ (pix - reference_imageij) + reference_imagexy.
"""
from typing import Optional, Sequence

import numpy as np
import torch

from . import AP_config
from .errors import InvalidData, InvalidImage, InvalidWindow, SpecificationConflict, ConflicingWCS

__all__ = [
    "Window", "Window_List", "Image", "Image_List", "Image_Header", "Target_Image",
    "Target_Image_List", "Model_Image", "Model_Image_List", "Jacobian_Image",
    "Jacobian_Image_List", "PSF_Image",
]


def _np(x, n=None):
    if isinstance(x, torch.Tensor):
        x = x.detach().cpu().numpy()
    a = np.array(x, dtype=np.float64)
    if n is not None:
        a = a.reshape(n)
    return a


def _ht(a):
    return torch.as_tensor(np.asarray(a), dtype=torch.float64)


def _dev(x, dtype=None):
    dtype = AP_config.ap_dtype if dtype is None else dtype
    if isinstance(x, torch.Tensor):
        return x.to(dtype=dtype, device=AP_config.ap_device)
    return torch.as_tensor(np.asarray(x), dtype=dtype, device=AP_config.ap_device)


class Window:
    """A parallelogram of pixels pinned to the tangent plane."""

    north = np.pi / 2

    def __init__(self, *, pixel_shape=None, origin=None, center=None, pixelscale=None,
                 reference_imageij=None, reference_imagexy=None, state=None, wcs=None, **kwargs):
        if wcs is not None or "origin_radec" in kwargs or "center_radec" in kwargs:
            raise InvalidWindow("astropy-WCS / RA-DEC placement is out of scope for astrophot_b200; "
                                "use pixelscale= with origin=/center=/reference_imageij=")
        if state is not None:
            self.set_state(state)
            return
        if pixelscale is None:
            AP_config.ap_logger.warning(
                "Assuming pixelscale of 1! To remove this message please provide the pixelscale explicitly")
            pixelscale = 1.0
        if origin is not None and center is not None:
            raise SpecificationConflict(
                "Please provide only one reference position for the window, otherwise the placement is ambiguous")
        self._set_scale(pixelscale)
        self.pixel_shape = pixel_shape
        if origin is not None:
            self._rij = np.array([-0.5, -0.5])
            self._rxy = _np(origin, 2)
        elif center is not None:
            self._rij = self._shape / 2.0 - 0.5
            self._rxy = _np(center, 2)
        else:
            self._rij = np.array([-0.5, -0.5]) if reference_imageij is None else _np(reference_imageij, 2)
            self._rxy = np.zeros(2) if reference_imagexy is None else _np(reference_imagexy, 2)

    # -- raw geometry --------------------------------------------------
    def _set_scale(self, pixelscale):
        S = _np(pixelscale)
        if S.size == 1:
            S = np.eye(2) * float(S.reshape(-1)[0])
        self._S = S.reshape(2, 2).copy()
        self._Sinv = np.linalg.inv(self._S)

    @property
    def pixelscale(self):
        return _ht(self._S)

    @pixelscale.setter
    def pixelscale(self, v):
        self._set_scale(v)

    @property
    def reference_imageij(self):
        return _ht(self._rij)

    @reference_imageij.setter
    def reference_imageij(self, v):
        self._rij = _np(v, 2)

    @property
    def reference_imagexy(self):
        return _ht(self._rxy)

    @reference_imagexy.setter
    def reference_imagexy(self, v):
        self._rxy = _np(v, 2)

    @property
    def pixel_shape(self):
        return torch.as_tensor(self._shape, dtype=torch.int32)

    @pixel_shape.setter
    def pixel_shape(self, shape):
        self._shape = np.round(_np(shape, 2)).astype(np.int64)

    @property
    def pixel_area(self):
        return _ht(abs(self._det()))

    @property
    def pixel_length(self):
        return _ht(np.sqrt(abs(self._det())))

    def _det(self):
        # (written out: LAPACK's LU gives 15.999999999999998 for diag(4, 4))
        S = self._S
        return S[0, 0] * S[1, 1] - S[0, 1] * S[1, 0]

    # -- coordinate maps (numpy core, tensor wrappers) ----------------------
    def _pix2plane(self, p):
        return self._S @ (np.asarray(p, dtype=np.float64).reshape(2, -1) - self._rij[:, None]) + self._rxy[:, None]

    def _plane2pix(self, c):
        return self._Sinv @ (np.asarray(c, dtype=np.float64).reshape(2, -1) - self._rxy[:, None]) + self._rij[:, None]

    @staticmethod
    def _wrap(fn, a, b=None):
        if b is None:
            a = _np(a)
            return _ht(fn(a).reshape(a.shape))
        a, b = _np(a), _np(b)
        out = fn(np.stack((a.reshape(-1), b.reshape(-1))))
        return _ht(out[0].reshape(a.shape)), _ht(out[1].reshape(b.shape))

    def pixel_to_plane(self, pixel_i, pixel_j=None):
        return self._wrap(self._pix2plane, pixel_i, pixel_j)

    def plane_to_pixel(self, plane_x, plane_y=None):
        return self._wrap(self._plane2pix, plane_x, plane_y)

    def pixel_to_plane_delta(self, di, dj=None):
        return self._wrap(lambda p: self._S @ p.reshape(2, -1), di, dj)

    def plane_to_pixel_delta(self, dx, dy=None):
        return self._wrap(lambda p: self._Sinv @ p.reshape(2, -1), dx, dy)

    # -- derived positions --------------------------------------------
    @property
    def _origin(self):
        return self._pix2plane([-0.5, -0.5])[:, 0]

    @property
    def _end(self):
        return self._S @ self._shape.astype(np.float64)

    @property
    def origin(self):
        return _ht(self._origin)

    @property
    def end(self):
        return _ht(self._end)

    @property
    def center(self):
        return _ht(self._origin + self._end / 2)

    @property
    def shape(self):
        sx = self._S @ np.array([float(self._shape[0]), 0.0])
        sy = self._S @ np.array([0.0, float(self._shape[1])])
        return _ht([np.linalg.norm(sx), np.linalg.norm(sy)])

    @property
    def size(self):
        return int(np.prod(self._shape))

    # -- copies / rescaling ------------------------------------------
    def copy(self, **kwargs):
        if "origin" in kwargs or "center" in kwargs:
            kw = {"pixelscale": self._S, "pixel_shape": self._shape}
            kw.update(kwargs)
            return Window(**kw)
        kw = {"pixelscale": self._S, "pixel_shape": self._shape,
              "reference_imageij": self._rij, "reference_imagexy": self._rxy}
        kw.update(kwargs)
        return Window(**kw)

    def rescale_pixel(self, scale, **kwargs):
        scale = float(scale)
        return self.copy(pixelscale=self._S * scale, pixel_shape=np.floor(self._shape / scale),
                         reference_imageij=(self._rij + 0.5) / scale - 0.5, **kwargs)

    def shift(self, shift):
        self._rxy = self._rxy + _np(shift, 2)
        return self

    def pixel_shift(self, shift):
        self._rij = self._rij - _np(shift, 2)
        return self

    def _edges(self, pixels):
        p = _np(pixels).reshape(-1)
        if p.size == 1:
            return np.array([p[0], p[0], p[0], p[0]])
        if p.size == 2:
            return np.array([p[0], p[1], p[0], p[1]])
        if p.size == 4:
            return p
        raise ValueError(f"Unrecognized pixel crop format: {pixels}")

    def crop_pixel(self, pixels):
        e = self._edges(pixels)
        self.pixel_shape = self._shape - e[:2] - e[2:]
        self._rij = self._rij - e[:2]
        return self

    def pad_pixel(self, pixels):
        e = self._edges(pixels)
        self.pixel_shape = self._shape + e[:2] + e[2:]
        self._rij = self._rij + e[:2]
        return self

    def crop_to_pixel(self, pixels):
        """``[[xmin, xmax], [ymin, ymax]]`` in this window's pixel indices."""
        p = _np(pixels).reshape(2, 2)
        self._rij = self._rij - p[:, 0]
        self.pixel_shape = p[:, 1] - p[:, 0]
        return self

    # -- index arithmetic ----------------------------------------------
    @staticmethod
    def _get_indices(ref_window, obj_window):
        """Slices of ``ref_window``'s pixel grid covered by ``obj_window``
        (reference: `window_object.py:241-253`; round-half-even like torch)."""
        lo = np.round(ref_window._plane2pix(obj_window._origin)[:, 0] + 0.5).astype(np.int64)
        hi = np.round(ref_window._plane2pix(obj_window._origin + obj_window._end)[:, 0] + 0.5).astype(np.int64)
        lo = np.maximum(0, lo)
        hi = np.minimum(ref_window._shape, hi)
        return slice(int(lo[1]), int(hi[1])), slice(int(lo[0]), int(hi[0]))

    def get_self_indices(self, obj):
        return self._get_indices(self, obj if isinstance(obj, Window) else obj.window)

    def get_other_indices(self, obj):
        return self._get_indices(obj if isinstance(obj, Window) else obj.window, self)

    def overlap_frac(self, other):
        ov = self & other
        a = float(torch.prod(ov.shape)) if np.all(ov._shape > 0) else 0.0
        full = float(torch.prod(self.shape)) + float(torch.prod(other.shape)) - a
        return _ht(a / full)

    def _span(self, other):
        lo = self._plane2pix(other._origin)[:, 0]
        hi = self._plane2pix(other._origin + other._end)[:, 0]
        return lo, hi

    def __or__(self, other):
        lo, hi = self._span(other)
        lo = np.minimum(-0.5, lo)
        hi = np.maximum(self._shape.astype(np.float64), hi)
        return self.copy(origin=self._pix2plane(lo)[:, 0], pixel_shape=hi - lo)

    def __ior__(self, other):
        lo, hi = self._span(other)
        lo = np.minimum(-0.5, lo)
        hi = np.maximum(self._shape.astype(np.float64), hi)
        self._rij = self._rij - (lo + 0.5)
        self.pixel_shape = hi - lo
        return self

    def __and__(self, other):
        lo, hi = self._span(other)
        lo = np.maximum(-0.5, lo)
        hi = np.minimum(self._shape.astype(np.float64) - 0.5, hi)
        return self.copy(origin=self._pix2plane(lo)[:, 0], pixel_shape=np.maximum(hi - lo, 0))

    def __iand__(self, other):
        lo, hi = self._span(other)
        lo = np.maximum(-0.5, lo)
        hi = np.minimum(self._shape.astype(np.float64), hi)
        self._rij = self._rij - (lo + 0.5)
        self.pixel_shape = np.maximum(hi - lo, 0)
        return self

    def __eq__(self, other):
        return (isinstance(other, Window) and np.all(self._shape == other._shape)
                and np.all(self._S == other._S)
                and np.all(self._pix2plane([0, 0]) == other._pix2plane([0, 0])))

    def __ne__(self, other):
        return not self == other

    __hash__ = object.__hash__

    # -- coordinate grids (API parity; the kernels never materialise these) --
    def _grid(self, xs, ys):
        mx, my = np.meshgrid(xs, ys, indexing="xy")
        c = self._pix2plane(np.stack((mx.reshape(-1), my.reshape(-1))))
        return _dev(c.reshape(2, *mx.shape))

    def get_coordinate_meshgrid(self):
        return self._grid(np.arange(self._shape[0], dtype=np.float64), np.arange(self._shape[1], dtype=np.float64))

    def get_coordinate_corner_meshgrid(self):
        return self._grid(np.arange(self._shape[0] + 1) - 0.5, np.arange(self._shape[1] + 1) - 0.5)

    def get_coordinate_simps_meshgrid(self):
        return self._grid(0.5 * np.arange(2 * self._shape[0] + 1) - 0.5, 0.5 * np.arange(2 * self._shape[1] + 1) - 0.5)

    # -- state ------------------------------------------------------------
    def get_state(self):
        return {"pixelscale": self._S.tolist(), "reference_imageij": self._rij.tolist(),
                "reference_imagexy": self._rxy.tolist(), "pixel_shape": self._shape.tolist()}

    def set_state(self, state):
        self._set_scale(state.get("pixelscale", 1.0))
        self._rij = _np(state.get("reference_imageij", (-0.5, -0.5)), 2)
        self._rxy = _np(state.get("reference_imagexy", (0.0, 0.0)), 2)
        self.pixel_shape = state["pixel_shape"]

    def to(self, dtype=None, device=None):
        return self

    def __str__(self):
        return (f"window origin: {self._origin.tolist()}, shape: {self.shape.tolist()}, "
                f"center: {self.center.tolist()}, pixelscale: {self._S.tolist()}")

    def __repr__(self):
        return f"window pixel_shape: {self._shape.tolist()}, shape: {self.shape.tolist()}\n{self.get_state()}"


class Window_List(Window):
    """One window per image of an ``Image_List`` (reference:
    `window_object.py:563-700`)."""

    def __init__(self, window_list=None, state=None):
        if state is not None:
            self.window_list = [Window(state=s) for s in state["window_list"]]
        else:
            self.window_list = list(window_list or [])

    @property
    def origin(self):
        return tuple(w.origin for w in self.window_list)

    @property
    def shape(self):
        return tuple(w.shape for w in self.window_list)

    @property
    def center(self):
        return tuple(w.center for w in self.window_list)

    @property
    def pixel_shape(self):
        return tuple(w.pixel_shape for w in self.window_list)

    def copy(self):
        return Window_List([w.copy() for w in self.window_list])

    def shift(self, shift):
        raise NotImplementedError("Window_List cannot be shifted as a whole")

    def get_state(self):
        return {"window_list": [w.get_state() for w in self.window_list]}

    def _zip(self, other, op):
        return Window_List([op(a, b) for a, b in zip(self.window_list, other.window_list)])

    def __or__(self, other):
        return self._zip(other, lambda a, b: a | b)

    def __and__(self, other):
        return self._zip(other, lambda a, b: a & b)

    def __ior__(self, other):
        for a, b in zip(self.window_list, other.window_list):
            a |= b
        return self

    def __iand__(self, other):
        for a, b in zip(self.window_list, other.window_list):
            a &= b
        return self

    def __eq__(self, other):
        return isinstance(other, Window_List) and all(a == b for a, b in zip(self.window_list, other.window_list))

    __hash__ = object.__hash__

    def __len__(self):
        return len(self.window_list)

    def __iter__(self):
        return iter(self.window_list)

    def __str__(self):
        return "Window List: \n" + "\n".join(str(w) for w in self.window_list)

    __repr__ = __str__


class Image:
    """Pixel data + window.  ``image.header`` is the image itself (the
    reference splits this into ``Image_Header``; the attributes are the same)."""

    def __init__(self, *, data=None, header=None, pixelscale=None, window: Optional[Window] = None,
                 zeropoint=None, metadata=None, origin=None, center=None, identity=None,
                 note=None, wcs=None, filename=None, state=None, **kwargs):
        if wcs is not None or filename is not None or state is not None:
            raise InvalidData("FITS / WCS / saved-state image construction is out of scope for astrophot_b200")
        if header is not None:
            window = header.window
            zeropoint = header.zeropoint if zeropoint is None else zeropoint
            metadata = header.metadata if metadata is None else metadata
            identity = header.identity if identity is None else identity
        if data is None and window is None:
            raise InvalidData("Image must have either data or a window to construct itself.")
        self.identity = str(id(self)) if identity is None else identity
        self.zeropoint = None if zeropoint is None else float(_np(zeropoint).reshape(-1)[0])
        self.metadata = metadata
        self.note = note
        self._data = None
        if window is None:
            shp = tuple(data.shape)
            window = Window(pixel_shape=(shp[1], shp[0]), pixelscale=pixelscale, origin=origin, center=center,
                            **{k: kwargs[k] for k in ("reference_imageij", "reference_imagexy") if k in kwargs})
        self.window = window
        if data is None:
            self._data = torch.zeros((int(window._shape[1]), int(window._shape[0])),
                                     dtype=AP_config.ap_dtype, device=AP_config.ap_device)
        else:
            self.set_data(data)

    # -- header facade ---------------------------------------------------------
    @property
    def header(self):
        return self

    @property
    def north(self):
        return Window.north

    @property
    def pixelscale(self):
        return self.window.pixelscale

    @property
    def pixel_area(self):
        return self.window.pixel_area

    @property
    def pixel_length(self):
        return self.window.pixel_length

    @property
    def origin(self):
        return self.window.origin

    @property
    def shape(self):
        return self.window.shape

    @property
    def center(self):
        return self.window.center

    @property
    def size(self):
        return self.window.size

    def pixel_to_plane(self, *a):
        return self.window.pixel_to_plane(*a)

    def plane_to_pixel(self, *a):
        return self.window.plane_to_pixel(*a)

    def pixel_to_plane_delta(self, *a):
        return self.window.pixel_to_plane_delta(*a)

    def plane_to_pixel_delta(self, *a):
        return self.window.plane_to_pixel_delta(*a)

    def get_coordinate_meshgrid(self):
        return self.window.get_coordinate_meshgrid()

    def get_coordinate_corner_meshgrid(self):
        return self.window.get_coordinate_corner_meshgrid()

    def get_coordinate_simps_meshgrid(self):
        return self.window.get_coordinate_simps_meshgrid()

    def pixel_shift(self, shift):
        self.window.pixel_shift(shift)

    def shift(self, shift):
        self.window.shift(shift)

    # -- data ------------------------------------------------------------
    @property
    def data(self):
        return self._data

    @data.setter
    def data(self, data):
        self.set_data(data)

    def set_data(self, data, require_shape=True):
        if self._data is not None and require_shape and tuple(data.shape) != tuple(self._data.shape):
            raise SpecificationConflict(
                f"Attempting to change image data with tensor that has a different shape! ({tuple(self._data.shape)} vs {tuple(data.shape)})")
        self._data = _dev(data)

    def _extra(self):
        return {}

    def copy(self, **kwargs):
        kw = dict(data=self._data.clone(), window=self.window.copy(), zeropoint=self.zeropoint,
                  metadata=self.metadata, identity=self.identity)
        kw.update(self._extra())
        kw.update(kwargs)
        return self.__class__(**kw)

    def blank_copy(self, **kwargs):
        kw = dict(data=torch.zeros_like(self._data), window=self.window.copy(), zeropoint=self.zeropoint,
                  metadata=self.metadata, identity=self.identity)
        kw.update(kwargs)
        return self.__class__(**kw)

    def get_window(self, window, **kwargs):
        """View of the pixels this image shares with ``window``."""
        rows, cols = self.window.get_self_indices(window)
        kw = dict(data=self._data[rows, cols], window=self.window & window, zeropoint=self.zeropoint,
                  metadata=self.metadata, identity=self.identity)
        kw.update(kwargs)
        return self.__class__(**kw)

    def __getitem__(self, item):
        if isinstance(item, Image):
            item = item.window
        if isinstance(item, Window):
            return self.get_window(item)
        raise ValueError("Unrecognized Image getitem request!")

    def to(self, dtype=None, device=None):
        self._data = self._data.to(dtype=dtype or AP_config.ap_dtype, device=device or AP_config.ap_device)
        return self

    def crop(self, pixels):
        e = self.window._edges(pixels).astype(np.int64)
        h, w = self._data.shape[:2]
        self._data = self._data[e[1] : h - e[3], e[0] : w - e[2]]
        self.window.crop_pixel(pixels)
        return self

    def flatten(self, attribute="data"):
        return getattr(self, attribute).reshape(-1)

    def reduce(self, scale, **kwargs):
        """Sum ``scale`` x ``scale`` pixel blocks (reference: `image_object.py:348-376`)."""
        if isinstance(scale, torch.Tensor) and scale.dtype in (torch.int32, torch.int64):
            scale = int(scale)
        if not isinstance(scale, (int, np.integer)) or isinstance(scale, bool):
            raise SpecificationConflict(f"Reduce scale must be an integer! not {type(scale)}")
        scale = int(scale)
        if scale == 1:
            return self
        MS, NS = self._data.shape[0] // scale, self._data.shape[1] // scale
        data = self._data[: MS * scale, : NS * scale].reshape(MS, scale, NS, scale).sum(dim=(1, 3))
        kw = dict(data=data, window=self.window.rescale_pixel(scale), zeropoint=self.zeropoint,
                  metadata=self.metadata, identity=self.identity)
        kw.update(kwargs)
        return self.__class__(**kw)

    # -- arithmetic through windows -----------------------------------------
    def _binary(self, other, sign, inplace):
        tgt = self if inplace else self.copy()
        if isinstance(other, Image):
            if np.any(np.abs(self.window._S - other.window._S) > 1e-12 * np.abs(self.window._S).max()):
                raise ConflicingWCS("images have different pixelscale, cannot add/subtract")
            mine = tgt.window.get_self_indices(other)
            theirs = other.window.get_self_indices(tgt)
            tgt._data[mine] += sign * other._data[theirs]
        else:
            tgt._data += sign * other
        return tgt

    def __iadd__(self, other):
        return self._binary(other, 1, True)

    def __isub__(self, other):
        return self._binary(other, -1, True)

    def __add__(self, other):
        return self._binary(other, 1, False)

    def __sub__(self, other):
        return self._binary(other, -1, False)

    def __str__(self):
        return f"image pixelscale: {self.window._S.tolist()} origin: {self.window._origin.tolist()}\ndata: {self._data}"

    __repr__ = __str__


Image_Header = Image  # the header facade is the image


class Model_Image(Image):
    """Where sampled model flux accumulates (reference: `model_image.py`)."""

    def __init__(self, *, target_identity=None, **kwargs):
        super().__init__(**kwargs)
        self.target_identity = target_identity

    def _extra(self):
        return {"target_identity": self.target_identity}

    def clear_image(self):
        self._data.zero_()

    def blank_copy(self, **kwargs):
        return super().blank_copy(target_identity=self.target_identity, **kwargs)

    def get_window(self, window, **kwargs):
        return super().get_window(window, target_identity=self.target_identity, **kwargs)

    def reduce(self, scale, **kwargs):
        return super().reduce(scale, target_identity=self.target_identity, **kwargs)

    def replace(self, other, data=None):
        """Overwrite the overlap with another image, a window's worth of ``data``, or everything (reference:
        `model_image.py:51-63`)."""
        if isinstance(other, Image):
            mine = self.window.get_self_indices(other)
            theirs = other.window.get_self_indices(self)
            if self._data[mine].numel() == 0 or other._data[theirs].numel() == 0:
                return
            self._data[mine] = other._data[theirs]
        elif isinstance(other, Window):
            self._data[self.window.get_self_indices(other)] = torch.as_tensor(data, dtype=self._data.dtype,
                                                                              device=self._data.device)
        else:
            self.data = other


class Jacobian_Image(Image):
    """(H, W, P) derivative stack with one parameter identity per column
    (reference: `jacobian_image.py:13-130`).  Only ever materialised for small
    problems and tests; the LM path never builds it."""

    def __init__(self, *, parameters, target_identity=None, **kwargs):
        super().__init__(**kwargs)
        self.target_identity = target_identity
        self.parameters = list(parameters)
        if len(set(self.parameters)) != len(self.parameters):
            raise SpecificationConflict("Every parameter should be unique upon jacobian creation")

    def _extra(self):
        return {"parameters": list(self.parameters), "target_identity": self.target_identity}

    def set_data(self, data, require_shape=True):
        self._data = _dev(data)

    def flatten(self, attribute="data"):
        return getattr(self, attribute).reshape(-1, len(self.parameters))

    def get_window(self, window, **kwargs):
        return super().get_window(window, parameters=self.parameters, target_identity=self.target_identity, **kwargs)

    def __iadd__(self, other):
        if not isinstance(other, Jacobian_Image):
            raise InvalidImage("Jacobian images can only add with each other, not: type(other)")
        if other._data is None or other._data.numel() == 0 or len(other.parameters) == 0:
            return self
        mine = self.window.get_self_indices(other)
        theirs = other.window.get_self_indices(self)
        for k, ident in enumerate(other.parameters):
            if ident in self.parameters:
                col = self.parameters.index(ident)
            else:
                col = len(self.parameters)
                self.parameters.append(ident)
                pad = torch.zeros(*self._data.shape[:2], 1, dtype=self._data.dtype, device=self._data.device)
                self._data = torch.cat((self._data, pad), dim=2)
            self._data[mine[0], mine[1], col] += other._data[theirs[0], theirs[1], k]
        return self


class PSF_Image(Image):
    """Odd-sized PSF stamp centred on (0, 0) (reference: `psf_image.py:17-93`)."""

    has_mask = False
    has_variance = False

    def __init__(self, *, psf_upscale=1, **kwargs):
        super().__init__(**kwargs)
        self.psf_upscale = psf_upscale
        h, w = self._data.shape
        self.window._rij = np.array([(w - 1) / 2.0, (h - 1) / 2.0])
        self.window._rxy = np.zeros(2)

    def set_data(self, data, require_shape=True):
        super().set_data(data, require_shape)
        if any(s % 2 != 1 for s in self._data.shape):
            raise SpecificationConflict(f"psf must have odd shape, not {tuple(self._data.shape)}")
        if bool(torch.any(self._data < 0)):
            AP_config.ap_logger.warning("psf data should be non-negative")

    def normalize(self):
        self._data /= torch.sum(self._data)

    @property
    def mask(self):
        return torch.zeros_like(self._data, dtype=torch.bool)

    @property
    def psf_border_int(self):
        h, w = self._data.shape
        return torch.tensor([int(np.ceil((1 + w) / 2)), int(np.ceil((1 + h) / 2))], dtype=torch.int32)

    @property
    def psf_border(self):
        return self.window.pixel_to_plane_delta(self.psf_border_int.to(torch.float64))

    def _extra(self):
        return {"psf_upscale": self.psf_upscale}

    def model_image(self, data=None, **kwargs):
        return Model_Image(data=torch.zeros_like(self._data) if data is None else data,
                           window=self.window, zeropoint=self.zeropoint, target_identity=self.identity, **kwargs)

    def jacobian_image(self, parameters=None, data=None, **kwargs):
        if parameters is None:
            parameters, data = [], torch.zeros(*self._data.shape, 0)
        elif data is None:
            data = torch.zeros(*self._data.shape, len(parameters), dtype=AP_config.ap_dtype, device=AP_config.ap_device)
        return Jacobian_Image(parameters=parameters, target_identity=self.identity, data=data,
                              window=self.window, zeropoint=self.zeropoint, **kwargs)

    def blank_copy(self, **kwargs):
        return super().blank_copy(psf_upscale=self.psf_upscale, **kwargs)

    def expand(self, padding):
        raise NotImplementedError("expand not available for PSF_Image")


class Target_Image(Image):
    """Data to fit: pixels + optional variance/weight, mask and PSF
    (reference: `target_image.py:17-532`)."""

    image_count = 0

    def __init__(self, *args, mask=None, variance=None, weight=None, psf=None, psf_upscale=1, **kwargs):
        super().__init__(*args, **kwargs)
        self._weight = self._mask = self._psf = None
        self.set_mask(mask)
        if weight is not None:
            self.set_weight(weight)
        elif variance is not None:
            self.set_variance(variance)
        self.set_psf(psf, psf_upscale)
        if bool(torch.any(torch.isnan(self._data))):
            self.set_mask(torch.logical_or(self.mask, torch.isnan(self._data)))

    @property
    def has_variance(self):
        return self._weight is not None

    has_weight = has_variance

    @property
    def has_mask(self):
        return self._mask is not None

    @property
    def has_psf(self):
        return self._psf is not None

    @property
    def weight(self):
        return self._weight if self.has_weight else torch.ones_like(self._data)

    @weight.setter
    def weight(self, w):
        self.set_weight(w)

    @property
    def variance(self):
        return 1.0 / self._weight if self.has_variance else torch.ones_like(self._data)

    @variance.setter
    def variance(self, v):
        self.set_variance(v)

    @property
    def standard_deviation(self):
        return torch.sqrt(self.variance)

    @property
    def mask(self):
        return self._mask if self.has_mask else torch.zeros_like(self._data, dtype=torch.bool)

    @mask.setter
    def mask(self, m):
        self.set_mask(m)

    @property
    def psf(self):
        if self.has_psf:
            return self._psf
        raise AttributeError("This image does not have a PSF")

    @psf.setter
    def psf(self, psf):
        self.set_psf(psf)

    def set_variance(self, variance):
        if variance is None:
            self._weight = None
            return
        if isinstance(variance, str) and variance == "auto":
            self.set_weight("auto")
            return
        self.set_weight(1.0 / _dev(variance))

    def set_weight(self, weight):
        if weight is None:
            self._weight = None
            return
        if isinstance(weight, str) and weight == "auto":
            # estimated from the image itself (reference: `target_image.py:291-292`, `utils/initialize/variance.py`)
            from .utils.initialize import auto_variance
            weight = torch.as_tensor(1 / auto_variance(self._data, self._mask))
        if tuple(weight.shape) != tuple(self._data.shape):
            raise SpecificationConflict(
                f"weight/variance must have same shape as data ({tuple(weight.shape)} vs {tuple(self._data.shape)})")
        self._weight = _dev(weight)

    def set_mask(self, mask):
        if mask is None:
            self._mask = None
            return
        if tuple(mask.shape) != tuple(self._data.shape):
            raise SpecificationConflict(
                f"mask must have same shape as data ({tuple(mask.shape)} vs {tuple(self._data.shape)})")
        self._mask = _dev(mask, torch.bool)

    def set_psf(self, psf, psf_upscale=1):
        from .models import AstroPhot_Model  # local: models imports image

        if psf is None:
            self._psf = None
        elif isinstance(psf, (PSF_Image, AstroPhot_Model)):
            self._psf = psf
        else:
            self._psf = PSF_Image(data=psf, psf_upscale=psf_upscale,
                                  pixelscale=self.window._S / psf_upscale, identity=self.identity)

    def or_mask(self, mask):
        self._mask = torch.logical_or(self.mask, _dev(mask, torch.bool))

    def and_mask(self, mask):
        self._mask = torch.logical_and(self.mask, _dev(mask, torch.bool))

    def _extra(self):
        return {"mask": self._mask, "weight": self._weight, "psf": self._psf}

    def blank_copy(self, **kwargs):
        return super().blank_copy(mask=self._mask, psf=self._psf, **kwargs)

    def get_window(self, window, **kwargs):
        rows, cols = self.window.get_self_indices(window)
        return super().get_window(
            window,
            weight=self._weight[rows, cols] if self.has_weight else None,
            mask=self._mask[rows, cols] if self.has_mask else None,
            psf=self._psf, **kwargs)

    def model_image(self, data=None, **kwargs):
        return Model_Image(data=torch.zeros_like(self._data) if data is None else data,
                           window=self.window, zeropoint=self.zeropoint,
                           target_identity=self.identity, **kwargs)

    def jacobian_image(self, parameters=None, data=None, **kwargs):
        if parameters is None:
            parameters = []
            data = torch.zeros(*self._data.shape, 0, dtype=AP_config.ap_dtype, device=AP_config.ap_device)
        elif data is None:
            data = torch.zeros(*self._data.shape, len(parameters), dtype=AP_config.ap_dtype, device=AP_config.ap_device)
        return Jacobian_Image(parameters=parameters, target_identity=self.identity, data=data,
                              window=self.window, zeropoint=self.zeropoint, **kwargs)

    def reduce(self, scale, **kwargs):
        """Lower resolution copy: data and variance summed over ``scale`` x ``scale`` blocks, a block masked when any
        of its pixels is, the PSF reduced alike (reference: `target_image.py:443-478`)."""
        if scale == 1:
            return self
        MS, NS = self._data.shape[0] // scale, self._data.shape[1] // scale

        def blocks(t):
            return t[: MS * scale, : NS * scale].reshape(MS, scale, NS, scale)

        return super().reduce(
            scale,
            variance=blocks(self.variance).sum(dim=(1, 3)) if self.has_variance else None,
            mask=blocks(self.mask).amax(dim=(1, 3)) if self.has_mask else None,
            psf=self.psf.reduce(scale) if self.has_psf else None,
            **kwargs)


class Image_List(Image):
    """Ordered collection of images treated as one data set (multi-band)."""

    def __init__(self, image_list, window=None):
        self.image_list = list(image_list)
        if len(set(im.identity for im in self.image_list)) != len(self.image_list):
            raise InvalidImage("Images in an Image_List must have unique identities")

    @property
    def window(self):
        return Window_List([im.window for im in self.image_list])

    @property
    def pixelscale(self):
        return tuple(im.pixelscale for im in self.image_list)

    @property
    def zeropoint(self):
        return tuple(im.zeropoint for im in self.image_list)

    @property
    def data(self):
        return tuple(im.data for im in self.image_list)

    @data.setter
    def data(self, data):
        for im, d in zip(self.image_list, data):
            im.data = d

    def copy(self):
        return self.__class__([im.copy() for im in self.image_list])

    def blank_copy(self):
        return self.__class__([im.blank_copy() for im in self.image_list])

    def get_window(self, window):
        return self.__class__([im.get_window(w) for im, w in zip(self.image_list, window)])

    def index(self, other):
        """Position of ``other`` (or, for a model / Jacobian image, of the target it was made for; reference:
        `target_image.py:602-630`)."""
        keys = {other.identity, getattr(other, "target_identity", None)} - {None}
        for i, im in enumerate(self.image_list):
            if im.identity in keys:
                return i
        raise ValueError("Could not find identity match between image list and input image")

    def flatten(self, attribute="data"):
        return torch.cat([im.flatten(attribute) for im in self.image_list])

    def _each(self, other, op):
        if isinstance(other, Image_List):
            try:
                pairs = [(self.image_list[self.index(o)], o) for o in other.image_list]
            except ValueError:
                # unrelated images: element by element, like the reference's base list (image_object.py:619-638)
                pairs = list(zip(self.image_list, other.image_list))
            for a, b in pairs:
                op(a, b)
        elif isinstance(other, Image):
            op(self.image_list[self.index(other)], other)
        else:
            for im, o in zip(self.image_list, other):
                op(im, o)
        return self

    def __iadd__(self, other):
        return self._each(other, lambda a, b: a.__iadd__(b))

    def __isub__(self, other):
        return self._each(other, lambda a, b: a.__isub__(b))

    def __add__(self, other):
        return self.copy().__iadd__(other)

    def __sub__(self, other):
        return self.copy().__isub__(other)

    def __getitem__(self, item):
        if isinstance(item, Window_List):
            return self.get_window(item)
        if isinstance(item, Image_List):
            return self.get_window(item.window)
        if isinstance(item, int):
            return self.image_list[item]
        raise ValueError("Unrecognized Image_List getitem request!")

    def __iter__(self):
        return iter(self.image_list)

    def __len__(self):
        return len(self.image_list)

    def __str__(self):
        return f"image list of:\n" + "\n".join(str(im) for im in self.image_list)

    __repr__ = __str__


class Model_Image_List(Image_List, Model_Image):
    def __init__(self, image_list, window=None):
        Image_List.__init__(self, image_list)
        if not all(isinstance(im, Model_Image) for im in self.image_list):
            raise InvalidImage("Model_Image_List can only hold Model_Image objects")

    def clear_image(self):
        for im in self.image_list:
            im.clear_image()

    def replace(self, other, data=None):
        """Element by element (reference: `model_image.py:110-116`)."""
        for k, (im, oth) in enumerate(zip(self.image_list, other)):
            im.replace(oth, None if data is None else data[k])

    @property
    def target_identity(self):
        ids = tuple(im.target_identity for im in self.image_list)
        return None if any(i is None for i in ids) else ids

    def index(self, other):
        key = getattr(other, "target_identity", None) or other.identity
        for i, im in enumerate(self.image_list):
            if key in (im.target_identity, im.identity):
                return i
        raise ValueError("Could not find identity match between image list and input image")


class Jacobian_Image_List(Image_List, Jacobian_Image):
    def __init__(self, image_list, window=None):
        Image_List.__init__(self, image_list)

    def flatten(self, attribute="data"):
        if len(set(tuple(im.parameters) for im in self.image_list)) > 1:
            raise SpecificationConflict("Jacobian image list sub-images track different parameters")
        return torch.cat([im.flatten(attribute) for im in self.image_list])

    @property
    def parameters(self):
        return self.image_list[0].parameters

    def index(self, other):
        key = getattr(other, "target_identity", None) or other.identity
        for i, im in enumerate(self.image_list):
            if key in (im.target_identity, im.identity):
                return i
        raise ValueError("Could not find identity match between image list and input image")


class Target_Image_List(Image_List, Target_Image):
    """Several ``Target_Image`` fitted jointly (reference: `target_image.py:535-706`)."""

    def __init__(self, image_list, window=None):
        Image_List.__init__(self, image_list)
        if not all(isinstance(im, Target_Image) for im in self.image_list):
            raise InvalidImage("Target_Image_List can only hold Target_Image objects")

    @property
    def variance(self):
        return tuple(im.variance for im in self.image_list)

    @property
    def weight(self):
        return tuple(im.weight for im in self.image_list)

    @property
    def has_variance(self):
        return any(im.has_variance for im in self.image_list)

    has_weight = has_variance

    @property
    def mask(self):
        return tuple(im.mask for im in self.image_list)

    @property
    def has_mask(self):
        return any(im.has_mask for im in self.image_list)

    @property
    def psf(self):
        return tuple(im.psf for im in self.image_list)

    @property
    def has_psf(self):
        return any(im.has_psf for im in self.image_list)

    def model_image(self, data=None):
        if data is None:
            data = [None] * len(self.image_list)
        return Model_Image_List([im.model_image(data=d) for im, d in zip(self.image_list, data)])

    def jacobian_image(self, parameters=None, data=None):
        if data is None:
            data = [None] * len(self.image_list)
        return Jacobian_Image_List([im.jacobian_image(parameters, d) for im, d in zip(self.image_list, data)])

    def match_indices(self, other):
        if isinstance(other, Image_List):
            return [self.index(o) for o in other.image_list]
        return self.index(other)
