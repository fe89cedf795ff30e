"""Flat "scene" tables: what a lowered model tree looks like at the C-ABI.

These dataclasses are the Python mirror of the structs in
``include/astrophot_b200.h`` (``apb_image_t``, ``apb_source_t``, ``apb_psf_t``,
``apb_param_t``).  ``lowering.py`` builds them from the model/parameter/image
objects; ``cabi.py`` packs them into ctypes structs for the sm_100a library;
the CPU oracle under ``oracle/`` (test infrastructure only) consumes the very
same tables, which is what makes oracle-vs-CUDA parity tests meaningful.
"""
from dataclasses import dataclass, field
from typing import List, Optional

import numpy as np

# -- enumerations (values are part of the C ABI) -------------------------------
KIND_SERSIC, KIND_EXPONENTIAL, KIND_GAUSSIAN, KIND_MOFFAT, KIND_SPLINE, KIND_POINT, KIND_FLAT_SKY, KIND_PLANE_SKY = range(8)
KIND_NAMES = ["sersic", "exponential", "gaussian", "moffat", "spline", "point", "flat_sky", "plane_sky"]

FLAG_RADIAL = 1        # no rotation / axis ratio (PSF-model kinds): elements skip q, PA
FLAG_NORMALIZE = 2     # divide the sampled stamp by its sum (PSF models)
FLAG_AMP = 4           # the LAST element is a log10 amplitude applied after sampling / normalisation, never seen by the
                       # profile: a point source drawn from a PSF *model* (point_source.py:122-140) is that model's
                       # profile centred on the point source, times 10^flux

TR_NONE, TR_LOWER, TR_UPPER, TR_BOTH, TR_CYCLIC = range(5)

SAMPLE_MIDPOINT, SAMPLE_SIMPSONS, SAMPLE_QUAD, SAMPLE_TRAPEZOID = range(4)
INTEGRATE_NONE, INTEGRATE_THRESHOLD = range(2)
REF_MEAN, REF_SERSIC_FLUX = range(2)
SHIFT_NONE, SHIFT_BILINEAR = 0, 1   # SHIFT_LANCZOS + order = 10 + order
SHIFT_LANCZOS = 10
CONV_AUTO, CONV_DIRECT, CONV_FFT = range(3)   # psf_convolve_mode: "fft" -> AUTO (fastest), "direct" -> DIRECT

MAX_ELEM = 24
MAX_PROF = 20

# number of leading elements of each kind (before any spline nodes)
ELEMS = {
    KIND_SERSIC: ("cx", "cy", "q", "PA", "n", "Re", "Ie"),
    KIND_EXPONENTIAL: ("cx", "cy", "q", "PA", "Re", "Ie"),
    KIND_GAUSSIAN: ("cx", "cy", "q", "PA", "sigma", "flux"),
    KIND_MOFFAT: ("cx", "cy", "q", "PA", "n", "Rd", "I0"),
    KIND_SPLINE: ("cx", "cy", "q", "PA"),
    KIND_POINT: ("cx", "cy", "flux"),
    KIND_FLAT_SKY: ("cx", "cy", "F"),
    KIND_PLANE_SKY: ("cx", "cy", "F", "dx", "dy"),
}


@dataclass
class SceneImage:
    """One target image region (the fit window of one band)."""
    H: int
    W: int
    S: np.ndarray            # (2,2) pixelscale
    rij: np.ndarray          # (2,) reference pixel
    rxy: np.ndarray          # (2,) reference plane position
    data: object = None      # (H,W) tensor/array or None
    weight: object = None    # (H,W) or None (= ones)
    mask: object = None      # (H,W) bool/uint8, True = ignore; or None
    aux: bool = False        # grid of an auxiliary PSF model (a PSF_Image): sampled, never an output or a chi^2 term;
                             # aux images come after the target images


@dataclass
class ScenePSF:
    data: object             # (h,w) odd-shaped stamp, un-normalised; None when produced by a source
    source: int = -1         # index of the PSF-model source (on an aux image) that produces the stamp on every
                             # sampling pass (model_object.py:133-147,307-310), or -1
    shape: tuple = None      # (h, w) when data is None


@dataclass
class SceneSource:
    kind: int
    image: int
    out: tuple               # (x0, y0, w, h) output window, image pixel indices
    fwd: tuple               # working window when sampled inside the forward model
    jac: tuple               # working window when differentiated
    slot: List[int]          # per element: index into x, or -1 = locked
    cval: List[float]        # per element: value when locked (natural units)
    flags: int = 0
    prof: List[float] = field(default_factory=list)   # spline node radii
    sampling_mode: int = SAMPLE_MIDPOINT
    quad_init: int = 3       # N of "quad:N"
    integrate_mode: int = INTEGRATE_THRESHOLD
    quad_level: int = 3
    gridding: int = 5
    max_depth: int = 3
    tolerance: float = 1e-2
    softening: float = 1e-3
    ref_mode: int = REF_MEAN
    psf: int = -1
    psf_shift: int = SHIFT_BILINEAR
    conv_mode: int = CONV_AUTO
    name: str = ""
    owner: int = -1          # index into Scene.owners (the model this source is a tile-clipped piece of)
    mask: object = None      # the model's own mask (model_object.py:370-371): 2-D bool array, True = the model contributes
                             # nothing there (value and derivatives); element [0, 0] is image pixel `mask_origin`
    mask_origin: tuple = (0, 0)
    upscale: int = 1         # super-sampled PSF (model_object.py:312-315,348-349): the source is sampled and convolved on
                             # pixels 1 / upscale of the image's and block-summed back; windows stay in image pixels

    @property
    def n_elem(self):
        return len(self.slot)


@dataclass
class Scene:
    images: List[SceneImage]
    sources: List[SceneSource]
    psfs: List[ScenePSF]
    transform: np.ndarray    # (P,) int32
    lo: np.ndarray           # (P,) float64 (nan when absent)
    hi: np.ndarray           # (P,)
    identities: Optional[list] = None   # (P,) parameter identity strings (host only)
    # models of the whole fit when the scene holds tile-clipped pieces (lowering.tile_scene):
    # (uncut image index, (x0, y0, w, h) on it, [free parameter slots]) -- identical on every rank
    owners: Optional[list] = None

    @property
    def n_par(self):
        return int(len(self.transform))
