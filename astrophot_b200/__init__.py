"""astrophot_b200 — B200-native forward-model-and-fit hot path with AstroPhot's API.

``import astrophot_b200 as ap`` then use ``ap.image``, ``ap.models``, ``ap.fit``
as with the reference (`astrophot/__init__.py:4`).  All pixel arithmetic runs
in hand-written sm_100a kernels behind the C ABI in ``include/astrophot_b200.h``.
"""
from . import AP_config, errors, param, image, scene, lowering, models, fit, utils  # noqa: F401

__version__ = "0.1.0"
