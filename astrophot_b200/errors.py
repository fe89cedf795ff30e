"""Exception tree at the drop-in boundary.  Same names as the reference
(`astrophot/errors/*.py`) so user code that catches them keeps working."""

__all__ = (
    "AstroPhotError", "NameNotAllowed", "SpecificationConflict", "OptimizeStop",
    "InvalidWindow", "ConflicingWCS", "InvalidData", "InvalidImage", "InvalidWCS",
    "InvalidModel", "InvalidTarget", "UnrecognizedModel", "InvalidParameter",
    "NativeLibraryError",
)


class AstroPhotError(Exception):
    """Root of every error this package raises on purpose."""


class NameNotAllowed(AstroPhotError):
    """Object name uses a reserved character."""


class SpecificationConflict(AstroPhotError):
    """Inputs are contradictory, ambiguous, or name an unknown mode."""


class OptimizeStop(AstroPhotError):
    """An optimiser cannot continue (internal to LM)."""


class InvalidWindow(AstroPhotError):
    """A window specification cannot be interpreted."""


class ConflicingWCS(InvalidWindow):
    """Two windows disagree on their world coordinate system."""


class InvalidData(AstroPhotError):
    """Pixel data of the wrong shape or kind."""


class InvalidImage(AstroPhotError):
    """An image object of the wrong kind was supplied."""


class InvalidWCS(AstroPhotError):
    """Bad WCS specification."""


class InvalidModel(AstroPhotError):
    """A model was combined in a way that is not allowed."""


class InvalidTarget(AstroPhotError):
    """A model was handed something that is not a Target_Image."""


class UnrecognizedModel(AstroPhotError):
    """``model_type`` string names no known model."""


class InvalidParameter(AstroPhotError):
    """Parameter value outside its limits, or a cyclic parameter graph."""


class NativeLibraryError(AstroPhotError):
    """The sm_100a shared library is missing, failed to load, or returned an
    error status.  There is deliberately no CPU fallback."""
