"""Unit, profile and coordinate conversions users of the reference call from their scripts
(`utils/conversions/{units,functions,coordinates}.py`).  Host helpers; nothing here is on the device path."""
from . import coordinates, functions, units  # noqa: F401
