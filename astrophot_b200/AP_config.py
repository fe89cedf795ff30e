"""Run-time configuration, mirroring the module globals of the reference
(`astrophot/AP_config.py:7-20`): every layer reads ``ap_dtype`` / ``ap_device``
when it creates a tensor.  Pixel data lives on ``ap_device``; the parameter
DAG and window geometry are host-side bookkeeping."""
import logging
import sys

import torch

__all__ = ["ap_dtype", "ap_device", "ap_verbose", "ap_logger", "set_logging_output"]

ap_dtype = torch.float64
ap_device = "cuda:0" if torch.cuda.is_available() else "cpu"
ap_verbose = 0

ap_logger = logging.getLogger("astrophot_b200")
ap_logger.setLevel(logging.INFO)
if not ap_logger.handlers:
    _h = logging.StreamHandler(sys.stdout)
    _h.setFormatter(logging.Formatter("%(message)s"))
    ap_logger.addHandler(_h)


def set_logging_output(stdout=True, filename=None, **kwargs):
    """Choose where log records go (reference: `AP_config.py:23-66`)."""
    for h in list(ap_logger.handlers):
        ap_logger.removeHandler(h)
    if stdout:
        h = logging.StreamHandler(sys.stdout)
        h.setLevel(kwargs.get("stdout_level", logging.INFO))
        h.setFormatter(logging.Formatter(kwargs.get("stdout_formatter", "%(message)s")))
        ap_logger.addHandler(h)
    if filename is not None:
        h = logging.FileHandler(filename)
        h.setLevel(kwargs.get("filename_level", logging.INFO))
        h.setFormatter(
            logging.Formatter(kwargs.get("filename_formatter", "%(asctime)s:%(levelname)s: %(message)s"))
        )
        ap_logger.addHandler(h)
    if not ap_logger.handlers:
        ap_logger.addHandler(logging.NullHandler())
