"""`utils.script_util.create_gaussian_diffusion` of the reference (utils/script_util.py:7-61) lives in model/respace.py."""
from ..model.respace import create_gaussian_diffusion  # noqa: F401
